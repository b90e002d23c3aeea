// driver_common.hpp -- what the reference drivers take from deal.II (Utilities::MPI::MPI_InitFinalize,
// Timer, ConvergenceTable) reduced to what the three mains use, plus the override channel the
// reference lacks (every constant there is a compile-time literal):
//   NSB_MAX_STEPS=<n>   stop after n time steps      NSB_T=<T>  final time
//   NSB_ILU_ORDERING=1  multicolour ILU(0) (throughput mode)     NSB_DEVICE=<id>
//   NSB_OUTPUT=1        write the reference's side outputs (.vtu, gmres.csv, coeff_2.csv); off by default
//   NSB_ORTHOGONALISATION=1  batched Gram-Schmidt (throughput mode)   NSB_ILU_ORDERING_SCHUR=<k>  Schur factors' ordering
// One process per GPU (the reference under mpirun): start the binary N times with RANK / WORLD_SIZE /
// LOCAL_RANK / MASTER_ADDR set -- `python -m torch.distributed.run --no-python --nproc-per-node N <binary> <mesh>`
// or scripts/nsb_launch.sh -- see rendezvous.hpp.  NSB_P2P=0 keeps NCCL instead of peer memory.
#pragma once
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "NavierStokes.hpp"
#include "rendezvous.hpp"

// Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv) + this_mpi_process (main3D.cpp:9,28)
namespace Utilities { namespace MPI {
class MPI_InitFinalize
{
public:
  MPI_InitFinalize(int &, char **&) {}
  Rendezvous comm;
};
inline unsigned this_mpi_process(const MPI_InitFinalize &m) { return unsigned(m.comm.rank()); }
inline unsigned n_mpi_processes(const MPI_InitFinalize &m) { return unsigned(m.comm.size()); }
} } // namespace Utilities::MPI

namespace dealii {
class Timer
{
public:
  void restart() { t0 = std::chrono::steady_clock::now(); running = true; }
  void stop() { if (running) acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); running = false; }
  double wall_time() const { return acc; }
private:
  std::chrono::steady_clock::time_point t0;
  double acc = 0.0;
  bool running = false;
};
} // namespace dealii

inline double env_double(const char *name, double def) { const char *e = std::getenv(name); return e ? std::atof(e) : def; }
inline int env_int(const char *name, int def) { const char *e = std::getenv(name); return e ? std::atoi(e) : def; }

// NSB_RDV_SELFTEST=1: exercise the launcher contract and the all-gather without touching a GPU, then exit
inline bool rendezvous_selftest(Utilities::MPI::MPI_InitFinalize &mpi)
{
  if (env_int("NSB_RDV_SELFTEST", 0) == 0) return false;
  Rendezvous &comm = mpi.comm;
  for (size_t bytes : {size_t(1), size_t(64), size_t(128), size_t(100000)}) {
    std::vector<unsigned char> mine(bytes), all(bytes * size_t(comm.size()));
    for (size_t i = 0; i < bytes; ++i) mine[i] = (unsigned char)(31 * comm.rank() + 7 * i + bytes);
    comm.allgather(mine.data(), all.data(), bytes);
    for (int r = 0; r < comm.size(); ++r)
      for (size_t i = 0; i < bytes; ++i)
        if (all[size_t(r) * bytes + i] != (unsigned char)(31 * r + 7 * i + bytes)) throw std::runtime_error("rendezvous self-test: wrong data");
  }
  comm.barrier();
  std::cout << "rendezvous ok: rank " << comm.rank() << " of " << comm.size() << ", local rank " << comm.local_rank() << std::endl;
  return true;
}

inline void apply_env(NavierStokes &problem, Utilities::MPI::MPI_InitFinalize &mpi)
{
  Rendezvous &comm = mpi.comm;
  if (comm.size() > 1)
    problem.set_parallel(comm.size(), comm.rank(), [&comm](const void *mine, void *all, size_t bytes) { comm.allgather(mine, all, bytes); });
  problem.max_steps = env_int("NSB_MAX_STEPS", -1);
  problem.ilu_ordering = env_int("NSB_ILU_ORDERING", 0);
  problem.orthogonalisation = env_int("NSB_ORTHOGONALISATION", 0);
  problem.ilu_ordering_schur = env_int("NSB_ILU_ORDERING_SCHUR", -1);
  problem.device = env_int("NSB_DEVICE", comm.local_rank());
  problem.forces_after = env_double("NSB_FORCES_AFTER", 0.1);
  problem.write_output = env_int("NSB_OUTPUT", 0) != 0;
}

inline int write_forces_csv(const std::string &output_filename, const NavierStokes &problem, double deltat)
{
  std::ofstream outputFile(output_filename);
  if (!outputFile.is_open()) { std::cerr << "Error opening output file" << std::endl; return -1; }
  outputFile << "Iteration, Drag, Lift, Coeff Drag, CoeffLift, time prec, time solve" << std::endl;
  // the reference bounds this loop by vec_drag.size(), which it never fills (main2D.cpp:52): header only
  for (size_t ite = 0; ite < problem.vec_drag.size(); ite++)
    outputFile << ite * deltat << ", " << problem.vec_drag[ite] << ", " << problem.vec_lift_coeff[ite] << ", "
               << problem.vec_drag_coeff[ite] << ", " << problem.vec_lift_coeff[ite] << ", " << problem.time_prec[ite] << ", "
               << problem.time_solve[ite] << std::endl;
  return 0;
}
