// driver_common.hpp -- what the reference drivers take from deal.II (Utilities::MPI::MPI_InitFinalize,
// Timer, ConvergenceTable) reduced to what the three mains use, plus the override channel the
// reference lacks (every constant there is a compile-time literal):
//   NSB_MAX_STEPS=<n>   stop after n time steps      NSB_T=<T>  final time
//   NSB_ILU_ORDERING=1  multicolour ILU(0) (throughput mode)     NSB_DEVICE=<id>
//   NSB_OUTPUT=1        write the reference's side outputs (.vtu, gmres.csv, coeff_2.csv); off by default
#pragma once
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "NavierStokes.hpp"

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

inline void apply_env(NavierStokes &problem)
{
  problem.max_steps = env_int("NSB_MAX_STEPS", -1);
  problem.ilu_ordering = env_int("NSB_ILU_ORDERING", 0);
  problem.device = env_int("NSB_DEVICE", 0);
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
