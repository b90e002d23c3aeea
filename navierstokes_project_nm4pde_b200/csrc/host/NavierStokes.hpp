// NavierStokes.hpp -- C++ host side above the C ABI: the reference's `NavierStokes` class with
// the same public surface (`setup()`, `solve()`, `compute_error()`, the public result vectors;
// Navier-Stokes/include/NavierStokes2D.hpp:84-119, Convergence3D.hpp:268-292) and the same
// protected methods (`assemble`, `assemble_time_step`, `solve_time_step`, `compute_forces`),
// each of which is one call through include/nsb.h.  The reference has three copies of the class
// (2D cylinder, 3D cylinder, Ethier-Steinman); here the copy is chosen by `Variant`.
//
// deal.II is not available, so the pieces of setup() the reference gets from it (GridIn,
// DoFHandler, boundary interpolation) come from the nsh_* host entry points of the same library.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../../include/nsb.h"

namespace VectorTools { enum NormType { L2_norm, H1_norm }; }

class NavierStokes
{
public:
  enum class Variant { Cylinder2D = NSB_VARIANT_2D, Cylinder3D = NSB_VARIANT_3D, Convergence3D = NSB_VARIANT_CONV };

  // NavierStokes(mesh_file_name, degree_velocity, degree_pressure, T, deltat[, test_case]).
  // mesh_file_name is a Gmsh .msh (v2 / v4.1 ASCII) or, because the reference ships only .geo
  // scripts, a generator spec "gen:cylinder2d:<s>", "gen:cylinder3d:<s>:<nz>", "gen:cube:<n>".
  NavierStokes(Variant variant, const std::string &mesh_file_name, unsigned degree_velocity, unsigned degree_pressure,
               double T, double deltat, int test_case = 2);
  ~NavierStokes();
  NavierStokes(const NavierStokes &) = delete;
  NavierStokes &operator=(const NavierStokes &) = delete;

  // One process per GPU -- the reference under `mpirun -n P` (main3D.cpp:9,28; the partition of
  // NavierStokes2D.cpp:16-19).  Call before setup(); `allgather(mine, all, bytes)` is MPI_Allgather of small
  // blobs (host/rendezvous.hpp in the drivers) and is used during setup() only: per-step traffic is
  // device-to-device inside the library.  Only rank 0 prints (pcout).
  using AllGather = std::function<void(const void *mine, void *all, size_t bytes)>;
  void set_parallel(int nranks, int rank, AllGather allgather);
  int this_mpi_process() const { return rank; }
  int n_mpi_processes() const { return nranks; }
  const char *transport() const { return transport_name; } // "single", "nccl" or "p2p" (peer memory over NVLink)

  void setup();
  void solve();
  double compute_error(const VectorTools::NormType &norm_type); // Convergence3D.cpp:766-794

  std::vector<double> vec_drag, vec_lift, vec_drag_coeff, vec_lift_coeff; // never filled by the reference either
  std::vector<double> time_prec, time_solve;
  std::vector<int> gmres_iterations;
  std::vector<std::array<double, 2>> coefficients_history; // (c_d, c_l) of every step that computed forces
  double pressure_difference = 0.0;                        // P(A) - P(B) of the last compute_pressure_difference()

  // knobs the reference hard-codes; the drivers override them from the environment
  int max_steps = -1;        // stop after this many steps (reference: run to T)
  int ilu_ordering = 0;      // 0: reference replay, 1 / 2 / 3: throughput orderings (nsb_params)
  int orthogonalisation = 0; // 0: modified Gram-Schmidt as SolverGMRES, 1: batched (throughput mode)
  int ilu_ordering_schur = -1; // ordering of the Schur-complement factors (-1: same as ilu_ordering)
  int device = 0;
  double forces_after = 0.1; // NavierStokes3D.cpp:728 computes forces only for time > 0.1
  bool write_output = false; // VTU per step (2D) / every 20 steps (3D), gmres.csv, coeff_2.csv as the reference
  bool verbose = true;

  // local vector [u (owned, ghost) | p (owned, ghost)]; the whole vector on one rank
  const std::vector<double> &get_solution() { sync_solution(); return solution; }
  int n_dofs() const { return N_global; }

protected:
  void assemble(const double &time);           // NavierStokes2D.cpp:164-357
  void assemble_time_step(const double &time); // NavierStokes2D.cpp:361-527
  void solve_time_step(double time);           // NavierStokes2D.cpp:530-639
  std::vector<double> compute_forces();        // NavierStokes2D.cpp:752-859 / NavierStokes3D.cpp:744-840
  void compute_pressure_difference();          // NavierStokes2D.cpp:862-936 / NavierStokes3D.cpp:843-923
  void output(unsigned time_step, const std::vector<double> &coeff) const; // NavierStokes2D.cpp:642-695
  void dirichlet_values(double time, std::vector<double> &vals) const;
  void neumann_rhs(double time, std::vector<double> &rhs) const; // Convergence3D.cpp:309-330
  void initial_condition(std::vector<double> &x) const;          // NavierStokes2D.cpp:708
  void check(int rc, const char *what) const;
  void sync_solution() const; // device -> host copy of the solution, only when a host consumer needs it
  bool point_value(const double *x, double *out) const; // VectorTools::point_value + sum over the ranks
  const double *node_xyz(int local_node) const;
  const double *p_xyz(int local_p) const;

  Variant variant;
  int dim;
  std::string mesh_file_name;
  double T, deltat;
  int test_case;
  double nu;
  const double rho = 1.0;

  nsh_mesh mesh = nullptr;
  nsh_dofs dofs = nullptr;
  nsh_local local = nullptr; // subdomain of this rank (nranks > 1)
  nsb_handle engine = nullptr;
  int nranks = 1, rank = 0;
  AllGather allgather;
  const char *transport_name = "single";
  // sizes and index arrays below are LOCAL to this rank (owned + ghost); on one rank local == global
  int n_nodes = 0, n_u = 0, n_p = 0, N = 0, N_global = 0, dpc = 0, n_cells = 0;
  const int32_t *cell_dofs = nullptr;  // [n_cells][dpc]
  const double *cell_coords = nullptr; // [n_cells][dim+1][dim]
  const int32_t *node_gid = nullptr, *p_gid = nullptr; // local -> global (null: identity)
  const int32_t *cell_gid = nullptr, *cell_owner = nullptr, *g2l_cell = nullptr;
  std::vector<int32_t> dir_nodes, dir_rows;
  std::vector<char> dir_is_inlet;
  // boundary id 3 faces lying in local cells: local cell, local face, and whether this rank owns the cell
  std::vector<int32_t> obstacle_cells, obstacle_faces;
  std::vector<char> obstacle_owned;
  mutable std::vector<double> solution; // host mirror of the device solution (see sync_solution)
  mutable bool solution_stale = false;
  double time_now = 0.0;
};

// InletVelocity::getMeanVelocity and the exact solution are exposed for the drivers / tests
double inlet_mean_velocity(int dim, int test_case, double t);
void ethier_steinman(const double x[3], double t, double u[3], double &p, double grad[3][3]);
