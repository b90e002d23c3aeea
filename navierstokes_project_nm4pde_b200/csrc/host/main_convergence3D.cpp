// main_convergence3D.cpp -- Ethier-Steinman convergence driver (Navier-Stokes/src/main_convergence3D.cpp):
// one time step on each mesh, L2 / H1 velocity errors, convergence.csv and the rate table.
// argv[1..] = meshes (default: generated Kuhn cubes with the reference's h values where they exist).
#include <cmath>
#include <iomanip>
#include <vector>

#include "driver_common.hpp"

int main(int argc, char *argv[])
{
  std::vector<std::string> meshes = {"gen:cube:5", "gen:cube:10", "gen:cube:20"}; // h = 0.4, 0.2, 0.1 (side 2)
  std::vector<double> h_vals = {1.0 / 2.5, 1.0 / 5.0, 1.0 / 10.0};               // main_convergence3D.cpp:20-23
  if (argc > 1) {
    meshes.assign(argv + 1, argv + argc);
    h_vals.clear();
    for (size_t i = 0; i < meshes.size(); ++i) h_vals.push_back(std::ldexp(1.0, -int(i)));
  }
  const unsigned int degree_velocity = 2, degree_pressure = 1;
  const double T = 0.0003, deltat = 0.0004;                                       // main_convergence3D.cpp:35-36

  dealii::Timer timer;
  timer.restart();
  std::vector<double> errors_L2, errors_H1;
  unsigned int mpi_rank = 0;
  try {
    Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);                        // main_convergence3D.cpp:13
    mpi_rank = Utilities::MPI::this_mpi_process(mpi_init);
    if (rendezvous_selftest(mpi_init)) return 0;
    std::ofstream convergence_file;
    if (mpi_rank == 0) {
      convergence_file.open("convergence.csv");
      convergence_file << "h,eL2,eH1" << std::endl;
    }
    for (unsigned int i = 0; i < meshes.size(); ++i) {
      NavierStokes problem(NavierStokes::Variant::Convergence3D, meshes[i], degree_velocity, degree_pressure, T, deltat);
      apply_env(problem, mpi_init);
      problem.setup();
      problem.solve();
      const double error_L2 = problem.compute_error(VectorTools::L2_norm);
      const double error_H1 = problem.compute_error(VectorTools::H1_norm);
      errors_L2.push_back(error_L2);
      errors_H1.push_back(error_H1);
      if (mpi_rank == 0) convergence_file << h_vals[i] << "," << error_L2 << "," << error_H1 << std::endl;
      timer.stop();
    }
  } catch (const std::exception &e) {
    std::cerr << "convergence: " << e.what() << std::endl;
    return 1;
  }
  if (mpi_rank != 0) return 0;
  std::cout << "Time taken to solve ENTIRE Navier Stokes problem: " << timer.wall_time() << " seconds" << std::endl;
  // ConvergenceTable::evaluate_all_convergence_rates(reduction_rate_log2)
  std::cout << "h        L2          rate   H1          rate" << std::endl;
  for (size_t i = 0; i < errors_L2.size(); ++i) {
    std::cout << std::fixed << std::setprecision(4) << h_vals[i] << "  " << std::scientific << std::setprecision(4) << errors_L2[i] << "  ";
    if (i) std::cout << std::fixed << std::setprecision(2) << std::log2(errors_L2[i - 1] / errors_L2[i]); else std::cout << "   -";
    std::cout << "  " << std::scientific << std::setprecision(4) << errors_H1[i] << "  ";
    if (i) std::cout << std::fixed << std::setprecision(2) << std::log2(errors_H1[i - 1] / errors_H1[i]); else std::cout << "   -";
    std::cout << std::endl;
  }
  return 0;
}
