// main2D.cpp -- driver of the 2D flow past a cylinder (Navier-Stokes/src/main2D.cpp): same constants,
// same calls; argv[1] = mesh (.msh or "gen:cylinder2d:<s>").
#include "driver_common.hpp"

int main(int argc, char *argv[])
{
  int test_case = 2;                                                                       // main2D.cpp:7
  const std::string mesh_file_name = argc > 1 ? argv[1] : "../mesh/Cylinder2D.msh";        // main2D.cpp:14
  const unsigned int degree_velocity = 2, degree_pressure = 1;                             // Taylor-Hood
  const double T = env_double("NSB_T", 8.0), deltat = 0.01;                                // main2D.cpp:21-22

  dealii::Timer timer;
  timer.restart();
  try {
    Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);                                 // main2D.cpp:9
    const unsigned int mpi_rank = Utilities::MPI::this_mpi_process(mpi_init);
    if (rendezvous_selftest(mpi_init)) return 0;
    NavierStokes problem(NavierStokes::Variant::Cylinder2D, mesh_file_name, degree_velocity, degree_pressure, T, deltat,
                         test_case);
    apply_env(problem, mpi_init);
    problem.setup();
    problem.solve();
    timer.stop();
    if (mpi_rank != 0) return 0;
    std::cout << "Time taken to solve ENTIRE Navier Stokes problem: " << timer.wall_time() << " seconds" << std::endl;
    return write_forces_csv("forces_results_2D_2case.csv", problem, deltat);
  } catch (const std::exception &e) {
    std::cerr << "navier_stokes2D: " << e.what() << std::endl;
    return 1;
  }
}
