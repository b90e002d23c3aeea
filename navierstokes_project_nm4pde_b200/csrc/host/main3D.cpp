// main3D.cpp -- driver of the 3D flow past a cylinder (Navier-Stokes/src/main3D.cpp).
#include "driver_common.hpp"

int main(int argc, char *argv[])
{
  int test_case = 2;                                                                       // main3D.cpp:8
  const std::string mesh_file_name = argc > 1 ? argv[1] : "../mesh/Parallelepiped3D.msh";  // main3D.cpp:31
  const unsigned int degree_velocity = 2, degree_pressure = 1;
  const double T = env_double("NSB_T", 4.0), deltat = 0.0002;                              // main3D.cpp:37-38

  dealii::Timer timer;
  timer.restart();
  try {
    Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);                                 // main3D.cpp:9
    const unsigned int mpi_rank = Utilities::MPI::this_mpi_process(mpi_init);
    if (rendezvous_selftest(mpi_init)) return 0;
    NavierStokes problem(NavierStokes::Variant::Cylinder3D, mesh_file_name, degree_velocity, degree_pressure, T, deltat,
                         test_case);
    apply_env(problem, mpi_init);
    problem.setup();
    problem.solve();
    timer.stop();
    if (mpi_rank != 0) return 0;
    std::cout << "Time taken to solve ENTIRE Navier Stokes problem: " << timer.wall_time() << " seconds" << std::endl;
    return write_forces_csv("forces_results_3D_2case.csv", problem, deltat);
  } catch (const std::exception &e) {
    std::cerr << "navier_stokes3D: " << e.what() << std::endl;
    return 1;
  }
}
