// Driver with the same five lines of work as the reference's src/main.cpp (which also compiles against
// NavierStokesSolver.hpp unchanged — see tests/test_host_shim.py); T and deltat may be overridden
// by NS_T / NS_DELTAT, and --history prints the Newton/GMRES record as JSON lines.
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>

#include "NavierStokesSolver.hpp"

int main(int argc, char *argv[]) {
  Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);
  const unsigned int degree_velocity = 2, degree_pressure = 1;
  const double T = std::getenv("NS_T") ? std::atof(std::getenv("NS_T")) : 1.0;
  const double deltat = std::getenv("NS_DELTAT") ? std::atof(std::getenv("NS_DELTAT")) : 0.05;
  try {
    NavierStokesSolver problem(degree_velocity, degree_pressure, T, deltat);
    problem.setup();
    problem.solve();
    if (argc > 1 && !std::strcmp(argv[1], "--history") && Utilities::MPI::this_mpi_process() == 0)
      for (const auto &r : problem.history())
        std::cout << "{\"time_step\": " << r.time_step << ", \"newton\": " << r.newton_iteration << ", \"residual\": "
                  << std::setprecision(17) << r.residual_norm << ", \"gmres\": " << r.gmres_steps << "}" << std::endl;
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
