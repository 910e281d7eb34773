// oracle/t0_link_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Link-time stubs for the three out-of-scope symbols that the reference's monte_carlo.cpp mentions on its
// "rate type":"davoody" branch (monte_carlo.cpp:42-46, 80-84, 123).  The tight-binding / Bethe-Salpeter
// solver behind them (cnt.cpp, exciton_transfer.cpp) needs Armadillo's complex cubes and LAPACK zheev,
// which do not exist in this image, so the T0 reference build supports "forster" and "wong" tables only and
// aborts loudly if the davoody branch is ever reached.
#include <cstdio>
#include <cstdlib>

#include "exciton_transfer/cnt.h"
#include "exciton_transfer/exciton_transfer.h"

static void unsupported(const char* what) {
  std::fprintf(stderr, "[oracle/T0] %s needs the davoody rate path, which is not built in this oracle\n", what);
  std::abort();
}

void cnt::calculate_exciton_dispersion() { unsupported("cnt::calculate_exciton_dispersion"); }

double exciton_transfer::first_order(const double&, const std::array<double, 2>, const double&, const bool&) {
  unsupported("exciton_transfer::first_order");
  return 0;
}

void exciton_transfer::save_atom_locations(path_t, const std::array<double, 2>&, const double&, const double&, std::string) {
  unsupported("exciton_transfer::save_atom_locations");
}
