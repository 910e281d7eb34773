// oracle/f1_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's OWN tight-binding / exciton-transfer translation units
// (/root/reference/src/exciton_transfer/{cnt,exciton_transfer}.cpp, compiled from where they lie against
// oracle/arma_full/armadillo), so that tests can obtain from the reference's code
//   * the exciton dispersion of a tube                      (cnt::calculate_exciton_dispersion, cnt.cpp)
//   * first-order transfer rates between two tubes          (exciton_transfer::first_order, exciton_transfer.cpp:395)
// which together are what monte_carlo::create_davoody_scatt_table (monte_carlo.cpp:64-153) tabulates.
//
// The reference writes result files below the directory it is given and below $HOME/research; the driver points both at
// a scratch directory of the caller's choosing.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "exciton_transfer/cnt.h"
#include "exciton_transfer/exciton_transfer.h"

namespace {
std::vector<std::unique_ptr<cnt>> g_cnts;
std::string                       g_error;

const cnt::exciton_struct* pick(const cnt& c, int which) {
  switch (which) {
    case 0: return &c.A1();
    case 1: return &c.A2_singlet();
    case 2: return &c.A2_triplet();
  }
  return nullptr;
}
}  // namespace

extern "C" {

const char* f1_last_error() { return g_error.c_str(); }

// builds the tube (n,m) of `length` cnt unit cells and runs the whole exciton calculation; returns a handle >= 0
int f1_cnt_create(int n, int m, int length, const char* scratch_dir) {
  try {
    setenv("HOME", scratch_dir, 1);
    nlohmann::json j;
    j["chirality"] = {n, m};
    j["length"] = {length, "cnt unit cells"};
    j["keep old results"] = false;
    g_cnts.emplace_back(new cnt(j, std::string(scratch_dir) + "/exciton_energy"));
    g_cnts.back()->calculate_exciton_dispersion();
    return (int)g_cnts.size() - 1;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

double f1_radius(int h) { return g_cnts[h]->radius(); }
double f1_length_in_meter(int h) { return g_cnts[h]->length_in_meter(); }
double f1_Au(int h) { return g_cnts[h]->Au(); }
int    f1_cells(int h) { return (int)g_cnts[h]->pos_u_3d().n_rows; }

// which: 0 = A1, 1 = A2 singlet, 2 = A2 triplet.  dims = {nk_cm, n_principal, nk_c, ik_cm_range[0], ik_cm_range[1]}
int f1_exciton_dims(int h, int which, int* dims) {
  const cnt::exciton_struct* ex = pick(*g_cnts[h], which);
  if (!ex) return -1;
  dims[0] = ex->nk_cm;
  dims[1] = ex->n_principal;
  dims[2] = ex->nk_c;
  dims[3] = ex->ik_cm_range[0];
  dims[4] = ex->ik_cm_range[1];
  return 0;
}
// energy(ik_cm_idx, n) row-major into out[nk_cm * n_principal]
int f1_exciton_energy(int h, int which, double* out) {
  const cnt::exciton_struct* ex = pick(*g_cnts[h], which);
  if (!ex) return -1;
  for (int i = 0; i < ex->nk_cm; ++i)
    for (int n = 0; n < ex->n_principal; ++n) out[(size_t)i * ex->n_principal + n] = ex->energy(i, n);
  return 0;
}

double f1_first_order(int donor, int acceptor, double z_shift, double axis_shift_1, double axis_shift_2, double theta) {
  try {
    exciton_transfer ex_transfer(*g_cnts[donor], *g_cnts[acceptor]);
    return ex_transfer.first_order(z_shift, {axis_shift_1, axis_shift_2}, theta, false);
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

// the same through the reference's other constructor (exciton_transfer.h:51-98), which takes temperature [Kelvin] and
// broadening [meV] from JSON and finds its tubes by name in a vector; the copies made for that vector keep pointing at
// the originals' band data, which stay alive in g_cnts
double f1_first_order_at(int donor, int acceptor, double temperature, double broadening_mev, double z_shift, double axis_shift_1,
                         double axis_shift_2, double theta) {
  try {
    if (acceptor != donor && g_cnts[donor]->name() == g_cnts[acceptor]->name())
      throw std::invalid_argument("the reference finds its tubes by name: two different tubes of one chirality cannot be told apart");
    std::vector<cnt> pool;
    pool.reserve(2);
    pool.emplace_back(*g_cnts[donor]);
    if (acceptor != donor) pool.emplace_back(*g_cnts[acceptor]);
    nlohmann::json j;
    j["cnt 1"] = g_cnts[donor]->name();
    j["cnt 2"] = g_cnts[acceptor]->name();
    j["temperature"] = {temperature, "Kelvin"};
    j["broadening factor"] = {broadening_mev, "meV"};
    j["keep old results"] = false;
    const char* home = getenv("HOME");
    exciton_transfer ex_transfer(j, pool, std::string(home ? home : ".") + "/exciton_transfer_json");
    return ex_transfer.first_order(z_shift, {axis_shift_1, axis_shift_2}, theta, false);
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

// rate[((i_th * nz + i_z) * n1 + i_1) * n2 + i_2], the loop nest of monte_carlo.cpp:114-137
int f1_table(int donor, int acceptor, int nth, const double* theta, int nz, const double* z, int n1, const double* a1, int n2,
             const double* a2, double* rate) {
  try {
    exciton_transfer ex_transfer(*g_cnts[donor], *g_cnts[acceptor]);
    // the reference's own team: "#pragma omp parallel" + "#pragma omp for" over theta (monte_carlo.cpp:107-112)
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nth; ++i)
      for (int k = 0; k < nz; ++k)
        for (int p = 0; p < n1; ++p)
          for (int q = 0; q < n2; ++q)
            rate[(((size_t)i * nz + k) * n1 + p) * n2 + q] = ex_transfer.first_order(z[k], {a1[p], a2[q]}, theta[i], false);
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

void f1_reset() { g_cnts.clear(); }
}
