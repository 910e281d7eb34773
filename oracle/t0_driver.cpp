// oracle/t0_driver.cpp -- TEST INFRASTRUCTURE ONLY (the "T0" oracle of SURVEY.md §8c).
//
// A thin extern "C" window onto the reference's OWN code: this file is compiled together with the unmodified
// translation units /root/reference/src/monte_carlo/{monte_carlo,particle,scatterer}.cpp (see oracle/Makefile)
// against oracle/arma_standin/armadillo, into oracle/_ref/libt0.so.  Nothing of the reference is copied into
// this repository; the sources are compiled from where they lie.
//
// What it adds on top of the reference objects:
//   * access to private members of mc::monte_carlo / mc::particle (via the access-specifier macro below; member
//     layout is unaffected on GCC, and every reference TU sees the same class definitions);
//   * an interposed rand() that forwards to glibc random() -- which is exactly what glibc's rand() does -- and
//     logs every draw together with the exciton that consumed it, so the draws can be replayed on the GPU
//     (SURVEY.md App. A.7);
//   * t0_kubo_step_logged(): the 12-line particle loop of monte_carlo::kubo_step (monte_carlo.cpp:319-342) written
//     out so the consuming exciton is known for every draw; it calls the reference's particle::step,
//     update_delta_pos etc.  tests/test_oracle_t0.py proves it equal, bit for bit, to the verbatim kubo_step();
//   * t0_contact_iteration_logged(): monte_carlo::step and monte_carlo::repopulate (monte_carlo.h:343-355, 458-491) written
//     out with an id carried along every particle through repopulate's swaps, so that the draws of the contact loop are
//     known per exciton (ids in order of birth); the same test proves it equal to the verbatim loop.
#include <array>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <list>
#include <map>
#include <regex>
#include <thread>
#include <chrono>
#include <iomanip>
#include <algorithm>
#include <experimental/filesystem>
#include <omp.h>
#include <armadillo>
#include "../lib/json.hpp"

#define private public
#include "monte_carlo/monte_carlo.h"
#undef private

namespace {

std::unique_ptr<mc::monte_carlo> g_sim;
std::string                      g_error;

// ---- rand() log -----------------------------------------------------------------------------------------------
bool                              g_log_on = false;
int64_t                           g_cur = -1;  // exciton currently consuming draws (-1: unattributed)
std::vector<std::vector<int32_t>> g_draws;     // per exciton, in consumption order
int64_t                           g_total_draws = 0;
int                               g_threads = 1;  // OpenMP team size for the reference's parallel regions
std::vector<int32_t>              g_seq;          // draws of a stretch whose consumers are known by position (g_cur == -2)
std::vector<int64_t>              g_ids;          // contact mode: id (order of birth) of _particle_list[i]
int64_t                           g_next_id = 0;

struct cout_silencer {
  std::streambuf* old;
  std::ostringstream sink;
  cout_silencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~cout_silencer() { std::cout.rdbuf(old); }
};

inline const mc::scatterer* site0() { return &g_sim->_all_scat_list[0]; }

void attribute_to(int64_t i) {
  g_cur = i;
  if (g_log_on && i >= 0 && (size_t)i >= g_draws.size()) g_draws.resize(i + 1);
}

}  // namespace

// glibc: int rand(void) { return (int) __random(); }   -- same stream, now observable.
extern "C" int rand(void) {
  const int r = (int)random();
  __atomic_fetch_add(&g_total_draws, 1, __ATOMIC_RELAXED);
  if (g_log_on && g_cur >= 0) g_draws[g_cur].push_back(r);
  if (g_log_on && g_cur == -2) g_seq.push_back(r);
  return r;
}

extern "C" {

const char* t0_last_error() { return g_error.c_str(); }

void t0_srand(unsigned seed) { srandom(seed); }  // glibc: srand is an alias of srandom
int64_t t0_total_draws() { return g_total_draws; }

// main.cpp:41-66 for a JSON file that holds an "exciton monte carlo" block; runs kubo_init().
int t0_open(const char* json_path, unsigned seed) {
  try {
    cout_silencer quiet;
    omp_set_num_threads(g_threads);
    srandom(seed);
    std::ifstream  f(json_path);
    nlohmann::json j;
    f >> j;
    nlohmann::json json_mc = j["exciton monte carlo"];
    g_sim.reset(new mc::monte_carlo(json_mc));
    g_sim->_time = 0;  // never initialised by the reference (monte_carlo.h:45)
    g_sim->kubo_init();
    g_draws.clear();
    g_cur = -1;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// contact-mode initialisation: monte_carlo::init() (monte_carlo.h:157-195), creates the particle list too.
int t0_open_contacts(const char* json_path, unsigned seed) {
  try {
    cout_silencer quiet;
    omp_set_num_threads(g_threads);
    srandom(seed);
    std::ifstream  f(json_path);
    nlohmann::json j;
    f >> j;
    nlohmann::json json_mc = j["exciton monte carlo"];
    g_sim.reset(new mc::monte_carlo(json_mc));
    g_sim->_time = 0;
    g_draws.clear();
    g_cur = g_log_on ? -2 : -1;  // logging: the creation draws go to g_seq (see t0_contacts_attribute_creation)
    g_sim->init();
    g_cur = -1;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// the same, with the creation draws of the initial population attributed: create_particles (monte_carlo.h:274-316) draws,
// per particle and in list order, a site, a free-flight time (again while the draw is zero) and a heading.  Returns the
// population, or -1 (also when a zero draw makes the positions ambiguous -- pick another seed).
int64_t t0_open_contacts_logged(const char* json_path, unsigned seed) {
  g_log_on = true;
  g_seq.clear();
  if (t0_open_contacts(json_path, seed) != 0) return -1;
  return 0;
}
int64_t t0_contacts_attribute_creation() {
  const int64_t P0 = (int64_t)g_sim->_particle_list.size();
  if ((int64_t)g_seq.size() != 3 * P0) {
    g_error = "creation consumed " + std::to_string(g_seq.size()) + " draws for " + std::to_string(P0) + " particles";
    return -1;
  }
  g_draws.assign((size_t)P0, {});
  g_ids.resize((size_t)P0);
  for (int64_t i = 0; i < P0; ++i) {
    g_draws[(size_t)i].assign(g_seq.begin() + 3 * i, g_seq.begin() + 3 * i + 3);
    g_ids[(size_t)i] = i;
  }
  g_next_id = P0;
  g_seq.clear();
  g_cur = -1;
  return P0;
}

void t0_close() { g_sim.reset(); }

void t0_set_threads(int n) {
  g_threads = n > 0 ? n : 1;
  omp_set_num_threads(g_threads);
}

// ---- read-only views of the set-up state ------------------------------------------------------------------------
int64_t t0_num_sites() { return (int64_t)g_sim->_all_scat_list.size(); }

void t0_sites(double* pos /*[3][N]*/, double* orient /*[3][N]*/, int32_t* left, int32_t* right, double* max_rate,
              double* inv_max_rate) {
  const auto&   s = g_sim->_all_scat_list;
  const int64_t N = (int64_t)s.size();
  for (int64_t i = 0; i < N; ++i) {
    for (int c = 0; c < 3; ++c) {
      pos[c * N + i] = s[i].pos(c);
      orient[c * N + i] = s[i].orientation(c);
    }
    left[i] = s[i].left;
    right[i] = s[i].right;
    max_rate[i] = s[i]._max_rate;
    inv_max_rate[i] = s[i]._inverse_max_rate;
  }
}

void t0_table_dims(int32_t* dims /*[4]: theta,z,a1,a2*/) {
  const auto& t = g_sim->_scat_table;
  dims[0] = (int32_t)t.theta.n_elem;
  dims[1] = (int32_t)t.z_shift.n_elem;
  dims[2] = (int32_t)t.axis_shift_1.n_elem;
  dims[3] = (int32_t)t.axis_shift_2.n_elem;
}

void t0_table(double* theta, double* z, double* a1, double* a2, double* rates /*[th][z][a1][a2]*/) {
  const auto& t = g_sim->_scat_table;
  for (unsigned i = 0; i < t.theta.n_elem; ++i) theta[i] = t.theta(i);
  for (unsigned i = 0; i < t.z_shift.n_elem; ++i) z[i] = t.z_shift(i);
  for (unsigned i = 0; i < t.axis_shift_1.n_elem; ++i) a1[i] = t.axis_shift_1(i);
  for (unsigned i = 0; i < t.axis_shift_2.n_elem; ++i) a2[i] = t.axis_shift_2(i);
  size_t k = 0;
  for (unsigned i = 0; i < t.theta.n_elem; ++i)
    for (unsigned a = 0; a < t.z_shift.n_elem; ++a)
      for (unsigned b = 0; b < t.axis_shift_1.n_elem; ++b)
        for (unsigned c = 0; c < t.axis_shift_2.n_elem; ++c) rates[k++] = t.rate(i)(a, b, c);
}

double t0_get_rate(double theta, double z, double a1, double a2) { return g_sim->_scat_table.get_rate(theta, z, a1, a2); }

void t0_domain(double* lo_hi /*[6]: lo xyz, hi xyz*/) {
  for (int c = 0; c < 3; ++c) {
    lo_hi[c] = g_sim->_domain.first(c);
    lo_hi[3 + c] = g_sim->_domain.second(c);
  }
}
void t0_removal_domain(double* lo_hi) {
  for (int c = 0; c < 3; ++c) {
    lo_hi[c] = g_sim->_removal_domain.first(c);
    lo_hi[3 + c] = g_sim->_removal_domain.second(c);
  }
}
int64_t t0_num_inject() { return (int64_t)g_sim->_inject_scats.size(); }
void    t0_inject(int32_t* ids) {
  for (size_t i = 0; i < g_sim->_inject_scats.size(); ++i) ids[i] = (int32_t)(g_sim->_inject_scats[i] - site0());
}
double t0_time() { return g_sim->time(); }
double t0_max_time() { return g_sim->kubo_max_time(); }
double t0_cutoff() { return g_sim->_max_hopping_radius; }

// scatterer::find_neighbors (scatterer.cpp:34-83) for one site: neighbour ids and cumulative rates.
int64_t t0_row(int64_t i, int32_t* ids, double* cum, int64_t cap) {
  auto          nl = g_sim->_all_scat_list[i].find_neighbors(g_sim->_max_hopping_radius);
  const int64_t d = (int64_t)nl.size();
  for (int64_t k = 0; k < d && k < cap; ++k) {
    ids[k] = (int32_t)(nl[k].second - site0());
    cum[k] = nl[k].first;
  }
  return d;
}

// degree of every site (scatterer::no_of_neighbors, scatterer.cpp:85-104)
void t0_degrees(int32_t* deg) {
  const int64_t N = (int64_t)g_sim->_all_scat_list.size();
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < N; ++i) deg[i] = g_sim->_all_scat_list[i].no_of_neighbors(g_sim->_max_hopping_radius);
}

// whole table in CSR form, rows computed by the reference's find_neighbors
void t0_csr(const int64_t* row_ptr, int32_t* ids, double* cum) {
  const int64_t N = (int64_t)g_sim->_all_scat_list.size();
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < N; ++i) {
    auto nl = g_sim->_all_scat_list[i].find_neighbors(g_sim->_max_hopping_radius);
    for (size_t k = 0; k < nl.size(); ++k) {
      ids[row_ptr[i] + k] = (int32_t)(nl[k].second - site0());
      cum[row_ptr[i] + k] = nl[k].first;
    }
  }
}

// ---- excitons ----------------------------------------------------------------------------------------------------
void t0_log_draws(int on) { g_log_on = on != 0; }
void t0_clear_draws() {
  for (auto& v : g_draws) v.clear();
}

// monte_carlo::kubo_create_particles, verbatim (monte_carlo.cpp:308-316)
void t0_kubo_create_particles_verbatim() {
  g_cur = -1;
  g_sim->kubo_create_particles();
}

// same loop written out so that each draw is attributed to its exciton
void t0_kubo_create_particles_logged(int64_t n_particle) {
  for (int64_t i = 0; i < n_particle; ++i) {
    attribute_to(i);
    int                  dice = std::rand() % g_sim->_inject_scats.size();
    const mc::scatterer* s = g_sim->_inject_scats[dice];
    arma::vec            pos = s->pos();
    g_sim->_particle_list.push_back(mc::particle(pos, s, g_sim->_particle_velocity));
  }
  g_cur = -1;
}

int64_t t0_num_particles() { return (int64_t)g_sim->_particle_list.size(); }

void t0_particles(int32_t* site, double* pos /*[3][P]*/, double* old_pos, double* delta /*[3][P]*/, double* ff, int32_t* heading) {
  const auto&   pl = g_sim->_particle_list;
  const int64_t P = (int64_t)pl.size();
  for (int64_t i = 0; i < P; ++i) {
    site[i] = (int32_t)(pl[i].scat_ptr() - site0());
    for (int c = 0; c < 3; ++c) {
      pos[c * P + i] = pl[i].pos(c);
      if (old_pos) old_pos[c * P + i] = pl[i].old_pos(c);
      delta[c * P + i] = pl[i].delta_pos(c);
    }
    ff[i] = pl[i].ff_time();
    heading[i] = pl[i]._heading_right ? 1 : 0;
  }
}

// the ensemble averages of kubo_save_avg_dispalcement_squared (monte_carlo.cpp:396-406), same arithmetic
static void msd_row(double* out3) {
  double ax = 0, ay = 0, az = 0;
  for (const auto& p : g_sim->_particle_list) {
    ax += std::pow(p.delta_pos(0), 2);
    ay += std::pow(p.delta_pos(1), 2);
    az += std::pow(p.delta_pos(2), 2);
  }
  const double n = double(g_sim->_particle_list.size());
  out3[0] = ax / n;
  out3[1] = ay / n;
  out3[2] = az / n;
}

// nsteps x { kubo_step(dt) ; MSD row }  using the reference's kubo_step itself
void t0_kubo_step_verbatim(double dt, int64_t nsteps, double* msd /*[nsteps][3]*/, int write_file) {
  cout_silencer quiet;
  g_cur = -1;
  for (int64_t s = 0; s < nsteps; ++s) {
    g_sim->kubo_step(dt);
    if (write_file) g_sim->kubo_save_avg_dispalcement_squared();
    if (msd) msd_row(msd + 3 * s);
  }
}

// the same loop with per-exciton draw attribution (single thread, exciton index order = OMP_NUM_THREADS=1 order)
void t0_kubo_step_logged(double dt, int64_t nsteps, double* msd /*[nsteps][3]*/) {
  auto& sim = *g_sim;
  for (int64_t s = 0; s < nsteps; ++s) {
    for (unsigned i = 0; i < sim._particle_list.size(); ++i) {
      attribute_to(i);
      mc::particle& p = sim._particle_list[i];
      p.step(dt, sim._all_scat_list, sim._max_hopping_radius);
      p.update_delta_pos();
      if (arma::any(p.pos() < sim._removal_domain.first) || arma::any(sim._removal_domain.second < p.pos())) {
        int                  dice = std::rand() % sim._inject_scats.size();
        const mc::scatterer* sc = sim._inject_scats[dice];
        arma::vec            pos = sc->pos();
        p.set_pos(pos);
        p.set_scatterer(sc);
      }
    }
    sim._time += dt;
    if (msd) msd_row(msd + 3 * s);
  }
  g_cur = -1;
}

// CPU-baseline timing loop: monte_carlo::kubo_step (monte_carlo.cpp:319-342) with its OpenMP parallel-for, plus a count
// of re-injections so that hops = (draws - reinjections) / 2 is exact.  Draw logging must be off (not thread-safe).
int64_t t0_kubo_step_omp(double dt, int64_t nsteps) {
  auto&   sim = *g_sim;
  int64_t reinj = 0;
  g_cur = -1;
  for (int64_t s = 0; s < nsteps; ++s) {
#pragma omp parallel for reduction(+ : reinj)
    for (unsigned i = 0; i < sim._particle_list.size(); ++i) {
      mc::particle& p = sim._particle_list[i];
      p.step(dt, sim._all_scat_list, sim._max_hopping_radius);
      p.update_delta_pos();
      if (arma::any(p.pos() < sim._removal_domain.first) || arma::any(sim._removal_domain.second < p.pos())) {
        int                  dice = std::rand() % sim._inject_scats.size();
        const mc::scatterer* sc = sim._inject_scats[dice];
        arma::vec            pos = sc->pos();
        p.set_pos(pos);
        p.set_scatterer(sc);
        ++reinj;
      }
    }
    sim._time += dt;
  }
  return reinj;
}

int64_t t0_draw_count(int64_t i) { return (size_t)i < g_draws.size() ? (int64_t)g_draws[i].size() : 0; }
void    t0_draw_counts(int64_t* counts, int64_t P) {
  for (int64_t i = 0; i < P; ++i) counts[i] = t0_draw_count(i);
}
void t0_draws(int32_t* flat, int64_t P) {
  int64_t k = 0;
  for (int64_t i = 0; i < P && (size_t)i < g_draws.size(); ++i)
    for (int32_t r : g_draws[i]) flat[k++] = r;
}

// ---- contact mode (monte_carlo.h:343-355, 443-491, 525-643) ---------------------------------------------------------
int64_t t0_num_contact_sites(int which) { return (int64_t)(which == 1 ? g_sim->_c1_scat.size() : g_sim->_c2_scat.size()); }
void    t0_contact_sites(int which, int32_t* ids) {
  const auto& l = which == 1 ? g_sim->_c1_scat : g_sim->_c2_scat;
  for (size_t i = 0; i < l.size(); ++i) ids[i] = (int32_t)(l[i] - site0());
}
void t0_area(double* area) {
  for (size_t i = 0; i < g_sim->_area.size(); ++i) area[i] = g_sim->_area[i];
}
void t0_set_contact_pops(unsigned c1, unsigned c2) {
  g_sim->_c1_pop = c1;
  g_sim->_c2_pop = c2;
}
// one iteration of the (unreachable) contact loop in main.cpp:98-106
void t0_contact_iteration(double dt) {
  cout_silencer quiet;
  g_cur = -1;
  g_sim->step(dt);
  g_sim->save_metrics(dt);
  g_sim->repopulate_contacts();
}
// The same iteration with every draw attributed to the exciton that consumed it.  monte_carlo::step (monte_carlo.h:343-355)
// in list order (= OMP_NUM_THREADS=1), save_metrics verbatim, then repopulate_contacts (:443-455) with repopulate's list
// surgery (:458-491) written out so that g_ids follows every swap; a particle born here gets the next id.
static void repopulate_logged(const double ymin, const double ymax, const unsigned n_particle, const std::vector<const mc::scatterer*>& s_list) {
  auto&    p_list = g_sim->_particle_list;
  unsigned j = p_list.size();
  for (unsigned i = 0; i < j;) {
    if (p_list[i].pos(1) >= ymin && p_list[i].pos(1) <= ymax) {
      --j;
      std::swap(p_list[i], p_list[j]);
      std::swap(g_ids[i], g_ids[j]);
    } else {
      ++i;
    }
  }
  unsigned       n = 0;
  const unsigned final_size = j + n_particle;
  const unsigned j_lim = std::min(int(p_list.size()), int(final_size));
  for (; j < j_lim; ++j) {
    g_ids[j] = g_next_id++;
    attribute_to(g_ids[j]);
    const int dice = std::rand() % s_list.size();
    p_list[j] = mc::particle(s_list[dice]->pos(), s_list[dice], g_sim->_particle_velocity);
    ++n;
  }
  for (; n < n_particle; ++n) {
    g_ids.push_back(g_next_id++);
    attribute_to(g_ids.back());
    const int dice = std::rand() % s_list.size();
    p_list.emplace_back(mc::particle(s_list[dice]->pos(), s_list[dice], g_sim->_particle_velocity));
  }
  p_list.resize(final_size);
  g_ids.resize(final_size);
  g_cur = -1;
}
void t0_contact_iteration_logged(double dt) {
  cout_silencer quiet;
  auto&         sim = *g_sim;
  for (unsigned i = 0; i < sim._particle_list.size(); ++i) {
    attribute_to(g_ids[i]);
    sim._particle_list[i].step(dt, sim._all_scat_list, sim._max_hopping_radius);
  }
  g_cur = -1;
  sim._time += dt;
  sim.save_metrics(dt);
  const double ymin = sim._domain.first(1), ymax = sim._domain.second(1), dy = (ymax - ymin) / double(sim._n_seg);
  repopulate_logged(ymin, ymin + dy, sim._c1_pop, sim._c1_scat);
  repopulate_logged(ymin + double(sim._n_seg - 1) * dy, ymax, sim._c2_pop, sim._c2_scat);
}
int64_t t0_next_id() { return g_next_id; }
void    t0_particle_ids(int64_t* ids) {
  for (size_t i = 0; i < g_ids.size(); ++i) ids[i] = g_ids[i];
}
// monte_carlo::track_particle (monte_carlo.h:786-818), verbatim; every draw is attributed to exciton `log_slot`
void t0_track_particle(double dt, int file_no, int64_t log_slot) {
  cout_silencer quiet;
  attribute_to(log_slot);
  g_sim->track_particle(dt, file_no);
  g_cur = -1;
}
}  // extern "C"
