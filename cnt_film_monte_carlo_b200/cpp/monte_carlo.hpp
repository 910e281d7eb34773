// cpp/monte_carlo.hpp -- the reference's C++ class surface, re-implemented as a thin shim over the C ABI (include/cntmc.h).
//
// mc::monte_carlo here has the public methods main() of the reference calls on its simulation object
// (src/main.cpp:64-76 and :85-105; declarations at src/monte_carlo/monte_carlo.h:111-842), with the same names,
// argument meaning, exceptions (std::invalid_argument for bad input, like monte_carlo.cpp:59 / main.cpp:47), directory
// handling (helper/prepare_directory.hpp) and output files.  Everything numerical happens behind cntmc_* on the GPU.
// The JSON argument is the text of the "exciton monte carlo" block (the reference passes an nlohmann::json object, a
// vendored third-party type this repository does not copy; `j.dump()` on the caller's side bridges the two).
#pragma once
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <experimental/filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cntmc.h"
#include "../csrc/json_min.h"

namespace mc {

namespace fs = std::experimental::filesystem;

// helper/prepare_directory.hpp:8-65
inline fs::path prepare_directory(std::string path, const bool keep_old_files = true) {
  if (!path.empty() && path[0] == '~') path = std::string(getenv("HOME") ? getenv("HOME") : "") + path.substr(1);
  const fs::path dir(path);
  if (!fs::exists(dir)) {
    fs::create_directories(dir);
    if (!fs::is_directory(dir)) throw std::invalid_argument("The input value for output directory is not acceptable.");
    return dir;
  }
  if (!fs::is_directory(dir)) throw std::invalid_argument("The input value for output directory is not acceptable.");
  if (!fs::is_empty(dir)) {
    if (keep_old_files) {
      int count = 1;
      while (fs::exists(dir.string() + "." + std::to_string(count))) count++;
      fs::rename(dir, dir.string() + "." + std::to_string(count));
    } else {
      fs::remove_all(dir);
    }
    fs::create_directories(dir);
  }
  return dir;
}
// helper/prepare_directory.hpp:68-105
inline fs::path check_directory(std::string path, bool should_be_empty = false) {
  if (!path.empty() && path[0] == '~') path = std::string(getenv("HOME") ? getenv("HOME") : "") + path.substr(1);
  const fs::path dir(path);
  if (!fs::exists(dir)) throw std::invalid_argument("directory does NOT exists!!!");
  if (!fs::is_directory(dir)) throw std::invalid_argument("input path is NOT a directory!!!");
  if (fs::is_empty(dir) && !should_be_empty) throw std::invalid_argument("directory is empty!!!");
  return dir;
}

class monte_carlo {
 private:
  cntmc_t*             _h = nullptr;
  cntmc_multi_t*       _m = nullptr;  // more than one GPU: the same simulation sharded over the GPUs of the box (cntmc_multi_*)
  cntmc::json::Value   _json_prop;
  fs::path             _output_directory, _input_directory;
  std::fstream         _displacement_squard_file, _pop_file, _curr_file;
  std::fstream         _displacement_file_x, _displacement_file_y, _displacement_file_z;
  std::vector<double>  _last_msd = std::vector<double>(3, 0.0);
  std::vector<int64_t> _last_pop, _last_curr;
  std::vector<double>  _area;
  double               _domain[6] = {0, 0, 0, 0, 0, 0};
  unsigned             _n_seg = 0;
  uint64_t             _seed = 100;  // main.cpp:30 seeds glibc with 100; here the value keys the exciton streams

  void ok(int rc) const {
    if (rc == CNTMC_OK) return;
    const std::string msg = _m ? cntmc_multi_last_error(_m) : cntmc_last_error(_h);
    if (rc == CNTMC_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }

 public:
  monte_carlo() = delete;
  monte_carlo(const monte_carlo&) = delete;

  // monte_carlo.h:116-136
  // n_gpus > 1: the excitons are split over GPUs 0..n_gpus-1 (the reference's OpenMP team over the particle list,
  // monte_carlo.cpp:320-338, becomes one GPU per shard and one NCCL all-reduce of the ensemble sums per call)
  explicit monte_carlo(const std::string& json_text, int n_gpus = 1) {
    std::cout << "\n" << "ready properties from json file" << "\n";
    _json_prop = cntmc::json::parse(json_text);
    if (n_gpus > 1) {
      if (cntmc_multi_create(json_text.c_str(), n_gpus, nullptr, &_m) != CNTMC_OK) throw std::invalid_argument(cntmc_multi_last_error(nullptr));
      _h = cntmc_multi_handle(_m, 0);  // read-backs that are the same on every GPU (domain, sites, areas) use GPU 0's handle
    } else {
      const int rc = cntmc_create(json_text.c_str(), &_h);
      if (rc != CNTMC_OK) throw std::invalid_argument(cntmc_last_error(nullptr));
    }
    bool keep_old_data = true;
    if (const auto* v = _json_prop.find("keep old results")) keep_old_data = v->as_bool();
    _output_directory = prepare_directory(_json_prop.at("output directory").as_string(), keep_old_data);
    _input_directory = check_directory(_json_prop.at("mesh input directory").as_string(), false);
  }
  ~monte_carlo() {
    if (_m)
      cntmc_multi_destroy(_m);
    else
      cntmc_destroy(_h);
  }

  double          time() const { return _m ? cntmc_multi_time(_m) : cntmc_time(_h); }   // monte_carlo.h:139
  const fs::path& output_path() const { return _output_directory; }                     // monte_carlo.h:148
  const fs::path& input_path() const { return _input_directory; }                       // monte_carlo.h:151
  unsigned        number_of_particles() const { return (unsigned)(_m ? cntmc_multi_number_of_particles(_m) : cntmc_number_of_particles(_h)); }  // monte_carlo.h:154
  int64_t         hops() const { return _m ? cntmc_multi_hops(_m) : cntmc_hops(_h); }
  double          kubo_max_time() const { return cntmc_kubo_max_time(_h); }             // monte_carlo.h:833
  double          time_step() const { return cntmc_time_step(_h); }
  cntmc_t*        handle() { return _h; }
  void            set_seed(uint64_t seed) { _seed = seed; }

  // monte_carlo.h:319-324
  void save_json_properties() {
    std::ofstream json_file((_output_directory / "input.json").string(), std::ios::out);
    cntmc::json::dump(_json_prop, json_file, 4);
    json_file << std::endl;
  }

  // ---- Green-Kubo flavour ----------------------------------------------------------------------------------------------
  // monte_carlo.cpp:254-305
  void kubo_init() {
    if (_m) {
      ok(cntmc_multi_load_mesh(_m, _input_directory.string().c_str()));
      ok(cntmc_multi_kubo_init(_m));
    } else {
      ok(cntmc_load_mesh(_h, _input_directory.string().c_str()));
      ok(cntmc_kubo_init(_h));
    }
    // create_davoody_scatt_table leaves its table in the output directory (monte_carlo.cpp:150); the closed-form ones do not
    if (_json_prop.at("rate type").as_string() == "davoody") save_scat_table();
    ok(cntmc_get_domain(_h, _domain));
    int64_t n = 0;
    ok(cntmc_num_sites(_h, &n));
    std::ios::fmtflags f(std::cout.flags());
    std::cout << std::fixed << std::showpos << "\n"
              << "simulation domain AFTER trimming:\n"
              << "    x (" << _domain[0] * 1e9 << " , " << _domain[3] * 1e9 << ") [nm]\n"
              << "    y (" << _domain[1] * 1e9 << " , " << _domain[4] * 1e9 << ") [nm]\n"
              << "    z (" << _domain[2] * 1e9 << " , " << _domain[5] * 1e9 << ") [nm]\n"
              << std::endl;
    std::cout.flags(f);
    std::cout << "total number of scatterers: " << n << std::endl;
  }
  // monte_carlo.cpp:308-316
  void kubo_create_particles() { ok(_m ? cntmc_multi_kubo_create_particles(_m, 0, _seed) : cntmc_kubo_create_particles(_h, 0, _seed, 0)); }
  // monte_carlo.cpp:319-342
  void kubo_step(double dt) { ok(_m ? cntmc_multi_kubo_step(_m, dt, 1, _last_msd.data()) : cntmc_kubo_step(_h, dt, 1, _last_msd.data())); }
  // monte_carlo.cpp:382-409
  void kubo_save_avg_dispalcement_squared() { write_msd_rows(&_last_msd[0], 1, time(), 0.0); }
  // monte_carlo.cpp:345-380: particle_dispalcement.{x,y,z}.dat, one column per exciton, one row per call
  void kubo_save_individual_particle_dispalcements() {
    const size_t        P = number_of_particles();
    std::vector<double> delta(3 * P);
    ok(_m ? cntmc_multi_get_particles(_m, nullptr, nullptr, delta.data(), nullptr, nullptr, nullptr)
          : cntmc_get_particles(_h, nullptr, nullptr, delta.data(), nullptr, nullptr, nullptr));
    std::fstream* files[3] = {&_displacement_file_x, &_displacement_file_y, &_displacement_file_z};
    const char*   names[3] = {"particle_dispalcement.x.dat", "particle_dispalcement.y.dat", "particle_dispalcement.z.dat"};
    for (int c = 0; c < 3; ++c) {
      std::fstream& f = *files[c];
      if (!f.is_open()) {
        f.open((_output_directory / names[c]).string(), std::ios::out);
        f << std::showpos << std::scientific;
        f << "time";
        for (int i = 0; i < int(P); ++i) f << "," << i;
        f << std::endl;
      }
      f << time();
      for (size_t i = 0; i < P; ++i) f << "," << delta[c * P + i];
      f << std::endl;
    }
  }
  // nsteps x { kubo_step(dt); kubo_save_avg_dispalcement_squared(); } in one engine call; the rows are the same
  void kubo_run(double dt, int64_t nsteps) {
    std::vector<double> msd((size_t)nsteps * 3);
    const double        t0 = time();
    ok(_m ? cntmc_multi_kubo_step(_m, dt, nsteps, msd.data()) : cntmc_kubo_step(_h, dt, nsteps, msd.data()));
    write_msd_rows(msd.data(), nsteps, t0, dt);
    for (int c = 0; c < 3; ++c) _last_msd[c] = msd[(size_t)(nsteps - 1) * 3 + c];
  }

  // ---- contact flavour ---------------------------------------------------------------------------------------------------
  // monte_carlo.h:157-195 (contact populations 1100 and 0 are hard-coded there, :191-192)
  void init(int64_t c1_pop = 1100, int64_t c2_pop = 0) {
    if (_m) {
      ok(cntmc_multi_load_mesh(_m, _input_directory.string().c_str()));
      ok(cntmc_multi_init(_m, c1_pop, c2_pop, _seed));
    } else {
      ok(cntmc_load_mesh(_h, _input_directory.string().c_str()));
      ok(cntmc_init(_h, c1_pop, c2_pop, _seed, 0));
    }
    if (_json_prop.at("rate type").as_string() == "davoody") save_scat_table();  // monte_carlo.cpp:150
    _n_seg = (unsigned)cntmc_number_of_segments(_h);
    _area.resize(_n_seg);
    ok(cntmc_get_area(_h, _area.data()));
    ok(cntmc_get_domain(_h, _domain));
    _last_pop.assign(_n_seg, 0);
    _last_curr.assign(_n_seg - 1, 0);
    std::cout << "number of segments: " << _n_seg << std::endl;
    get_scatterer_statistics();
  }
  // monte_carlo.h:691-719 (called from init(), :183)
  void get_scatterer_statistics() {
    std::vector<int64_t> pop(_n_seg, 0);
    ok(cntmc_get_scatterer_statistics(_h, pop.data()));
    int64_t n_sites = 0;
    ok(cntmc_num_sites(_h, &n_sites));
    const double ymin = _domain[1], ymax = _domain[4];
    const double dy = (ymax - ymin) / double(_n_seg);
    std::fstream f;
    f.open((_output_directory / "scatterer_statistics.dat").string(), std::ios::out);
    f << "position,distribution,population,density\n";
    for (unsigned i = 0; i < _n_seg; ++i)
      f << std::scientific << ymin + (double(i) + 0.5) * dy << "," << double(pop[i]) / double(n_sites) << "," << (long)pop[i] << ","
        << double(pop[i]) / (_area[i] * dy) << "\n";
    f.close();
  }
  // monte_carlo.h:786-818; the exciton's stream is (seed, 2^63 | fileNo) -- a domain of its own, so that a tracked exciton
  // never shares draws with exciton fileNo of the population; max_steps bounds the reference's unbounded loop
  bool track_particle(double dt, int fileNo, int64_t max_steps = int64_t(1) << 20) {
    std::vector<double> path((size_t)max_steps * 3);
    int64_t             n = 0;
    int32_t             reached = 0;
    ok(cntmc_track_particle(_h, dt, _seed, (1ull << 63) | (uint64_t)fileNo, 0, nullptr, nullptr, max_steps, path.data(), &n, &reached));
    std::ofstream file((_output_directory / ("particle_path." + std::to_string(fileNo) + ".dat")).string(), std::ios::out);
    file << std::scientific << std::showpos;
    for (int64_t s = 0; s < n; ++s) file << "   " << path[3 * s] << " " << path[3 * s + 1] << " " << path[3 * s + 2] << "\n";
    file << std::endl;
    return reached != 0;
  }
  // scattering_struct::save into the output directory (monte_carlo.cpp:150) and its inverse
  void save_scat_table() { ok(cntmc_save_rate_table(_h, _output_directory.string().c_str())); }
  void load_scat_table(const std::string& dir) { ok(cntmc_load_rate_table(_h, dir.c_str())); }
  // monte_carlo.h:343-355; the engine performs step, the counting of save_metrics and repopulate_contacts in one call
  void step(double dt) {
    ok(_m ? cntmc_multi_step(_m, dt, 1, _last_pop.data(), _last_curr.data()) : cntmc_step(_h, dt, 1, _last_pop.data(), _last_curr.data()));
  }
  // monte_carlo.h:519-522
  void save_metrics(double dt) {
    save_population_profile();
    save_currents(dt);
  }
  // monte_carlo.h:443-455: already applied on the device at the end of step(); kept so that main.cpp:98-106 compiles
  void repopulate_contacts() {}

 private:
  void write_msd_rows(const double* msd, int64_t n, double t0, double dt) {
    if (!_displacement_squard_file.is_open()) {
      _displacement_squard_file.open((_output_directory / "particle_dispalcement.avg.squared.dat").string(), std::ios::out);
      _displacement_squard_file << std::showpos << std::scientific;
      _displacement_squard_file << "# this file contains the average of dx^2, dy^2, and dz^2 of the particle ensemble over time" << std::endl
                                << "# number of particles: " << (size_t)number_of_particles() << std::endl
                                << std::endl;
      _displacement_squard_file << "time,x,y,z" << std::endl;
    }
    double t = t0;
    for (int64_t s = 0; s < n; ++s) {
      t += dt;  // monte_carlo.cpp:341
      _displacement_squard_file << t << "," << msd[s * 3] << "," << msd[s * 3 + 1] << "," << msd[s * 3 + 2] << std::endl;
    }
  }
  // monte_carlo.h:525-581
  void save_population_profile() {
    const double ymin = _domain[1], ymax = _domain[4];
    const double dy = (ymax - ymin) / double(_n_seg);
    if (!_pop_file.is_open()) {
      _pop_file.open((_output_directory / "population_profile.dat").string(), std::ios::out);
      _pop_file << "area";
      for (double a : _area) _pop_file << "," << std::scientific << std::showpos << a;
      _pop_file << std::endl << std::endl << "dy";
      for (unsigned i = 0; i < _n_seg; ++i) _pop_file << "," << dy;
      _pop_file << std::endl << std::endl << "section pos";
      for (unsigned i = 0; i < _n_seg; ++i) _pop_file << "," << ymin + (double(i) + 0.5) * dy;
      _pop_file << std::endl << std::endl << "time";
      for (unsigned i = 0; i < _n_seg; ++i) _pop_file << ",section" << i;
      _pop_file << std::endl;
    }
    _pop_file << std::showpos << std::scientific << time();
    for (unsigned j = 0; j < _n_seg; ++j) _pop_file << "," << double(_last_pop[j]) / (_area[j] * dy);
    _pop_file << std::endl;
  }
  // monte_carlo.h:584-643
  void save_currents(double dt) {
    const double        ymin = _domain[1], ymax = _domain[4];
    const int           n = (int)_n_seg;
    const double        dy = (ymax - ymin) / double(n);
    std::vector<double> area_at_interface;
    for (int i = 1; i < n; ++i) area_at_interface.push_back((_area[i - 1] + _area[i]) / 2);
    if (!_curr_file.is_open()) {
      _curr_file.open((_output_directory / "region_current.dat").string(), std::ios::out);
      _curr_file << "interface area";
      for (double a : area_at_interface) _curr_file << std::showpos << std::scientific << "," << a;
      _curr_file << std::endl << std::endl << "interface pos";
      for (int i = 1; i < n; ++i) _curr_file << "," << ymin + dy * double(i);
      _curr_file << std::endl << std::endl << "time";
      for (int i = 1; i < n; ++i) _curr_file << ",interface" << (i - 1);
      _curr_file << std::endl;
    }
    _curr_file << std::showpos << std::scientific << time();
    for (int i = 0; i + 1 < n; ++i) _curr_file << "," << double(_last_curr[i]) / (area_at_interface[i] * dt);
    _curr_file << std::endl;
  }
};

}  // namespace mc
