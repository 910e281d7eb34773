// host_setup.h -- order-defining set-up of the simulation, done once on the host.
//
// These steps are cheap (O(N)) but they fix the enumeration order of every neighbour list and therefore the bits of
// the cumulative rates, so they follow the reference exactly.  Reference citations are into /root/reference/src.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "json_min.h"

namespace cntmc {

// The "exciton monte carlo" block of input.json
struct Params {
  std::string           mesh_dir, output_dir;
  bool                  keep_old_results = true;
  std::string           rate_type;                         // davoody | forster | wong   (monte_carlo.cpp:27-59)
  std::array<double, 3> zshift{}, ashift1{}, ashift2{}, theta_deg{};  // [start, stop, count]
  double                max_hopping_radius = 0;            // monte_carlo.cpp:256
  double                velocity = 0;                      // monte_carlo.cpp:259
  int                   n_seg = 0;                         // monte_carlo.h:164 (contact mode)
  std::array<double, 2> xlim{}, ylim{}, zlim{};            // monte_carlo.cpp:276-278
  double                time_step = 0;                     // main.cpp:62
  int                   n_sections = 0;                    // monte_carlo.cpp:300
  double                max_time = 0;                      // monte_carlo.cpp:304
  int64_t               n_particles = 0;                   // monte_carlo.cpp:309
  bool                  has_table_grids = false;
  // "cnts" of the input (main.cpp:52-54 copies it into the block when the rate type is davoody): {n, m, length in cnt unit
  // cells} per tube, in the key order nlohmann::json iterates them (sorted), "directory" and "comment" left out
  std::vector<std::array<int, 3>> tubes;
};

// `doc` is either a whole input.json (with an "exciton monte carlo" member, main.cpp:46-49) or that block itself.
const json::Value& mc_block(const json::Value& doc);
Params             parse_params(const json::Value& block);

std::vector<double> linspace(double start, double end, int64_t n);

struct HostTable {
  std::vector<double> theta, z, a1, a2, rates;  // rates in [theta][z][a1][a2] C order
  bool                empty() const { return rates.empty(); }
};
// the four axes of the table (monte_carlo.cpp:65-75 / 157-167), rates left empty
HostTable make_table_axes(const Params& p);
// monte_carlo::create_scattering_table for "forster" (gamma0 = 1e15) and "wong" (1e13)  (monte_carlo.cpp:24-61, 156-200)
HostTable make_rate_table(const Params& p);
// the tubes of a "cnts" object (monte_carlo.cpp:33-47, cnt.h:159-193)
std::vector<std::array<int, 3>> parse_tubes(const json::Value& cnts);
// scattering_struct::save (scattering_struct.h:56-94): dir/scat_table.{theta,z_shift,axis_shift_1,axis_shift_2,rates}.dat.
// Same files and layout (axes one value per line; rates.dat = "sizes:" header, the four sizes, a blank line, then the
// rates theta-major); values are written with 17 significant digits so that a saved table loads back bit for bit
// (the reference prints 6, Armadillo 4 for the axes).
void save_rate_table(const HostTable& t, const std::string& dir);
// the inverse (the reference only writes; visualization/monte_carlo_results.py:261-282 reads the files the same way)
HostTable load_rate_table(const std::string& dir);

struct Mesh {
  int64_t             n_tubes = 0, n_cols = 0;
  std::vector<double> pos[3], orient[3];  // tube-major, positions in nm
};
// the six single_cnt.{pos,orient}.{x,y,z}.dat files (monte_carlo.h:206-245); Armadillo arma_ascii or raw ASCII
Mesh load_mesh(const std::string& dir);

struct Sites {
  int64_t              N = 0;
  std::vector<double>  pos[3], orient[3];  // metres
  std::vector<int32_t> left, right;
};
Sites create_sites(const Mesh& m);  // monte_carlo.h:247-263
// monte_carlo::trim_scats (monte_carlo.h:722-782): same surviving order and links as the reference's swap loop
void trim_sites(Sites& s, const std::array<double, 2>& xlim, const std::array<double, 2>& ylim,
                const std::array<double, 2>& zlim);

struct Domain {
  double lo[3], hi[3];
};
Domain find_domain(const Sites& s);  // monte_carlo.h:328-340

struct Buckets {
  int                  n[3] = {0, 0, 0};
  std::vector<int64_t> start;  // [ncell+1]
  std::vector<int32_t> sites;  // [N], ascending site index inside each cell
};
Buckets build_buckets(const Sites& s, const Domain& d, double radius);  // monte_carlo.h:375-395

struct Injection {
  std::vector<int32_t> sites;  // monte_carlo.cpp:203-228
  double               rem_lo[3], rem_hi[3];  // monte_carlo.cpp:231-251
};
Injection injection_region(const Sites& s, const Domain& d, int n_sections);

// contact mode
std::vector<double>  slab_areas(const Sites& s, const Domain& d, int n_seg);              // monte_carlo.h:646-688
std::vector<int32_t> contact_sites(const Sites& s, const Domain& d, int n_seg, int i);  // monte_carlo.h:494-516
std::vector<int32_t> slab_sites_half_open(const Sites& s, const Domain& d, int n_seg, int i);  // monte_carlo.h:296-301
// sites per slab as counted by monte_carlo::get_scatterer_statistics (monte_carlo.h:702-705): int(|y - ymin| / dy) % n_seg
std::vector<int64_t> slab_site_counts(const Sites& s, const Domain& d, int n_seg);

}  // namespace cntmc

// ---- device-table records built on the host --------------------------------------------------------------------------
#include "hop_core.h"
namespace cntmc {
// links and the two segment flight times of every site (hop_core.h SiteRec) plus the position records; the rate fields
// (total, 1/total, CSR row) are filled by the neighbour-table kernel.
std::vector<SiteRec> make_site_records(const Sites& s, double velocity, std::vector<PosRec>& pos);
// Tables::seg with its padding: element 4 + s is the segment time between sites s and s+1 where right[s] == s+1 and
// left[s+1] == s (and both records hold the same time, bit for bit), NaN elsewhere; four NaNs on either side
std::vector<DirRec> make_direction_records(const std::vector<SiteRec>& rec, const std::vector<PosRec>& pos);
constexpr int kSegPad = 4;
std::vector<double> make_segment_times(const std::vector<SiteRec>& rec);
}  // namespace cntmc
