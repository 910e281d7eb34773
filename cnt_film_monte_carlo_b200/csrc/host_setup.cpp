// host_setup.cpp -- see host_setup.h.  Host C++ only; compiled with -ffp-contract=off so that no FMA is ever formed
// (the reference's x86-64 baseline build has none), which keeps the rate table and the domain arithmetic bit-identical.
#include "host_setup.h"

#include <limits>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cntmc {

namespace {
constexpr double kRefPi = 3.141592;  // helper/constants.h:10 -- truncated in the reference; it defines the theta grid

std::array<double, 3> triple(const json::Value& v, const char* key) {
  const json::Value& a = v.at(key);
  return {a.at(0).as_number(), a.at(1).as_number(), a.at(2).as_number()};
}
std::array<double, 2> couple(const json::Value& v, const char* key) {
  const json::Value& a = v.at(key);
  return {a.at(0).as_number(), a.at(1).as_number()};
}

// Armadillo's two-accumulator reductions (see hop_core.h)
double dot3(const double a[3], const double b[3]) {
  double v1 = 0.0, v2 = 0.0;
  v1 += a[0] * b[0];
  v2 += a[1] * b[1];
  v1 += a[2] * b[2];
  return v1 + v2;
}
double norm3(const double a[3]) { return std::sqrt(dot3(a, a)); }
void   normalise3(const double a[3], double out[3]) {
  const double n = norm3(a);
  const double d = (n > 0) ? n : 1.0;
  for (int c = 0; c < 3; ++c) out[c] = a[c] / d;
}
}  // namespace

const json::Value& mc_block(const json::Value& doc) {
  if (!doc.is_object()) throw std::invalid_argument("json input is not an object");
  if (const json::Value* b = doc.find("exciton monte carlo")) return *b;
  if (doc.contains("rate type") || doc.contains("mesh input directory")) return doc;
  // message of main.cpp:47
  throw std::invalid_argument("json input file does not contain \"exciton monte carlo\"");
}

Params parse_params(const json::Value& j) {
  Params p;
  if (const auto* v = j.find("mesh input directory")) p.mesh_dir = v->as_string();
  if (const auto* v = j.find("output directory")) p.output_dir = v->as_string();
  if (const auto* v = j.find("keep old results")) p.keep_old_results = v->as_bool();
  if (!j.contains("rate type")) throw std::invalid_argument("json: missing key \"rate type\"");  // assert at monte_carlo.cpp:25
  p.rate_type = j.at("rate type").as_string();
  if (j.contains("zshift [m]") && j.contains("axis shift 1 [m]") && j.contains("axis shift 2 [m]") &&
      j.contains("theta [degrees]")) {
    p.zshift = triple(j, "zshift [m]");
    p.ashift1 = triple(j, "axis shift 1 [m]");
    p.ashift2 = triple(j, "axis shift 2 [m]");
    p.theta_deg = triple(j, "theta [degrees]");
    p.has_table_grids = true;
  }
  p.max_hopping_radius = j.at("max hopping radius [m]").as_number();
  p.velocity = j.at("exciton velocity [m/s]").as_number();
  if (const auto* v = j.find("number of segments")) p.n_seg = (int)v->as_number();
  const json::Value& t = j.at("trim limits");
  p.xlim = couple(t, "xlim");
  p.ylim = couple(t, "ylim");
  p.zlim = couple(t, "zlim");
  if (const auto* v = j.find("monte carlo time step")) p.time_step = v->as_number();
  if (const auto* v = j.find("number of sections for injection region")) p.n_sections = (int)v->as_number();
  if (const auto* v = j.find("maximum time for kubo simulation [seconds]")) p.max_time = v->as_number();
  if (const auto* v = j.find("number of particles for kubo simulation")) p.n_particles = (int64_t)v->as_number();
  return p;
}

// arma::linspace: x[i] = start + i*delta, last element set to `end` exactly
std::vector<double> linspace(double start, double end, int64_t n) {
  std::vector<double> x((size_t)std::max<int64_t>(n, 1));
  if (n >= 2) {
    const double delta = (end - start) / double(n - 1);
    for (int64_t i = 0; i < n - 1; ++i) x[(size_t)i] = start + double(i) * delta;
    x[(size_t)n - 1] = end;
  } else {
    x[0] = end;
  }
  return x;
}

std::vector<std::array<int, 3>> parse_tubes(const json::Value& cnts) {
  if (!cnts.is_object()) throw std::invalid_argument("json: \"cnts\" is not an object");
  std::vector<std::pair<std::string, std::array<int, 3>>> named;
  for (const auto& kv : cnts.obj) {
    if (kv.first == "directory" || kv.first == "comment") continue;  // erased at monte_carlo.cpp:35-37
    const json::Value& chirality = kv.second.at("chirality");
    const json::Value& length = kv.second.at("length");
    if (length.at(1).as_string() != "cnt unit cells")  // message of cnt.h:187
      throw std::invalid_argument("units other than \"cnt unit cells\" is not implemented yet!!!");
    named.push_back({kv.first, {(int)chirality.at(0).as_number(), (int)chirality.at(1).as_number(), (int)length.at(0).as_number()}});
  }
  std::sort(named.begin(), named.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  std::vector<std::array<int, 3>> out;
  for (const auto& kv : named) out.push_back(kv.second);
  return out;
}

HostTable make_table_axes(const Params& p) {
  if (!p.has_table_grids) throw std::invalid_argument("json: the four rate-table grids are required");
  HostTable t;
  t.z = linspace(p.zshift[0], p.zshift[1], (int64_t)p.zshift[2]);
  t.a1 = linspace(p.ashift1[0], p.ashift1[1], (int64_t)p.ashift1[2]);
  t.a2 = linspace(p.ashift2[0], p.ashift2[1], (int64_t)p.ashift2[2]);
  t.theta = linspace(p.theta_deg[0], p.theta_deg[1], (int64_t)p.theta_deg[2]);
  for (double& th : t.theta) th *= (kRefPi / 180);
  return t;
}

HostTable make_rate_table(const Params& p) {
  double gamma0;
  if (p.rate_type == "forster") {
    gamma0 = 1.e15;
  } else if (p.rate_type == "wong") {
    gamma0 = 1.e13;
  } else if (p.rate_type == "davoody") {
    throw std::invalid_argument(
        "rate type \"davoody\" needs the input's \"cnts\" (the table is then built at initialisation) or a table "
        "installed with cntmc_set_rate_table / cntmc_load_rate_table / cntmc_create_davoody_table");
  } else {
    // message of monte_carlo.cpp:59
    throw std::invalid_argument("rate type must be one of the following: \"davoody\", \"forster\", \"wong\"");
  }
  HostTable t = make_table_axes(p);
  t.rates.reserve(t.theta.size() * t.z.size() * t.a1.size() * t.a2.size());
  // monte_carlo.cpp:172-193
  for (double th : t.theta)
    for (double zsh : t.z)
      for (double ash1 : t.a1)
        for (double ash2 : t.a2) {
          const double r1[3] = {ash1, 0, 0};
          const double r2[3] = {ash2 * std::cos(th), ash2 * std::sin(th), zsh};
          const double dR[3] = {r1[0] - r2[0], r1[1] - r2[1], r1[2] - r2[2]};
          double       u1[3], u2[3], ud[3];
          normalise3(r1, u1);
          normalise3(r2, u2);
          normalise3(dR, ud);
          const double angle_factor = std::cos(th) - 3 * dot3(u1, ud) * dot3(u2, ud);
          t.rates.push_back(gamma0 * (angle_factor * angle_factor) * std::pow(1.e-9 / norm3(dR), 6));
        }
  return t;
}

namespace {
// arma::Mat<double>::load(std::istream&) with auto-detection: ARMA_MAT_TXT_FN008 header, else raw ASCII rows
void load_matrix(const std::string& path, int64_t& rows, int64_t& cols, std::vector<double>& out) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("cannot open mesh file " + path);
  std::stringstream buf;
  buf << f.rdbuf();
  const std::string text = buf.str();
  const char*       c = text.c_str();
  const char*       end = c + text.size();
  out.clear();
  if (text.compare(0, 8, "ARMA_MAT") == 0) {
    while (c < end && *c != '\n') ++c;
    char* e = nullptr;
    rows = std::strtoll(c, &e, 10);
    c = e;
    cols = std::strtoll(c, &e, 10);
    c = e;
    out.reserve((size_t)(rows * cols));
    for (int64_t k = 0; k < rows * cols; ++k) {
      const double v = std::strtod(c, &e);
      if (e == c) throw std::invalid_argument("truncated mesh file " + path);
      out.push_back(v);
      c = e;
    }
    return;
  }
  rows = 0;
  cols = 0;
  while (c < end) {
    const char* line_end = c;
    while (line_end < end && *line_end != '\n') ++line_end;
    int64_t n_in_line = 0;
    while (c < line_end) {
      char*        e = nullptr;
      const double v = std::strtod(c, &e);
      if (e == c || e > line_end) break;
      out.push_back(v);
      ++n_in_line;
      c = e;
    }
    if (n_in_line > 0) {
      if (cols == 0) cols = n_in_line;
      if (n_in_line != cols) throw std::invalid_argument("ragged mesh file " + path);
      ++rows;
    }
    c = line_end + 1;
  }
}
}  // namespace

// ---- scat_table.*.dat -------------------------------------------------------------------------------------------------
namespace {
void write_column(const std::string& path, const std::vector<double>& v) {
  std::ofstream f(path);
  if (!f) throw std::invalid_argument("cannot write " + path);
  char buf[40];
  for (double x : v) {
    snprintf(buf, sizeof buf, "%.17g\n", x);
    f << buf;
  }
}
std::vector<double> read_numbers(std::istream& f) {
  std::vector<double> v;
  std::string         tok;
  while (f >> tok) {
    char*        end = nullptr;
    const double x = strtod(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') throw std::invalid_argument("not a number: \"" + tok + "\"");
    v.push_back(x);
  }
  return v;
}
std::vector<double> read_column(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("cannot open " + path);
  return read_numbers(f);
}
}  // namespace

void save_rate_table(const HostTable& t, const std::string& dir) {
  const std::string base = dir + "/scat_table";
  write_column(base + ".theta.dat", t.theta);
  write_column(base + ".z_shift.dat", t.z);
  write_column(base + ".axis_shift_1.dat", t.a1);
  write_column(base + ".axis_shift_2.dat", t.a2);
  std::ofstream f(base + ".rates.dat");
  if (!f) throw std::invalid_argument("cannot write " + base + ".rates.dat");
  f << "sizes:\n"
    << "theta, z_shift, axis_shift_1, axis_shift_2\n"
    << t.theta.size() << "," << t.z.size() << "," << t.a1.size() << "," << t.a2.size() << "\n\n";
  char buf[40];
  for (double x : t.rates) {
    snprintf(buf, sizeof buf, "%.17g\n", x);
    f << buf;
  }
}

HostTable load_rate_table(const std::string& dir) {
  const std::string base = dir + "/scat_table";
  HostTable         t;
  t.theta = read_column(base + ".theta.dat");
  t.z = read_column(base + ".z_shift.dat");
  t.a1 = read_column(base + ".axis_shift_1.dat");
  t.a2 = read_column(base + ".axis_shift_2.dat");
  std::ifstream f(base + ".rates.dat");
  if (!f) throw std::invalid_argument("cannot open " + base + ".rates.dat");
  std::string line;
  std::getline(f, line);  // "sizes:"
  std::getline(f, line);  // axis names
  std::getline(f, line);  // n_theta,n_z,n_a1,n_a2
  size_t dims[4] = {0, 0, 0, 0};
  {
    std::stringstream ss(line);
    std::string       tok;
    int               k = 0;
    while (k < 4 && std::getline(ss, tok, ',')) dims[k++] = (size_t)strtoull(tok.c_str(), nullptr, 10);
    if (k != 4) throw std::invalid_argument("scat_table.rates.dat: line 3 must hold the four table sizes");
  }
  if (dims[0] != t.theta.size() || dims[1] != t.z.size() || dims[2] != t.a1.size() || dims[3] != t.a2.size())
    throw std::invalid_argument("scat_table.rates.dat: sizes do not match the axis files");
  t.rates = read_numbers(f);
  if (t.rates.size() != dims[0] * dims[1] * dims[2] * dims[3] || t.rates.empty())
    throw std::invalid_argument("scat_table.rates.dat: wrong number of rates");
  return t;
}

Mesh load_mesh(const std::string& dir) {
  Mesh        m;
  const char* ax[3] = {"x", "y", "z"};
  for (int c = 0; c < 3; ++c) {
    int64_t r = 0, k = 0;
    load_matrix(dir + "/single_cnt.pos." + ax[c] + ".dat", r, k, m.pos[c]);
    if (c == 0) {
      m.n_tubes = r;
      m.n_cols = k;
    } else if (r != m.n_tubes || k != m.n_cols) {
      throw std::invalid_argument("mesh files have different shapes");
    }
    load_matrix(dir + "/single_cnt.orient." + ax[c] + ".dat", r, k, m.orient[c]);
    if (r != m.n_tubes || k != m.n_cols) throw std::invalid_argument("mesh files have different shapes");
  }
  if (m.n_tubes * m.n_cols == 0) throw std::invalid_argument("mesh is empty");
  return m;
}

// site n = tube * n_cols + col; positions nm -> m (xcoor *= 1.e-9); links to n-1 / n+1 inside a tube
Sites create_sites(const Mesh& m) {
  Sites s;
  s.N = m.n_tubes * m.n_cols;
  for (int c = 0; c < 3; ++c) {
    s.pos[c].resize((size_t)s.N);
    s.orient[c] = m.orient[c];
    for (int64_t n = 0; n < s.N; ++n) s.pos[c][(size_t)n] = m.pos[c][(size_t)n] * 1.e-9;
  }
  s.left.resize((size_t)s.N);
  s.right.resize((size_t)s.N);
  for (int64_t n = 0; n < s.N; ++n) {
    const int64_t col = n % m.n_cols;
    s.left[(size_t)n] = (col > 0) ? (int32_t)(n - 1) : -1;
    s.right[(size_t)n] = (col + 1 < m.n_cols) ? (int32_t)(n + 1) : -1;
  }
  return s;
}

// The reference scans i upward and, on an out-of-box site, swaps it with the current tail element (--j) and looks at
// position i again (monte_carlo.h:752-772).  Net effect, computed here without the swaps: with K survivors, every
// "hole" (removed site at a position < K, ascending) is filled by the surviving site with the highest position >= K
// not used yet (descending); survivors below K stay put; links between survivors are kept (re-indexed), links to
// removed sites are cut.
void trim_sites(Sites& s, const std::array<double, 2>& xlim, const std::array<double, 2>& ylim,
                const std::array<double, 2>& zlim) {
  const int64_t     N = s.N;
  std::vector<char> keep((size_t)N);
  int64_t           K = 0;
  for (int64_t i = 0; i < N; ++i) {
    const double x = s.pos[0][(size_t)i], y = s.pos[1][(size_t)i], z = s.pos[2][(size_t)i];
    const bool   out = x < xlim[0] || y < ylim[0] || z < zlim[0] || x > xlim[1] || y > ylim[1] || z > zlim[1];
    keep[(size_t)i] = !out;
    K += !out;
  }
  if (K == N) return;
  std::vector<int32_t> new_index((size_t)N, -1), source((size_t)K);
  int64_t              filler = N;
  for (int64_t i = 0; i < K; ++i) {
    if (keep[(size_t)i]) {
      new_index[(size_t)i] = (int32_t)i;
      source[(size_t)i] = (int32_t)i;
    } else {
      do {
        --filler;
      } while (!keep[(size_t)filler]);
      new_index[(size_t)filler] = (int32_t)i;
      source[(size_t)i] = (int32_t)filler;
    }
  }
  Sites t;
  t.N = K;
  for (int c = 0; c < 3; ++c) {
    t.pos[c].resize((size_t)K);
    t.orient[c].resize((size_t)K);
  }
  t.left.resize((size_t)K);
  t.right.resize((size_t)K);
  for (int64_t i = 0; i < K; ++i) {
    const size_t src = (size_t)source[(size_t)i];
    for (int c = 0; c < 3; ++c) {
      t.pos[c][(size_t)i] = s.pos[c][src];
      t.orient[c][(size_t)i] = s.orient[c][src];
    }
    const int32_t l = s.left[src], r = s.right[src];
    t.left[(size_t)i] = (l > -1) ? new_index[(size_t)l] : -1;
    t.right[(size_t)i] = (r > -1) ? new_index[(size_t)r] : -1;
  }
  s = std::move(t);
}

Domain find_domain(const Sites& s) {
  if (s.N == 0) throw std::invalid_argument("no scatterer left inside the trim limits");
  Domain d;
  for (int c = 0; c < 3; ++c) {
    d.lo[c] = *std::min_element(s.pos[c].begin(), s.pos[c].end());
    d.hi[c] = *std::max_element(s.pos[c].begin(), s.pos[c].end());
  }
  return d;
}

Buckets build_buckets(const Sites& s, const Domain& d, double radius) {
  Buckets b;
  for (int c = 0; c < 3; ++c) b.n[c] = (int)(std::ceil((d.hi[c] - d.lo[c]) / radius) + 1);
  const int64_t        nc = (int64_t)b.n[0] * b.n[1] * b.n[2];
  std::vector<int64_t> cell((size_t)s.N);
  b.start.assign((size_t)nc + 1, 0);
  for (int64_t i = 0; i < s.N; ++i) {
    const int     ix = (int)((s.pos[0][(size_t)i] - d.lo[0]) / radius);
    const int     iy = (int)((s.pos[1][(size_t)i] - d.lo[1]) / radius);
    const int     iz = (int)((s.pos[2][(size_t)i] - d.lo[2]) / radius);
    const int64_t idx = (int64_t)ix + (int64_t)iy * b.n[0] + (int64_t)iz * b.n[0] * b.n[1];
    cell[(size_t)i] = idx;
    b.start[(size_t)idx + 1]++;
  }
  for (int64_t k = 0; k < nc; ++k) b.start[(size_t)k + 1] += b.start[(size_t)k];
  std::vector<int64_t> fill(b.start.begin(), b.start.end() - 1);
  b.sites.resize((size_t)s.N);
  for (int64_t i = 0; i < s.N; ++i) b.sites[(size_t)fill[(size_t)cell[(size_t)i]]++] = (int32_t)i;  // stable
  return b;
}

Injection injection_region(const Sites& s, const Domain& d, int n) {
  if (!(n > 1 && n % 2 == 1))  // asserts at monte_carlo.cpp:204 and :232
    throw std::invalid_argument("\"number of sections for injection region\" must be odd and > 1");
  std::vector<double> g[3];
  for (int c = 0; c < 3; ++c) {
    const double step = (d.hi[c] - d.lo[c]) / double(n);
    for (int i = 0; i <= n; ++i) g[c].push_back(double(i) * step + d.lo[c]);
  }
  Injection inj;
  for (int64_t i = 0; i < s.N; ++i) {
    bool inside = true;
    for (int c = 0; c < 3; ++c) {
      const double v = s.pos[c][(size_t)i];
      inside = inside && g[c][(size_t)(n / 2)] <= v && v <= g[c][(size_t)(n / 2 + 1)];
    }
    if (inside) inj.sites.push_back((int32_t)i);
  }
  for (int c = 0; c < 3; ++c) {
    inj.rem_lo[c] = g[c][1];
    inj.rem_hi[c] = g[c][(size_t)n - 1];
  }
  return inj;
}

std::vector<double> slab_areas(const Sites& s, const Domain& d, int n_seg) {
  const double        ymin = d.lo[1], ymax = d.hi[1];
  const double        dy = (ymax - ymin) / double(n_seg);
  std::vector<double> xmax((size_t)n_seg, d.lo[0]), xmin((size_t)n_seg, d.hi[0]), zmax((size_t)n_seg, d.lo[2]),
      zmin((size_t)n_seg, d.hi[2]);
  for (int64_t k = 0; k < s.N; ++k) {
    int i = (int)((s.pos[1][(size_t)k] - ymin) / dy);
    i = i < 0 ? 0 : (i < n_seg ? i : n_seg - 1);
    const double x = s.pos[0][(size_t)k], z = s.pos[2][(size_t)k];
    // the reference updates the minimum OR the maximum for a site, never both (monte_carlo.h:662-672)
    if (xmin[(size_t)i] > x) {
      xmin[(size_t)i] = x;
    } else if (xmax[(size_t)i] < x) {
      xmax[(size_t)i] = x;
    }
    if (zmin[(size_t)i] > z) {
      zmin[(size_t)i] = z;
    } else if (zmax[(size_t)i] < z) {
      zmax[(size_t)i] = z;
    }
  }
  std::vector<double> area((size_t)n_seg);
  for (int i = 0; i < n_seg; ++i) area[(size_t)i] = (zmax[(size_t)i] - zmin[(size_t)i]) * (xmax[(size_t)i] - xmin[(size_t)i]);
  return area;
}

std::vector<int64_t> slab_site_counts(const Sites& s, const Domain& d, int n_seg) {
  const double         ymin = d.lo[1];
  const double         dy = (d.hi[1] - ymin) / double(n_seg);
  std::vector<int64_t> pop((size_t)n_seg, 0);
  for (int64_t k = 0; k < s.N; ++k) pop[(size_t)(int(std::abs(s.pos[1][(size_t)k] - ymin) / dy) % n_seg)]++;
  return pop;
}

std::vector<int32_t> contact_sites(const Sites& s, const Domain& d, int n_seg, int i) {
  const double         dy = (d.hi[1] - d.lo[1]) / double(n_seg);
  const double         y1 = d.lo[1] + double(i - 1) * dy, y2 = d.lo[1] + double(i) * dy;
  std::vector<int32_t> out;
  for (int64_t k = 0; k < s.N; ++k)
    if (s.pos[1][(size_t)k] >= y1 && s.pos[1][(size_t)k] <= y2) out.push_back((int32_t)k);
  return out;
}

std::vector<int32_t> slab_sites_half_open(const Sites& s, const Domain& d, int n_seg, int i) {
  const double         dy = (d.hi[1] - d.lo[1]) / double(n_seg);
  const double         y1 = d.lo[1] + double(i) * dy, y2 = y1 + dy;
  std::vector<int32_t> out;
  for (int64_t k = 0; k < s.N; ++k)
    if (y1 <= s.pos[1][(size_t)k] && s.pos[1][(size_t)k] < y2) out.push_back((int32_t)k);
  return out;
}

}  // namespace cntmc

namespace cntmc {
std::vector<SiteRec> make_site_records(const Sites& s, double velocity, std::vector<PosRec>& pos) {
  std::vector<SiteRec> rec((size_t)s.N);
  pos.resize((size_t)s.N);
  for (int64_t i = 0; i < s.N; ++i) {
    SiteRec&     r = rec[(size_t)i];
    const double x = s.pos[0][(size_t)i], y = s.pos[1][(size_t)i], z = s.pos[2][(size_t)i];
    pos[(size_t)i] = PosRec{x, y, z, 0.0};
    r.left = s.left[(size_t)i];
    r.right = s.right[(size_t)i];
    // create_scatterers + trim_scats always produce symmetric links (monte_carlo.h:249-262, 730-772); anything else is a
    // corrupted site list.
    if ((r.left > -1 && s.right[(size_t)r.left] != (int32_t)i) || (r.right > -1 && s.left[(size_t)r.right] != (int32_t)i))
      throw std::invalid_argument("chain links of the site list are not symmetric");
    if (r.left > -1 && r.left == r.right) throw std::invalid_argument("a site is linked to the same neighbour on both sides");
    r.q_right = r.q_left = 0.0;
    if (r.left > -1)
      r.q_left = segment_time(x, y, z, s.pos[0][(size_t)r.left], s.pos[1][(size_t)r.left], s.pos[2][(size_t)r.left], velocity);
    if (r.right > -1)
      r.q_right = segment_time(x, y, z, s.pos[0][(size_t)r.right], s.pos[1][(size_t)r.right], s.pos[2][(size_t)r.right], velocity);
    r.inv_total = 0.0;
    r.row_begin = r.row_len = 0;
    for (int k = 0; k < 8; ++k) r.guide[k] = 0;
    for (int k = 0; k < 6; ++k) r.spare[k] = 0.0;
    r.top = TopRec{0, 0, 0, -1, -1, -1, 0.0};  // filled by the table build
  }
  return rec;
}

// normalise(next.pos - pos) with the arithmetic of move_along (hop_core.h): w / (|w| > 0 ? |w| : 1)
std::vector<DirRec> make_direction_records(const std::vector<SiteRec>& rec, const std::vector<PosRec>& pos) {
  std::vector<DirRec> dir(rec.size(), DirRec{0, 0, 0, 0, 0, 0, 0, 0});
  for (size_t i = 0; i < rec.size(); ++i) {
    for (int side = 0; side < 2; ++side) {
      const int32_t n = side == 0 ? rec[i].right : rec[i].left;
      if (n < 0) continue;
      const double wx = pos[(size_t)n].x - pos[i].x, wy = pos[(size_t)n].y - pos[i].y, wz = pos[(size_t)n].z - pos[i].z;
      const double nn = norm3(wx, wy, wz);
      const double den = (nn > 0) ? nn : 1.0;
      double*      u = side == 0 ? &dir[i].rx : &dir[i].lx;
      u[0] = wx / den;
      u[1] = wy / den;
      u[2] = wz / den;
    }
  }
  return dir;
}

std::vector<double> make_segment_times(const std::vector<SiteRec>& rec) {
  const size_t        N = rec.size();
  const double        nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<double> seg(N + 2 * kSegPad, nan);
  for (size_t s = 0; s + 1 < N; ++s) {
    const SiteRec &a = rec[s], &b = rec[s + 1];
    if (a.right == (int32_t)(s + 1) && b.left == (int32_t)s && memcmp(&a.q_right, &b.q_left, sizeof(double)) == 0) seg[kSegPad + s] = a.q_right;
  }
  return seg;
}
}  // namespace cntmc
