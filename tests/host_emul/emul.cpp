// tests/host_emul/emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the engine's host/device-shared arithmetic (csrc/hop_core.h, csrc/csr_core.h) and its host set-up
// (csrc/host_setup.cpp) with g++ and drives them sequentially, so the CPU test-suite can compare the exact code the
// CUDA kernels inline against the oracle on a machine without a GPU.  The product never builds or loads this file.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../cnt_film_monte_carlo_b200/csrc/csr_core.h"
#include "../../cnt_film_monte_carlo_b200/csrc/hop_core.h"
#include "../../cnt_film_monte_carlo_b200/csrc/host_setup.h"

using namespace cntmc;

struct Emul {
  Params               prm;
  HostTable            table;
  Sites                sites;
  Domain               dom;
  Buckets              buckets;
  Injection            inj;
  std::vector<SiteRec> site;
  std::vector<double>  seg;
  std::vector<DirRec>  dir;
  bool                 runs = true;
  std::vector<PosRec>  pos;
  std::vector<double>  cum;
  std::vector<int32_t> nbr;
  std::vector<RowEntry> row;
  std::vector<int64_t> row_ptr;
  Tables               T;
  std::vector<Lane>    lanes;
  uint64_t             seed = 0, first_gid = 0;
  std::vector<int64_t> r_off;
  std::vector<int32_t> r_draws;
  std::vector<double>  r_logs;
  bool                 replay = false;
  int64_t              guards = 0, hops = 0, fast = 0;
  bool                 use_top = false;  // decide events by the top entries of the site record where possible
  std::string          err;
  std::vector<std::vector<int32_t>> trace;
};

extern "C" {

Emul* emul_create(const char* json_text) {
  Emul* e = new Emul;
  try {
    const json::Value doc = json::parse(json_text);
    e->prm = parse_params(mc_block(doc));
  } catch (const std::exception& ex) {
    e->err = ex.what();
  }
  return e;
}
const char* emul_error(Emul* e) { return e->err.c_str(); }
void emul_destroy(Emul* e) { delete e; }

int emul_kubo_init(Emul* e, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient) {
  try {
    Mesh         m;
    const size_t N = (size_t)(n_tubes * n_cols);
    m.n_tubes = n_tubes;
    m.n_cols = n_cols;
    for (int c = 0; c < 3; ++c) {
      m.pos[c].assign(pos_nm + c * N, pos_nm + (c + 1) * N);
      m.orient[c].assign(orient + c * N, orient + (c + 1) * N);
    }
    e->table = make_rate_table(e->prm);
    e->sites = create_sites(m);
    trim_sites(e->sites, e->prm.xlim, e->prm.ylim, e->prm.zlim);
    e->dom = find_domain(e->sites);
    e->buckets = build_buckets(e->sites, e->dom, e->prm.max_hopping_radius);
    e->inj = injection_region(e->sites, e->dom, e->prm.n_sections);
    const int64_t n = e->sites.N;
    std::vector<SiteGeom> geom((size_t)n);
    e->site = make_site_records(e->sites, e->prm.velocity, e->pos);
    for (int64_t i = 0; i < n; ++i) {
      geom[i] = SiteGeom{e->sites.pos[0][i], e->sites.pos[1][i], e->sites.pos[2][i],
                         e->sites.orient[0][i], e->sites.orient[1][i], e->sites.orient[2][i]};
    }
    RateTable R{e->table.theta.data(), e->table.z.data(), e->table.a1.data(), e->table.a2.data(), e->table.rates.data(),
                (int32_t)e->table.theta.size(), (int32_t)e->table.z.size(), (int32_t)e->table.a1.size(), (int32_t)e->table.a2.size()};
    R.guard_tol = 1e-9;
    grid_hint(R.theta, R.n_theta, &R.start[0], &R.inv_step[0]);  // as cntmc_api.cu does: evenly spaced grids are not scanned in full
    grid_hint(R.z, R.n_z, &R.start[1], &R.inv_step[1]);
    grid_hint(R.a1, R.n_a1, &R.start[2], &R.inv_step[2]);
    grid_hint(R.a2, R.n_a2, &R.start[3], &R.inv_step[3]);
    const double radius = e->prm.max_hopping_radius;
    e->row_ptr.assign((size_t)n + 1, 0);
    e->cum.clear();
    e->nbr.clear();
    // same enumeration as csr_rows_kernel
    for (int64_t i = 0; i < n; ++i) {
      const SiteGeom& s1 = geom[i];
      const int cx = cell_coord(s1.px, e->dom.lo[0], radius), cy = cell_coord(s1.py, e->dom.lo[1], radius),
                cz = cell_coord(s1.pz, e->dom.lo[2], radius);
      uint32_t d = 0;
      double   acc = 0;
      bool     guard = false;
      TopEntries top;
      top.clear();
      for (int ix = cx - 1; ix <= cx + 1; ++ix)
        for (int iy = cy - 1; iy <= cy + 1; ++iy)
          for (int iz = cz - 1; iz <= cz + 1; ++iz) {
            if (!(ix > -1 && ix < e->buckets.n[0] && iy > -1 && iy < e->buckets.n[1] && iz > -1 && iz < e->buckets.n[2])) continue;
            const int64_t b = (int64_t)ix + (int64_t)iy * e->buckets.n[0] + (int64_t)iz * e->buckets.n[0] * e->buckets.n[1];
            for (int64_t q = e->buckets.start[b]; q < e->buckets.start[b + 1]; ++q) {
              const SiteGeom& s2 = geom[e->buckets.sites[q]];
              if (!within_cutoff(s1, s2, radius)) continue;
              const double rate = pair_rate(s1, s2, R, &guard);
              const double below = (d == 0) ? -1.0 : acc;
              acc = (d == 0) ? rate : acc + rate;
              top.add(rate, below, acc, e->buckets.sites[q]);
              e->nbr.push_back(e->buckets.sites[q]);
              e->cum.push_back(acc);
              ++d;
            }
          }
      e->row_ptr[i + 1] = e->row_ptr[i] + d;
      e->site[i].inv_total = d ? 1. / acc : 0.0;
      {
        uint8_t g[kGuideBuckets];
        build_guide(e->cum.data() + e->row_ptr[i], d, acc, g);
        memcpy(e->site[i].guide, g, sizeof(g));
      }
      e->site[i].row_begin = (uint32_t)e->row_ptr[i];
      e->site[i].row_len = d;
      top.store(e->site[i].top, acc, d);
      e->guards += guard;
    }
    e->seg = make_segment_times(e->site);
    e->T.seg = e->runs ? e->seg.data() + kSegPad : nullptr;
    e->dir = make_direction_records(e->site, e->pos);
    e->T.dir = e->runs ? e->dir.data() : nullptr;
    e->T.site = e->site.data();
    e->T.pos = e->pos.data();
    e->row.resize(e->cum.size());
    for (size_t k = 0; k < e->cum.size(); ++k) e->row[k] = RowEntry{e->cum[k], e->nbr[k], 0};
    e->T.row = e->row.data();
    e->T.inject = e->inj.sites.data();
    e->T.n_inject = (int32_t)e->inj.sites.size();
    for (int c = 0; c < 3; ++c) {
      e->T.rem_lo[c] = e->inj.rem_lo[c];
      e->T.rem_hi[c] = e->inj.rem_hi[c];
    }
    e->T.velocity = e->prm.velocity;
    e->T.inv_velocity = 1.0 / e->prm.velocity;
    return 0;
  } catch (const std::exception& ex) {
    e->err = ex.what();
    return -1;
  }
}

int64_t emul_num_sites(Emul* e) { return e->sites.N; }
int64_t emul_nnz(Emul* e) { return (int64_t)e->cum.size(); }
int64_t emul_guards(Emul* e) { return e->guards; }
void emul_sites(Emul* e, double* pos, double* orient, int32_t* left, int32_t* right, double* max_rate, double* inv) {
  const size_t N = (size_t)e->sites.N;
  for (int c = 0; c < 3; ++c) {
    memcpy(pos + c * N, e->sites.pos[c].data(), N * 8);
    memcpy(orient + c * N, e->sites.orient[c].data(), N * 8);
  }
  memcpy(left, e->sites.left.data(), N * 4);
  memcpy(right, e->sites.right.data(), N * 4);
  for (size_t i = 0; i < N; ++i) {
    max_rate[i] = e->site[i].top.total;
    inv[i] = 1. / e->site[i].top.total;
  }
}
void emul_csr(Emul* e, int64_t* row_ptr, int32_t* nbr, double* cum) {
  memcpy(row_ptr, e->row_ptr.data(), e->row_ptr.size() * 8);
  memcpy(nbr, e->nbr.data(), e->nbr.size() * 4);
  memcpy(cum, e->cum.data(), e->cum.size() * 8);
}
void emul_domains(Emul* e, double* dom, double* rem) {
  for (int c = 0; c < 3; ++c) {
    dom[c] = e->dom.lo[c]; dom[3 + c] = e->dom.hi[c];
    rem[c] = e->inj.rem_lo[c]; rem[3 + c] = e->inj.rem_hi[c];
  }
}
int64_t emul_num_inject(Emul* e) { return (int64_t)e->inj.sites.size(); }
void emul_inject(Emul* e, int32_t* ids) { memcpy(ids, e->inj.sites.data(), e->inj.sites.size() * 4); }
void emul_table(Emul* e, double* rates) { memcpy(rates, e->table.rates.data(), e->table.rates.size() * 8); }

}  // extern "C"

template <typename Draws>
static void init_draws(Emul* e, Draws& D, int64_t i);
template <>
void init_draws<PhiloxDraws>(Emul* e, PhiloxDraws& D, int64_t i) { D.init(e->seed, e->first_gid + (uint64_t)i); }
template <>
void init_draws<ReplayDraws>(Emul* e, ReplayDraws& D, int64_t i) {
  D.init(e->r_draws.data(), e->r_logs.empty() ? nullptr : e->r_logs.data(), e->r_off[i], e->r_off[i + 1]);
}

template <typename Draws>
static void create(Emul* e, int64_t P) {
  e->lanes.assign((size_t)P, Lane{});
  e->trace.assign((size_t)P, {});
  for (int64_t i = 0; i < P; ++i) {
    Draws D;
    init_draws(e, D, i);
    create_exciton(e->lanes[i], e->T, D, e->T.inject, e->T.n_inject);
  }
}
extern "C" {
void emul_create_philox(Emul* e, int64_t P, uint64_t seed, uint64_t first_gid) {
  e->replay = false;
  e->seed = seed;
  e->first_gid = first_gid;
  create<PhiloxDraws>(e, P);
}
void emul_create_replay(Emul* e, int64_t P, const int64_t* off, const int32_t* draws, const double* logs) {
  e->replay = true;
  e->r_off.assign(off, off + P + 1);
  e->r_draws.assign(draws, draws + off[P]);
  if (logs) e->r_logs.assign(logs, logs + off[P]); else e->r_logs.clear();
  create<ReplayDraws>(e, P);
}

}  // extern "C"

// the flat loop of kubo_flat_kernel, one lane at a time; msd summed in exciton order
template <typename Draws>
static int step(Emul* e, double dt, int64_t nsteps, double* msd, int trace_cap) {
  const int64_t       P = (int64_t)e->lanes.size();
  std::vector<double> sums((size_t)nsteps * 3, 0.0);
  int                 bad = 0;
  std::vector<int32_t> tr((size_t)(trace_cap > 0 ? trace_cap : 1));
  for (int64_t i = 0; i < P; ++i) {
    Lane&  L = e->lanes[i];
    Draws  D;
    Cursor c;
    init_draws(e, D, i);
    L.nevent = 0;
    L.nfast = 0;
    c.step = 0;
    begin_step(c, L, dt);
    while (c.step < nsteps && !L.stuck) {
      if (advance(L, e->T, D, c, trace_cap > 0 ? tr.data() : nullptr, (uint32_t)trace_cap, e->use_top)) {
        sums[c.step * 3 + 0] += L.dx * L.dx;
        sums[c.step * 3 + 1] += L.dy * L.dy;
        sums[c.step * 3 + 2] += L.dz * L.dz;
        ++c.step;
        begin_step(c, L, dt);
      }
    }
    e->hops += L.nevent;
    e->fast += L.nfast;
    if (trace_cap > 0) e->trace[i].insert(e->trace[i].end(), tr.begin(), tr.begin() + std::min<uint32_t>(L.nevent, trace_cap));
    bad |= L.stuck || D.exhausted();
  }
  if (msd)
    for (size_t k = 0; k < sums.size(); ++k) msd[k] = sums[k] / double(P);
  return bad;
}
extern "C" {
int emul_kubo_step(Emul* e, double dt, int64_t nsteps, double* msd, int trace_cap) {
  return e->replay ? step<ReplayDraws>(e, dt, nsteps, msd, trace_cap) : step<PhiloxDraws>(e, dt, nsteps, msd, trace_cap);
}
int64_t emul_hops(Emul* e) { return e->hops; }
int64_t emul_top_events(Emul* e) { return e->fast; }
void emul_set_top_entries(Emul* e, int on) { e->use_top = on != 0; }
void emul_set_runs(Emul* e, int on) {
  e->runs = on != 0;
  e->T.seg = e->runs && !e->seg.empty() ? e->seg.data() + kSegPad : nullptr;
  e->T.dir = e->runs && !e->dir.empty() ? e->dir.data() : nullptr;
}
void emul_particles(Emul* e, int32_t* site, double* pos, double* delta, double* ff, int32_t* heading, uint32_t* ndraw) {
  const size_t P = e->lanes.size();
  for (size_t i = 0; i < P; ++i) {
    const Lane& L = e->lanes[i];
    site[i] = L.site;
    pos[i] = L.px; pos[P + i] = L.py; pos[2 * P + i] = L.pz;
    delta[i] = L.dx; delta[P + i] = L.dy; delta[2 * P + i] = L.dz;
    ff[i] = L.ff;
    heading[i] = L.heading_right;
    ndraw[i] = L.ndraw;
  }
}
void emul_trace_counts(Emul* e, int64_t* counts) {
  for (size_t i = 0; i < e->trace.size(); ++i) counts[i] = (int64_t)e->trace[i].size();
}
void emul_trace(Emul* e, int32_t* flat) {
  size_t k = 0;
  for (auto& v : e->trace)
    for (int32_t s : v) flat[k++] = s;
}
int64_t emul_select(const double* cum, int64_t d, double dice) { return (int64_t)select_entry(cum, 0u, (uint32_t)d - 1u, dice); }
// the engine's own search over interleaved entries (nbr[k] = k so that the destination is the index)
static int64_t select_via_entries(const double* cum, int64_t d, uint32_t lo, uint32_t hi, double dice) {
  std::vector<RowEntry> row((size_t)d);
  for (int64_t k = 0; k < d; ++k) row[(size_t)k] = RowEntry{cum[k], (int32_t)k, 0};
  return select_dest(row.data(), lo, hi, dice);
}
// guided search exactly as after_flight_scatter performs it for draw r
int64_t emul_select_guided(const double* cum, int64_t d, int32_t r) {
  const double total = cum[d - 1];
  uint8_t      g8[kGuideBuckets];
  build_guide(cum, (uint32_t)d, total, g8);
  uint32_t g[2];
  memcpy(g, g8, sizeof(g));
  const double dice = total * (double)r / kRandMax;
  uint32_t     lo, hi;
  guide_bracket(g[0], g[1], (uint32_t)d, r, lo, hi);
  return select_via_entries(cum, d, lo, hi, dice);
}
// the decision of the top entries (after_flight_scatter) for draw r: index of the top entry whose interval of the draw holds r,
// or -1 (ordinary search needed); *dice receives the dice the ordinary search would use for r
int64_t emul_select_top(const double* cum, int64_t d, int32_t r, double* dice) {
  TopEntries top;
  top.clear();
  for (int64_t k = 0; k < d; ++k) top.add(k ? cum[k] - cum[k - 1] : cum[0], k ? cum[k - 1] : -1.0, cum[k], (int32_t)k);
  TopRec t{};
  top.store(t, cum[d - 1], (uint32_t)d);
  *dice = dice_of(cum[d - 1], (uint32_t)r);
  const uint32_t u = (uint32_t)r >> kTopBlockShift;
  if (in_draw_blocks(t.iv0, u)) return t.nbr0;
  if (in_draw_blocks(t.iv1, u)) return t.nbr1;
  if (in_draw_blocks(t.iv2, u)) return t.nbr2;
  return -1;
}
// the packed block interval of the draws [rlo, rhi) (hop_core.h draw_blocks) and the membership test the event applies to it
uint32_t emul_draw_blocks(uint32_t rlo, uint32_t rhi) { return draw_blocks(rlo, rhi); }
int emul_in_draw_blocks(uint32_t iv, uint32_t r) { return in_draw_blocks(iv, r >> kTopBlockShift) ? 1 : 0; }
// first_draw_reaching against its definition: the smallest draw whose dice is >= x (checked on both sides of the answer)
int64_t emul_first_draw_reaching(double total, double x) {
  const uint32_t r = first_draw_reaching(total, x);
  if (r <= 0x7fffffffu && !(dice_of(total, r) >= x)) return -1;
  if (r > 0u && dice_of(total, r - 1u) >= x) return -2;
  return (int64_t)r;
}
int64_t emul_select_full(const double* cum, int64_t d, double dice) { return select_via_entries(cum, d, 0u, (uint32_t)d - 1u, dice); }
// div_by against the division it replaces: number of operands (out of n) whose quotients differ
int64_t emul_div_by_mismatches(const double* x, int64_t n, double c) {
  const double rc = 1.0 / c;
  int64_t      bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double a = x[i] / c, b = div_by(x[i], c, rc);
    bad += memcmp(&a, &b, 8) != 0;
  }
  return bad;
}
void emul_philox2x32(uint32_t c0, uint32_t c1, uint32_t k, uint32_t* out) { philox2x32_10(c0, c1, k, out); }
}  // extern "C"

// ---- scheduling study support: the per-exciton sequence of micro-operations of one launch ----------------------------
// one int32 per advance() call: (chain crossings << 1) | (1 if the call ended a time step, 0 if it was an event)
extern "C" int64_t emul_op_sequences(Emul* e, double dt, int64_t nsteps, int64_t* counts, int32_t* flat, int64_t cap) {
  const int64_t P = (int64_t)e->lanes.size();
  int64_t       k = 0;
  for (int64_t i = 0; i < P; ++i) {
    Lane&       L = e->lanes[i];
    PhiloxDraws D;
    Cursor      c;
    D.init(e->seed, e->first_gid + (uint64_t)i);
    L.nevent = 0;
    c.step = 0;
    begin_step(c, L, dt);
    int64_t n = 0;
    while (c.step < nsteps && !L.stuck) {
      const uint32_t c0 = L.ncross;
      const bool     ended = advance(L, e->T, D, c, nullptr, 0);
      if (k < cap) flat[k] = (int32_t)(((L.ncross - c0) << 1) | (ended ? 1 : 0));
      ++k;
      ++n;
      if (ended) {
        ++c.step;
        begin_step(c, L, dt);
      }
    }
    counts[i] = n;
  }
  return k;
}

// number of arguments for which the windowed nearest-grid search disagrees with the full scan (0 expected); *hinted = whether
// the grid qualified for the windowed search at all
extern "C" int64_t emul_argmin_mismatches(const double* grid, int n, const double* xs, int64_t m, int* hinted) {
  double start = 0, inv_step = 0;
  grid_hint(grid, n, &start, &inv_step);
  *hinted = inv_step > 0 ? 1 : 0;
  int64_t bad = 0;
  for (int64_t k = 0; k < m; ++k)
    if (argmin_abs(grid, n, xs[k]) != argmin_abs_near(grid, n, xs[k], start, inv_step)) ++bad;
  return bad;
}
