// kernels.cuh -- the CUDA kernels of the hop path (sm_100a).
//
//   K1  count_rows_kernel / fill_rows_kernel   neighbour table: scatterer::find_neighbors of every site, once, as CSR
//   K4  create_excitons_kernel                 monte_carlo::kubo_create_particles / create_particles / repopulate
//   K2  kubo_flat_kernel                       nsteps x monte_carlo::kubo_step for every exciton + per-step sum of dx^2
//       reduce_partials_kernel                 block partials -> [nsteps][4] sums, fixed order (deterministic)
//
// All arithmetic is FP64 / integer; there is no dense contraction anywhere on this path, so tensor cores (tcgen05)
// do not apply.  The kernels are bound by dependent gathers into the L2-resident site and CSR tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "csr_core.h"
#include "hop_core.h"

namespace cntmc {

// ---- exciton population, struct of arrays ----------------------------------------------------------------------------
struct ExcitonArrays {
  double *  px, *py, *pz;  // particle::_pos
  double *  dx, *dy, *dz;  // particle::_delta_pos
  double*   ff;            // particle::_ff_time
  double *  ox, *oy, *oz;  // particle::_old_pos (contact mode only; null otherwise)
  int32_t*  site;          // particle::_scat_ptr
  uint8_t*  heading;       // particle::_heading_right
  uint32_t* ndraw;         // next draw index of the exciton's stream
  uint32_t* last_events;   // events in the previous launch (load-balancing key)
  uint64_t* gid;           // stream id (contact mode, where excitons are created and destroyed); null = first_gid + index
};

struct DrawConfig {
  uint64_t       seed;
  uint64_t       first_gid;
  const int64_t* replay_off;  // non-null selects replay
  const int32_t* replay_draws;
  const double*  replay_logs;
};

__device__ __forceinline__ void load_lane(Lane& L, const ExcitonArrays& S, const Tables& T, int64_t e) {
  L.px = S.px[e];
  L.py = S.py[e];
  L.pz = S.pz[e];
  L.dx = S.dx[e];
  L.dy = S.dy[e];
  L.dz = S.dz[e];
  L.ff = S.ff[e];
  L.site = S.site[e];
  L.heading_right = S.heading[e] != 0;
  L.ndraw = S.ndraw[e];
  L.nevent = 0;
  L.nreinject = 0;
  L.ncross = 0;
  L.nprobe = 0;
  L.stuck = false;
  const FlyRec f = load_fly(T.fly + L.site);
  L.left = f.left;
  L.right = f.right;
}
__device__ __forceinline__ void store_lane(const Lane& L, const ExcitonArrays& S, int64_t e) {
  S.px[e] = L.px;
  S.py[e] = L.py;
  S.pz[e] = L.pz;
  S.dx[e] = L.dx;
  S.dy[e] = L.dy;
  S.dz[e] = L.dz;
  S.ff[e] = L.ff;
  S.site[e] = L.site;
  S.heading[e] = L.heading_right ? 1 : 0;
  S.ndraw[e] = L.ndraw;
  S.last_events[e] = L.nevent;
}

template <typename Draws>
__device__ __forceinline__ void init_draws(Draws& D, const DrawConfig& dc, const ExcitonArrays& S, int64_t e);
template <>
__device__ __forceinline__ void init_draws<PhiloxDraws>(PhiloxDraws& D, const DrawConfig& dc, const ExcitonArrays& S, int64_t e) {
  D.init(dc.seed, S.gid ? S.gid[e] : dc.first_gid + (uint64_t)e);
}
template <>
__device__ __forceinline__ void init_draws<ReplayDraws>(ReplayDraws& D, const DrawConfig& dc, const ExcitonArrays& S, int64_t e) {
  const int64_t g = S.gid ? (int64_t)S.gid[e] : (int64_t)dc.first_gid + e;
  D.init(dc.replay_draws, dc.replay_logs, dc.replay_off[g], dc.replay_off[g + 1]);
}

enum { FLAG_STUCK = 0, FLAG_REPLAY = 1, FLAG_EMPTY_ROW = 2, FLAG_COUNT = 4 };
enum { CTR_REINJECT = 0, CTR_GUARD = 1, CTR_CROSS = 2, CTR_PROBE = 3, CTR_COUNT = 4 };

// ---- K4: creation ------------------------------------------------------------------------------------------------------
struct CreateArgs {
  Tables         T;
  ExcitonArrays  S;
  DrawConfig     draws;
  int64_t        P;
  const int32_t* site_list;  // injection region (kubo) or a slab / contact list
  int32_t        n_list;
  int32_t*       flags;
};

template <typename Draws>
__global__ void __launch_bounds__(256) create_excitons_kernel(const CreateArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.P) return;
  Lane  L;
  Draws D;
  init_draws(D, a.draws, a.S, e);
  create_exciton(L, a.T, D, a.site_list, a.n_list);
  store_lane(L, a.S, e);
  if (a.S.ox) {
    a.S.ox[e] = L.px;
    a.S.oy[e] = L.py;
    a.S.oz[e] = L.pz;
  }
  if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
}

// ---- K2: the hop kernel, Green-Kubo flavour -------------------------------------------------------------------------------
struct KuboArgs {
  Tables          T;
  ExcitonArrays   S;
  DrawConfig      draws;
  const uint32_t* perm;  // thread -> exciton (null = identity); groups excitons of similar activity into warps
  int64_t         P;
  double          dt;
  int32_t         nsteps;
  double*         partial;  // [gridDim.x][nsteps][4]: sum dx^2, dy^2, dz^2, events
  int32_t*        trace_sites;
  int32_t*        trace_counts;
  int32_t         trace_cap;
  int32_t*        flags;
  unsigned long long* counters;
};

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// One thread per exciton.  The loop is flat: an iteration moves every running lane forward by one scattering event or
// by the end of one time step, whichever comes first for that lane, so lanes of a warp may be in different time
// steps.  When lanes finish a step their squared displacements are summed over the lanes that share the step index
// (shuffle tree with zeros for the others, hence a fixed order) and added to the warp's row for that step in shared
// memory; nothing in the loop needs a block barrier or an atomic.
template <typename Draws>
__global__ void __launch_bounds__(128) kubo_flat_kernel(const KuboArgs a) {
  extern __shared__ double acc[];  // [warps][nsteps][4]
  const int    lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int    row = a.nsteps * 4;
  double*      wacc = acc + (size_t)warp * row;
  for (int k = lane; k < row; k += 32) wacc[k] = 0.0;
  __syncwarp();

  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool    active = t < a.P;
  const int64_t e = active ? (a.perm ? (int64_t)a.perm[t] : t) : 0;
  Lane          L{};
  Draws         D{};
  Cursor        c{};
  int32_t*      trace = nullptr;
  int32_t       trace_base = 0;
  if (active) {
    load_lane(L, a.S, a.T, e);
    init_draws(D, a.draws, a.S, e);
    c.step = 0;
    begin_step(c, L, a.dt);
    if (a.trace_sites) {  // the trace continues where the previous launch stopped
      trace_base = a.trace_counts[e];
      trace = a.trace_sites + e * (int64_t)a.trace_cap + trace_base;
    }
  } else {
    c.step = a.nsteps;
  }

  for (;;) {
    const bool running = c.step < a.nsteps;
    if (!__any_sync(kFullMask, running)) break;
    bool ended = false;
    if (running) ended = advance(L, a.T, D, c, trace, (uint32_t)(a.trace_cap - trace_base));
    unsigned m = __ballot_sync(kFullMask, ended);
    while (m) {
      const int      s0 = __shfl_sync(kFullMask, c.step, __ffs(m) - 1);
      const bool     mine = ended && (c.step == s0);
      const unsigned grp = __ballot_sync(kFullMask, mine);
      const double   sx = warp_sum(mine ? L.dx * L.dx : 0.0);
      const double   sy = warp_sum(mine ? L.dy * L.dy : 0.0);
      const double   sz = warp_sum(mine ? L.dz * L.dz : 0.0);
      const int      ev = __reduce_add_sync(kFullMask, mine ? (int)(L.nevent - c.ev0) : 0);
      if (lane == 0) {
        double* r = wacc + s0 * 4;
        r[0] += sx;
        r[1] += sy;
        r[2] += sz;
        r[3] += (double)ev;
      }
      m &= ~grp;
    }
    if (ended) {
      ++c.step;
      begin_step(c, L, a.dt);
    }
    if (L.stuck) c.step = a.nsteps;  // give up on this lane; the host reports the error
  }

  if (active) {
    store_lane(L, a.S, e);
    if (a.trace_counts) a.trace_counts[e] = trace_base + (int32_t)L.nevent;
    if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
    if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
    if (L.nreinject) atomicAdd(a.counters + CTR_REINJECT, (unsigned long long)L.nreinject);
  }
  {
    const unsigned nc = __reduce_add_sync(kFullMask, active ? L.ncross : 0u);
    const unsigned np = __reduce_add_sync(kFullMask, active ? L.nprobe : 0u);
    if (lane == 0) {
      atomicAdd(a.counters + CTR_CROSS, (unsigned long long)nc);
      atomicAdd(a.counters + CTR_PROBE, (unsigned long long)np);
    }
  }
  __syncthreads();
  double* out = a.partial + (size_t)blockIdx.x * row;
  for (int k = threadIdx.x; k < row; k += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += acc[(size_t)w * row + k];
    out[k] = s;
  }
}

// sums[s][c] = sum over blocks of partial[b][s][c], in a fixed order: thread j adds blocks j, j+T, j+2T, ... then a
// shared-memory tree.  One block per time step.
__global__ void __launch_bounds__(128) reduce_partials_kernel(const double* partial, int nblocks, int nsteps, double* sums) {
  __shared__ double sh[128][4];
  const int         s = blockIdx.x;
  double            v[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double* p = partial + ((size_t)b * nsteps + s) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] += p[c];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) sh[threadIdx.x][c] = v[c];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int c = 0; c < 4; ++c) sh[threadIdx.x][c] += sh[threadIdx.x + o][c];
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) sums[(size_t)s * 4 + threadIdx.x] = sh[0][threadIdx.x];
}

__global__ void iota_kernel(uint32_t* v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

// ---- K1: neighbour table ---------------------------------------------------------------------------------------------------
struct CsrArgs {
  const SiteGeom* geom;         // [N] site order
  const SiteGeom* cell_geom;    // [N] bucket order (candidates of one cell are contiguous)
  const int32_t*  cell_sites;   // [N] site index of each bucket-ordered entry, ascending inside a cell
  const int64_t*  cell_start;   // [ncell+1]
  int             nb[3];
  double          lo[3];
  double          radius;
  int64_t         N;
  RateTable       R;
  // outputs
  uint32_t*       deg;        // [N]      (count pass)
  const uint64_t* row_begin;  // [N+1]    (fill pass; exclusive scan of deg)
  int32_t*        nbr;
  double*         cum;
  HopRec*         hop;
  int32_t*        flags;
  unsigned long long* counters;
};

// Enumerate the candidates of site i in the reference's order: the 27-cell stencil with x outermost and z innermost
// (monte_carlo.h:402-411), sites of a cell in ascending list order (monte_carlo.h:389-395).
template <bool kFill>
__global__ void __launch_bounds__(128) csr_rows_kernel(const CsrArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const SiteGeom s1 = a.geom[i];
  const int      cx = cell_coord(s1.px, a.lo[0], a.radius), cy = cell_coord(s1.py, a.lo[1], a.radius),
            cz = cell_coord(s1.pz, a.lo[2], a.radius);
  uint32_t d = 0;
  double   acc = 0.0;
  uint64_t base = 0;
  bool     guard = false;
  if (kFill) base = a.row_begin[i];
  for (int ix = cx - 1; ix <= cx + 1; ++ix)
    for (int iy = cy - 1; iy <= cy + 1; ++iy)
      for (int iz = cz - 1; iz <= cz + 1; ++iz) {
        if (!(ix > -1 && ix < a.nb[0] && iy > -1 && iy < a.nb[1] && iz > -1 && iz < a.nb[2])) continue;
        const int64_t b = (int64_t)ix + (int64_t)iy * a.nb[0] + (int64_t)iz * a.nb[0] * a.nb[1];
        const int64_t q1 = a.cell_start[b + 1];
        for (int64_t q = a.cell_start[b]; q < q1; ++q) {
          const SiteGeom s2 = a.cell_geom[q];
          if (!within_cutoff(s1, s2, a.radius)) continue;
          if (kFill) {
            const double rate = pair_rate(s1, s2, a.R, &guard);
            acc = (d == 0) ? rate : acc + rate;  // scatterer.cpp:78-80, sequential
            a.nbr[base + d] = a.cell_sites[q];
            a.cum[base + d] = acc;
          }
          ++d;
        }
      }
  if (!kFill) {
    a.deg[i] = d;
  } else {
    HopRec h;
    h.total = acc;                    // scatterer.h:91  _max_rate = neighbors.back().first
    h.inv_total = (d ? 1. / acc : 0.0);  // scatterer.h:92
    h.row_begin = (uint32_t)base;
    h.row_len = d;
    h.pad[0] = h.pad[1] = 0;
    a.hop[i] = h;
    if (d == 0) atomicOr(a.flags + FLAG_EMPTY_ROW, 1);
    if (guard) atomicAdd(a.counters + CTR_GUARD, 1ULL);
  }
}

}  // namespace cntmc
