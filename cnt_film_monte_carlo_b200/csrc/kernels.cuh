// kernels.cuh -- the CUDA kernels of the hop path (sm_100a).
//
//   K1  csr_rows_kernel<false/true>   neighbour table: scatterer::find_neighbors of every site, once, as CSR
//   K4  create_excitons_kernel        monte_carlo::kubo_create_particles / create_particles / repopulate
//   K2  kubo_kernel                   nsteps x monte_carlo::kubo_step for every exciton (persistent warps, lanes refill
//                                     from activity-class lists)
//       reduce_stage_kernel           per-(step, exciton) squared displacements -> [nsteps][4] sums, fixed order
//
// All arithmetic is FP64 / integer; there is no dense contraction anywhere on this path, so tensor cores (tcgen05)
// do not apply.  The kernels are bound by instruction issue and by dependent gathers into the L2-resident site and CSR
// tables (DESIGN.md has the measurements).
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
  int32_t*  site;          // particle::_scat_ptr
  uint8_t*  heading;       // particle::_heading_right
  uint32_t* ndraw;         // next draw index of the exciton's stream
  uint32_t* last_events;   // events in the previous launch (load-balancing key)
  uint64_t* gid;           // stream id where excitons are created and destroyed (contact mode); null = first_gid + index
};

struct DrawConfig {
  uint64_t       seed;
  uint64_t       first_gid;
  const int64_t* replay_off;  // non-null selects replay
  const int32_t* replay_draws;
  const double*  replay_logs;
};

// Exciton state streams through once per launch (ld.cs).  kCoherent: read through L2 only (ld.cg) -- in the two-engine
// kernel an exciton may come back to an SM that has seen an older version of it, and L1 is not coherent.
template <bool kCoherent = false>
__device__ __forceinline__ void load_lane(Lane& L, const ExcitonArrays& S, const Tables& T, int64_t e) {
  auto ld = [](auto* p) { return kCoherent ? __ldcg(p) : __ldcs(p); };
  L.px = ld(S.px + e);
  L.py = ld(S.py + e);
  L.pz = ld(S.pz + e);
  L.dx = ld(S.dx + e);
  L.dy = ld(S.dy + e);
  L.dz = ld(S.dz + e);
  L.ff = ld(S.ff + e);
  L.site = ld(S.site + e);
  L.heading_right = ld(S.heading + e) != 0;
  L.ndraw = ld(S.ndraw + e);
  L.nevent = 0;
  L.stuck = false;
  attach_site(L, T);
}
__device__ __forceinline__ void store_lane(const Lane& L, const ExcitonArrays& S, int64_t e) {  // needs L.pos_valid
  __stcs(S.px + e, L.px);
  __stcs(S.py + e, L.py);
  __stcs(S.pz + e, L.pz);
  __stcs(S.dx + e, L.dx);
  __stcs(S.dy + e, L.dy);
  __stcs(S.dz + e, L.dz);
  __stcs(S.ff + e, L.ff);
  __stcs(S.site + e, L.site);
  __stcs(S.heading + e, (uint8_t)(L.heading_right ? 1 : 0));
  __stcs(S.ndraw + e, L.ndraw);
  __stcs(S.last_events + e, L.nevent);
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
enum {
  CTR_REINJECT = 0, CTR_GUARD = 1, CTR_CROSS = 2, CTR_PROBE = 3, CTR_QUEUE = 4, CTR_EVENTS = 5, CTR_FAST = 6,
  // instrumented hop kernel only: warp residency (ns summed over warps), kernel span as seen from the device
  // (~first start and last exit, stored for atomicMax), number of warps, lane-iterations with / without an exciton
  CTR_WARP_NS = 8, CTR_T_FIRST_INV = 9, CTR_T_LAST = 10, CTR_WARPS = 11, CTR_LANE_BUSY = 12, CTR_LANE_IDLE = 13,
  // two-engine kernel, instrumented: ns per loop phase of lane 0, [engine][phase], and iterations per engine
  CTR_PHASE = 16, CTR_ITER = 28, CTR_COUNT = 32
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr unsigned kFullMask = 0xffffffffu;

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
  Lane  L{};
  Draws D{};
  init_draws(D, a.draws, a.S, e);
  create_exciton(L, a.T, D, a.site_list, a.n_list);
  store_lane(L, a.S, e);
  a.S.last_events[e] = 0;
  if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
}

// ---- K2: the hop kernel, Green-Kubo flavour -------------------------------------------------------------------------------
// Activity classes.  Exciton activity is extremely skewed (on the C2 film the median exciton scatters twice in 64 steps,
// the 99th percentile 1400 times: a few excitons sit in traps between closely crossing tubes) and it is predictable from
// the state: Gamma(site) * dt = expected events per step if the exciton stays where it is.  When a lane stores an
// exciton it files it under one of four classes; in the next launch "hot" blocks serve the two active classes and
// "cold" blocks the two quiet ones, so that warps mostly hold excitons that run the same branch of the loop.
constexpr int kClasses = 4;
__device__ __forceinline__ int activity_class(double expected_events_per_step) {
  return expected_events_per_step >= 8.0 ? 3 : expected_events_per_step >= 1.0 ? 2 : expected_events_per_step >= 0.125 ? 1 : 0;
}
struct ClassLists {
  const uint32_t*     list[kClasses];       // excitons of each class, filed by the previous launch
  const uint32_t*     count;                // [kClasses]
  unsigned long long* head;                 // [kClasses] next unassigned position of each list
  uint32_t*           next_list[kClasses];  // being filled for the next launch
  uint32_t*           next_count;           // [kClasses]
};

// file exciton e of the lanes flagged `mine` under their class `cls` (warp-aggregated append; call with the whole warp)
__device__ __forceinline__ void file_excitons(const ClassLists& q, bool mine, int cls, uint32_t e, int lane, unsigned lt_mask) {
#pragma unroll
  for (int c = 0; c < kClasses; ++c) {
    const unsigned m = __ballot_sync(0xffffffffu, mine && cls == c);
    if (m) {
      uint32_t  base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(q.next_count + c, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (mine && cls == c) q.next_list[c][base + __popc(m & lt_mask)] = e;
    }
  }
}
// hand an exciton to every lane that `want`s one, from the two classes of the warp's current role: hot warps serve
// classes 3 then 2, cold warps classes 0 then 1 (call with the whole warp).  Returns false for lanes left without work.
__device__ __forceinline__ bool take_exciton(const ClassLists& q, bool want, bool hot_role, int lane, unsigned lt_mask, int64_t& e) {
  bool     got = false;
  unsigned need = __ballot_sync(0xffffffffu, want);
  for (int k = 0; k < 2 && need; ++k) {
    const int          c = hot_role ? kClasses - 1 - k : k;
    const int          n = __popc(need), leader = __ffs(need) - 1;
    const long long    cnt = (long long)q.count[c];
    unsigned long long base = 0;
    // (an exhausted list is not touched again: idle lanes of the two-engine kernel ask repeatedly)
    if (lane == leader)
      base = (cnt > 0 && *(volatile unsigned long long*)(q.head + c) < (unsigned long long)cnt) ? atomicAdd(q.head + c, (unsigned long long)n)
                                                                                               : (unsigned long long)cnt;
    base = __shfl_sync(0xffffffffu, base, leader);
    const long long avail = cnt - (long long)base;
    if (want && !got) {
      const int rank = __popc(need & lt_mask);
      if ((long long)rank < avail) {
        e = (int64_t)q.list[c][base + (unsigned long long)rank];
        got = true;
      }
    }
    need = __ballot_sync(0xffffffffu, want && !got);
  }
  return got;
}


// ---- hand-over queues of the two-engine hop kernel -----------------------------------------------------------------------
// A bounded multi-producer multi-consumer queue of exciton indices (Vyukov's scheme, one 64-bit word per cell:
// sequence number in the high half, exciton index in the low half).  Cell i starts as (i, -); the producer of position
// p waits for (p, -), publishes (p+1, exciton) with release semantics -- the state it stored for the exciton is visible
// to whoever reads the cell -- and the consumer of position p waits for (p+1, .), then leaves (p + capacity, -).  The
// capacity is at least the number of excitons and an exciton waits in at most one queue, so producers never wait.
// head and tail keep counting from launch to launch.  Consumers read exciton state with ld.cg (L2): L1 is not coherent.
struct Ring {
  unsigned long long* cell;
  unsigned long long* head;
  unsigned long long* tail;
  long long*          avail;  // positions produced and not yet claimed by a consumer (a semaphore: no compare-and-swap loops)
  uint32_t            mask;   // capacity - 1
};
struct Cursors {  // where in the launch a handed-over exciton is
  int32_t*  step;
  double*   dt_rem;
  double *  ox, *oy, *oz;  // position at the start of its current time step (_old_pos)
  uint32_t* ev;            // events since the start of its current time step
};
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int kSpinLimit = 1 << 24;  // polls before a wait is declared stuck (reported by the host; never seen)
// append exciton e of the lanes flagged `mine` (call with the whole warp; the lanes' state stores come first)
__device__ __forceinline__ bool ring_push(const Ring& r, bool mine, uint32_t e, int lane, unsigned lt_mask) {
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (!m) return true;
  unsigned long long base = 0;
  const int          leader = __ffs(m) - 1;
  if (lane == leader) {
    base = atomicAdd(r.tail, (unsigned long long)__popc(m));
    atomicAdd((unsigned long long*)r.avail, (unsigned long long)__popc(m));  // consumers wait on the cell for the index itself
  }
  base = __shfl_sync(0xffffffffu, base, leader);
  bool ok = true;
  if (mine) {
    const unsigned long long pos = base + (unsigned long long)__popc(m & lt_mask);
    unsigned long long*      c = r.cell + (pos & r.mask);
    int                      spin = 0;
    while ((uint32_t)(ld_relaxed_u64(c) >> 32) != (uint32_t)pos && ++spin < kSpinLimit) {
    }
    ok = spin < kSpinLimit;
    st_release_u64(c, ((unsigned long long)(uint32_t)(pos + 1) << 32) | e);
  }
  return ok;
}
// hand an exciton to every lane that `want`s one, as far as the queue holds any (call with the whole warp)
__device__ __forceinline__ bool ring_pop(const Ring& r, bool want, int lane, unsigned lt_mask, uint32_t& e, bool& ok) {
  const unsigned need = __ballot_sync(0xffffffffu, want);
  if (!need) return false;
  const int          leader = __ffs(need) - 1;
  unsigned long long base = 0;
  unsigned           take = 0;
  if (lane == leader && (long long)ld_relaxed_u64((const unsigned long long*)r.avail) > 0) {
    // claim up to n of the available positions: fetch-and-add, hand back what was not there (thousands of warps use
    // this word; a compare-and-swap loop would convoy)
    const long long n = (long long)__popc(need);
    const long long old = (long long)atomicAdd((unsigned long long*)r.avail, (unsigned long long)(-n));
    const long long got = old <= 0 ? 0 : (old < n ? old : n);
    if (got < n) atomicAdd((unsigned long long*)r.avail, (unsigned long long)(n - got));
    if (got > 0) {
      base = atomicAdd(r.head, (unsigned long long)got);
      take = (unsigned)got;
    }
  }
  base = __shfl_sync(0xffffffffu, base, leader);
  take = __shfl_sync(0xffffffffu, take, leader);
  bool got = false;
  if (want && (unsigned)__popc(need & lt_mask) < take) {
    const unsigned long long pos = base + (unsigned long long)__popc(need & lt_mask);
    unsigned long long*      c = r.cell + (pos & r.mask);
    unsigned long long       v;
    int                      spin = 0;
    while ((uint32_t)((v = ld_relaxed_u64(c)) >> 32) != (uint32_t)(pos + 1) && ++spin < kSpinLimit) {
    }
    if (spin >= kSpinLimit) ok = false;
    e = (uint32_t)(v & 0xffffffffULL);
    st_relaxed_u64(c, (unsigned long long)(uint32_t)(pos + (unsigned long long)r.mask + 1ULL) << 32);
    got = true;
  }
  return got;
}
__device__ __forceinline__ bool ring_has_work(const Ring& r) { return (long long)ld_relaxed_u64((const unsigned long long*)r.avail) > 0; }
__global__ void ring_init_kernel(unsigned long long* cell, uint32_t capacity, unsigned long long first) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= capacity) return;
  // the cell of position p >= first with p % capacity == i
  const unsigned long long p = first + ((unsigned long long)i + capacity - (first % capacity)) % capacity;
  cell[i] = (unsigned long long)(uint32_t)p << 32;
}

#if defined(CNTMC_PROFILE_SEGMENTS)
constexpr int kWarpTimeCols = 12;  // + cycles per loop segment of lane 0
#else
constexpr int kWarpTimeCols = 4;
#endif
struct alignas(32) StageRec {  // one full 32-byte sector per (step, exciton)
  double dx2, dy2, dz2, events;
};

struct KuboArgs {
  Tables              T;
  ExcitonArrays       S;
  DrawConfig          draws;
  ClassLists          q;
  int32_t             hot_blocks;  // blocks [0, hot_blocks) serve the active classes first
  int32_t             fast_rounds; // events decided by the top entries of the site record, per iteration and lane (0 = off)
  // two-engine kernel (kubo_engines_kernel): hand-over queues and the cursors of the excitons waiting in them
  int32_t             refill_min;          // idle lanes that make a warp fetch new excitons
  int32_t             park_min, park_max;  // flight warps: lanes with an event due that make the warp scatter; iterations a lane waits at most
  Ring                ring[2];     // [0]: excitons waiting for a flight warp, [1]: for an event warp
  unsigned long long* remaining;   // excitons that have not completed the launch yet
  Cursors             cur;
  int64_t             P;
  double              dt;
  int32_t             nsteps;
  StageRec*           stage;  // [nsteps][P] by exciton index
  int32_t*            trace_sites;
  int32_t*            trace_counts;
  int32_t             trace_cap;
  int32_t*            flags;
  unsigned long long* counters;
  unsigned long long* warp_times;  // diagnostics (instrumented kernel): [warps][4] = enter, first failed take, exit, role
};

// Persistent warps; every lane owns one exciton at a time, carries it through all nsteps time steps of the launch, files
// it under its activity class for the next launch and takes another one (see ClassLists).
//
// The loop is flat: an iteration moves every busy lane forward by one scattering event or by the end of one time step,
// whichever comes first for that lane; lanes of a warp are in general in different time steps of different excitons.
// Nothing in the loop needs a barrier, shared memory or a floating-point atomic: when a lane ends a step it writes its
// squared displacement to the (step, exciton) record, and reduce_stage_kernel sums the records in a fixed order
// afterwards, so the ensemble sums do not depend on which lane happened to process which exciton, nor on any tuning
// option.
// kInstr adds what only tests and the roofline bookkeeping need (site traces, probe / crossing counters).
template <typename Draws, int kMinBlocks, bool kInstr>
__global__ void __launch_bounds__(128, kMinBlocks) kubo_kernel(const KuboArgs a) {
  // The displacement accumulator and the position at the start of the step are only touched when a time step ends;
  // they live in shared memory (one slot per thread) so that the event path does not carry 12 registers of them.
  __shared__ double s_delta[3][128], s_old[3][128];
  const int      tid = threadIdx.x, lane = threadIdx.x & 31;
  const bool     fast_path = a.fast_rounds > 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool     hot_role = (int)blockIdx.x < a.hot_blocks;
  uint32_t       e = 0;
  Lane           L{};
  Draws          D{};
  double         dt_rem = 0.0;
  int32_t        step = 0;
  uint32_t       ev0 = 0;
  int32_t*       trace = nullptr;
  int32_t        trace_base = 0;
  unsigned long long t_enter = 0, it_busy = 0, it_idle = 0, t_dry = 0;
  int                iter = 0;
  if (kInstr) t_enter = global_ns();
#if defined(CNTMC_PROFILE_SEGMENTS)
  L.seg_t = clock64();
#endif

  auto start = [&]() {
    const uint32_t nc = L.ncross, np = L.nprobe, nr = L.nreinject, nf = L.nfast;
    load_lane(L, a.S, a.T, (int64_t)e);
    L.ncross = nc;
    L.nprobe = np;
    L.nreinject = nr;
    L.nfast = nf;
    init_draws(D, a.draws, a.S, (int64_t)e);
    step = 0;
    dt_rem = a.dt;
    ev0 = 0;
    s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
    s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;  // _old_pos = _pos (particle.cpp:59)
    if (kInstr && a.trace_sites) {  // the trace continues where the previous launch stopped
      trace_base = a.trace_counts[e];
      trace = a.trace_sites + (int64_t)e * a.trace_cap + trace_base;
    }
  };
  // A warp keeps to one role so that its lanes run the same branch of the loop most of the time (hot: scattering
  // events, cold: chain walks and step ends); it changes role only once, when all its lanes have run dry.
  bool role = hot_role;
  for (int pass = 0; pass < 2; ++pass, role = !role) {
    int64_t e64 = 0;
    bool    have = take_exciton(a.q, true, role, lane, lt_mask, e64);
    if (have) {
      e = (uint32_t)e64;
      start();
    }
    while (__any_sync(kFullMask, have)) {
      bool finished = false;
      if (kInstr) {
        if (have) ++it_busy; else ++it_idle;
        ++iter;
      }
      // One iteration = up to two operations per lane: first the end of a time step for the lanes whose free flight
      // outlasts the step, then a scattering event for the lanes whose flight ends inside the (possibly new) step.
      // The warp pays the latency of both code paths anyway whenever both kinds are present, so a lane that ends a
      // step and scatters right away gets both done in the same pass.
      CNTMC_SEG(L, 6);  // loop head, refill
      if (have && !(L.ff <= dt_rem)) {  // particle.cpp:62 false: the flight outlasts the step
        const double t = dt_rem;
        const Leg    leg = fly(L, a.T, t, true);
        L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
        after_flight_step_end(L, a.T, D, leg, t, s_old[0][tid], s_old[1][tid], s_old[2][tid]);
        // streaming stores: written once, read once by the reduction, must not evict the tables from L2
        double2* rec = reinterpret_cast<double2*>(a.stage + ((size_t)step * (size_t)a.P + (size_t)e));
        __stcs(rec, make_double2(L.dx * L.dx, L.dy * L.dy));  // std::pow(delta_pos, 2), monte_carlo.cpp:397-399
        __stcs(rec + 1, make_double2(L.dz * L.dz, (double)(L.nevent - ev0)));
        s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
        s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;
        ++step;
        dt_rem = a.dt;
        ev0 = L.nevent;
        finished = (step >= a.nsteps);
      }
      CNTMC_SEG(L, 5);  // step-end path (own, or waiting for the lanes that run it)
      // Events that the record of the exciton's site decides (fast_event): short straight-line code, one trip to
      // memory, no search and no chain walk.  The lanes that can take it run a few rounds of it together; the other
      // lanes wait, and whatever a lane cannot do this way is left to the ordinary event path below.
      for (int k = 1; k < a.fast_rounds; ++k) {  // (fast_rounds = 1: the top entries are only consulted inside the ordinary event path)
        bool did = false;
        if (have && !finished && (L.ff <= dt_rem))
          did = fast_event(L, a.T, D, dt_rem, kInstr ? trace : nullptr, kInstr ? (uint32_t)(a.trace_cap - trace_base) : 0u);
        if (!__any_sync(kFullMask, did)) break;
      }
      if (have && !finished && (L.ff <= dt_rem)) {
        const double t = L.ff;
        const Leg    leg = fly(L, a.T, t, false);
        dt_rem -= t;  // particle.cpp:63
        after_flight_scatter(L, a.T, D, leg, kInstr ? trace : nullptr, kInstr ? (uint32_t)(a.trace_cap - trace_base) : 0u, fast_path);
      }
      CNTMC_SEG(L, 7);  // waiting for the other lanes of the warp to finish their events
      finished = have && (finished || L.stuck);
      if (__any_sync(kFullMask, finished)) {
        int cls = 0;
        if (finished) {
          materialize(L, a.T);
          L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
          store_lane(L, a.S, (int64_t)e);
          cls = activity_class(hop_info(L, a.T).total * a.dt);
          if (kInstr && a.trace_counts) a.trace_counts[e] = trace_base + (int32_t)L.nevent;
          if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
          if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
        }
        file_excitons(a.q, finished, cls, e, lane, lt_mask);
        const bool got = take_exciton(a.q, finished, role, lane, lt_mask, e64);
        if (finished) {
          have = got;
          if (have) {
            e = (uint32_t)e64;
            start();
          }
        }
        if (kInstr && t_dry == 0 && __any_sync(kFullMask, finished && !got)) t_dry = global_ns();
      }
    }
  }

  const unsigned nr = __reduce_add_sync(kFullMask, L.nreinject);
  if (lane == 0 && nr) atomicAdd(a.counters + CTR_REINJECT, (unsigned long long)nr);
  if (kInstr) {
    const unsigned nc = __reduce_add_sync(kFullMask, L.ncross);
    const unsigned np = __reduce_add_sync(kFullMask, L.nprobe);
    const unsigned nf = __reduce_add_sync(kFullMask, L.nfast);
    if (lane == 0 && nf) atomicAdd(a.counters + CTR_FAST, (unsigned long long)nf);
    const unsigned long long t_exit = global_ns();
    atomicAdd(a.counters + CTR_LANE_BUSY, it_busy);
    atomicAdd(a.counters + CTR_LANE_IDLE, it_idle);
    if (lane == 0) {
      if (nc) atomicAdd(a.counters + CTR_CROSS, (unsigned long long)nc);
      if (np) atomicAdd(a.counters + CTR_PROBE, (unsigned long long)np);
      atomicAdd(a.counters + CTR_WARP_NS, t_exit - t_enter);
      atomicMax(a.counters + CTR_T_FIRST_INV, ~t_enter);
      atomicMax(a.counters + CTR_T_LAST, t_exit);
      atomicAdd(a.counters + CTR_WARPS, 1ULL);
      if (a.warp_times) {
        unsigned long long* w = a.warp_times + kWarpTimeCols * ((size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5));
        w[0] = t_enter;
        w[1] = t_dry ? t_dry : t_exit;
        w[2] = t_exit;
        w[3] = (hot_role ? 1ull : 0ull) | ((unsigned long long)iter << 8);
#if defined(CNTMC_PROFILE_SEGMENTS)
        for (int k = 0; k < 8; ++k) w[4 + k] = (unsigned long long)L.seg[k];  // lane 0 of the warp
#endif
      }
    }
  }
}


// ---- K2, two-engine form -------------------------------------------------------------------------------------------------------
// The same per-exciton arithmetic as kubo_kernel, scheduled so that the lanes of a warp run the same code.
//
// On the bench workload 97 % of the scattering events happen on a few "hot" sites, one after the other (tens of events
// per time step), while 97 % of the excitons are in free flight and need one end-of-step operation per time step.
// Which of the two an exciton needs changes inside a launch (a flight ends on a hot site; an exciton leaves one), and a
// warp that holds both kinds pays for both code paths in every iteration with one or two lanes active.  So warps are
// specialised and excitons change hands:
//   * a FLIGHT warp ends time steps (chain walk, last leg, displacement, removal box, staging record).  The rare events
//     of its excitons are gathered and run for several lanes at once; an exciton that scatters onto a hot site is
//     handed over -- state and cursor to memory, index to ring[1] -- and the lane takes another.
//   * an EVENT warp scatters: a few rounds of fast_event (decided by the site record, straight-line code), then the
//     ordinary event path for what is left, and the step ends of its own excitons.  An exciton that ends a step on a
//     quiet site with no event due within the next step goes back through ring[0].
// A warp refills its lanes from its engine's queue and the launch's class lists; a warp that has run empty joins the
// engine that has work, and leaves when every exciton has completed the launch (`remaining`).  Results cannot depend on
// any of this: an exciton's trajectory is a function of its own state and stream, and the ensemble sums are reduced
// from the (step, exciton) records in a fixed order.
template <typename Draws, int kMinBlocks, bool kInstr>
__global__ void __launch_bounds__(128, kMinBlocks) kubo_engines_kernel(const KuboArgs a) {
  __shared__ double s_delta[3][128], s_old[3][128];
  const int      tid = threadIdx.x, lane = threadIdx.x & 31;
  const bool     fast_path = a.fast_rounds > 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t       e = 0;
  Lane           L{};
  Draws          D{};
  double         dt_rem = 0.0;
  int32_t        step = 0;
  uint32_t       ev0 = 0;
  int32_t*       trace = nullptr;
  int32_t        trace_base = 0;
  bool           ok = true;
  unsigned long long t_enter = 0, it_busy = 0, it_idle = 0;
  int                iter = 0;
  if (kInstr) t_enter = global_ns();
  // instrumented build: time of lane 0 per phase {refill, step ends, fast events, ordinary events, hand-over, waiting}
  unsigned long long ph_t = kInstr ? global_ns() : 0ULL, ph_ns[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}}, ph_it[2] = {0, 0};
  auto mark = [&](int v, int k) {
    if (kInstr) {
      const unsigned long long now = global_ns();
      ph_ns[v][k] += now - ph_t;
      ph_t = now;
    }
  };

  auto start = [&](bool resumed, bool event_engine) {
    const uint32_t nc = L.ncross, np = L.nprobe, nr = L.nreinject, nf = L.nfast;
    load_lane<true>(L, a.S, a.T, (int64_t)e);
    L.ncross = nc;
    L.nprobe = np;
    L.nreinject = nr;
    L.nfast = nf;
    init_draws(D, a.draws, a.S, (int64_t)e);
    s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
    if (resumed) {
      step = __ldcg(a.cur.step + e);
      dt_rem = __ldcg(a.cur.dt_rem + e);
      s_old[0][tid] = __ldcg(a.cur.ox + e); s_old[1][tid] = __ldcg(a.cur.oy + e); s_old[2][tid] = __ldcg(a.cur.oz + e);
      ev0 = 0u - __ldcg(a.cur.ev + e);  // L.nevent restarts at 0: nevent - ev0 keeps counting the events of the step
    } else {
      step = 0;
      dt_rem = a.dt;
      ev0 = 0;
      s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;  // _old_pos = _pos (particle.cpp:59)
    }
    if (event_engine && fast_path && L.at_site) {  // fetch the rate fields now: the next event can take the fast path
      (void)hop_info(L, a.T);
      prefetch_l1(a.T.top + L.site);
    }
    if (kInstr && a.trace_sites) {  // the trace continues where the previous owner stopped
      trace_base = __ldcg(a.trace_counts + e);
      trace = a.trace_sites + (int64_t)e * a.trace_cap + trace_base;
    }
  };

  const bool dedicated = (int)blockIdx.x < a.hot_blocks;  // these blocks stay with the event engine: arrivals find a warp
  bool       role_v = dedicated;                          // event engine?
  bool       have = false;
  bool       try_fill = true;
  bool       f_lists_dry = false, v_lists_dry = false;  // the launch's class lists never refill once they are empty
  int        polls = 0, idle_polls = 0, park_age = 0;
  for (;;) {
    // ---- refill: migrated excitons of this engine first, then the launch's class lists
    // (a refill is a chain of trips to memory -- queue counters, list, exciton state, site records -- that the whole warp
    // waits for: it is worth it for several idle lanes at once, or when nothing else is left to do)
    const unsigned idle_m = __ballot_sync(kFullMask, !have);
    if (try_fill && idle_m && (__popc(idle_m) >= a.refill_min || (iter & 15) == 0 || idle_m == kFullMask)) {
      try_fill = false;
      uint32_t   e_new = 0;
      const bool from_ring = ring_pop(a.ring[role_v ? 1 : 0], !have, lane, lt_mask, e_new, ok);
      int64_t    e64 = 0;
      bool       from_list = false;
      if (!(role_v ? v_lists_dry : f_lists_dry)) {
        from_list = take_exciton(a.q, !have && !from_ring, role_v, lane, lt_mask, e64);
        if (__any_sync(kFullMask, !have && !from_ring && !from_list)) {
          if (role_v) v_lists_dry = true; else f_lists_dry = true;
        }
      }
      if (from_ring || from_list) {
        e = from_ring ? e_new : (uint32_t)e64;
        start(from_ring, role_v);
        have = true;
      }
    }
    mark(role_v, 0);
    if (!__any_sync(kFullMask, have)) {
      // the warp is empty: join the engine that has work, leave when the launch is complete
      // (every counter sits on its own 128-byte line and idle warps back off: thousands of them may be polling)
      bool v_work = false, f_work = false, all_done = false;
      if (lane == 0) {
        v_work = !v_lists_dry || ring_has_work(a.ring[1]);
        f_work = !f_lists_dry || ring_has_work(a.ring[0]);
        all_done = ld_relaxed_u64(a.remaining) == 0ULL;
      }
      v_work = __shfl_sync(kFullMask, v_work, 0);
      f_work = __shfl_sync(kFullMask, f_work, 0);
      all_done = __shfl_sync(kFullMask, all_done, 0);
      try_fill = true;
      if (role_v ? v_work : f_work) {
        idle_polls = 0;
        continue;
      }
      if ((role_v ? f_work : v_work) && !(dedicated && role_v)) {
        role_v = !role_v;
        idle_polls = 0;
        continue;
      }
      if (all_done || !ok) break;
      if (++polls > (1 << 18)) {  // seconds: something is lost; report instead of hanging
        ok = false;
        break;
      }
      __nanosleep(idle_polls < 4 ? (2000u << idle_polls) : 32000u);
      ++idle_polls;
      mark(role_v, 5);
      continue;
    }
    if (kInstr) {
      if (have) ++it_busy; else ++it_idle;
      ++iter;
    }
    if ((iter & 15) == 0) try_fill = true;  // idle lanes look for new arrivals every few iterations
    if (!kInstr) ++iter;
    if (kInstr) ++ph_it[role_v];

    bool finished = false, handoff = false;
    // ---- end of a time step, for the lanes whose flight outlasts it (both engines)
    if (have && !(L.ff <= dt_rem)) {  // particle.cpp:62 false
      const double t = dt_rem;
      const Leg    leg = fly(L, a.T, t, true);
      L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
      after_flight_step_end(L, a.T, D, leg, t, s_old[0][tid], s_old[1][tid], s_old[2][tid]);
      double2* rec = reinterpret_cast<double2*>(a.stage + ((size_t)step * (size_t)a.P + (size_t)e));
      __stcs(rec, make_double2(L.dx * L.dx, L.dy * L.dy));  // std::pow(delta_pos, 2), monte_carlo.cpp:397-399
      __stcs(rec + 1, make_double2(L.dz * L.dz, (double)(L.nevent - ev0)));
      s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
      s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;
      ++step;
      dt_rem = a.dt;
      ev0 = L.nevent;
      finished = (step >= a.nsteps);
      // an exciton that went quiet leaves the event engine: nothing due within the next step, and not on a hot site
      if (role_v && !finished && L.ff > a.dt && hop_info(L, a.T).total * a.dt < 1.0) handoff = true;
    }
    mark(role_v, 1);
    if (role_v) {
      // ---- events decided by the site record: a few rounds of short straight-line code for the lanes that can
      for (int k = 1; k < a.fast_rounds; ++k) {  // (fast_rounds = 1: the top entries are only consulted inside the ordinary event path)
        bool did = false;
        if (have && !finished && !handoff && (L.ff <= dt_rem))
          did = fast_event(L, a.T, D, dt_rem, kInstr ? trace : nullptr, kInstr ? (uint32_t)(a.trace_cap - trace_base) : 0u);
        if (!__any_sync(kFullMask, did)) break;
      }
      mark(role_v, 2);
      // ---- the ordinary event path for what is left
      if (have && !finished && !handoff && (L.ff <= dt_rem)) {
        const double t = L.ff;
        const Leg    leg = fly(L, a.T, t, false);
        dt_rem -= t;  // particle.cpp:63
        after_flight_scatter(L, a.T, D, leg, kInstr ? trace : nullptr, kInstr ? (uint32_t)(a.trace_cap - trace_base) : 0u, fast_path);
      }
    } else {
      // A flight warp scatters rarely (its excitons sit on quiet sites: an event every fifty steps or so).  Lanes with
      // an event due wait until a few of them have gathered, so that the event path runs once for several lanes
      // instead of in every other iteration for one; an exciton whose event lands it on a hot site changes engine.
      const bool     due = have && !finished && (L.ff <= dt_rem);
      const unsigned dm = __ballot_sync(kFullMask, due);
      if (dm) {
        const unsigned busy = __ballot_sync(kFullMask, have && !finished && !due);
        if (__popc(dm) >= a.park_min || park_age >= a.park_max || busy == 0u) {
          park_age = 0;
          if (due) {
            const double t = L.ff;
            const Leg    leg = fly(L, a.T, t, false);
            dt_rem -= t;  // particle.cpp:63
            after_flight_scatter(L, a.T, D, leg, kInstr ? trace : nullptr, kInstr ? (uint32_t)(a.trace_cap - trace_base) : 0u);
            if (hop_info(L, a.T).total * a.dt >= 1.0) handoff = true;
          }
        } else {
          ++park_age;
        }
      }
    }

    mark(role_v, 3);
    // ---- lanes that give their exciton up: it completed the launch, or it changes engine
    const bool leave = have && (finished || handoff || L.stuck);
    if (__any_sync(kFullMask, leave)) {
      int cls = 0;
      const bool done = leave && !(handoff && !L.stuck);
      if (leave) {
        materialize(L, a.T);
        L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
        store_lane(L, a.S, (int64_t)e);
        if (kInstr && a.trace_counts) __stcg(a.trace_counts + e, trace_base + (int32_t)L.nevent);
        if (done) {
          cls = activity_class(hop_info(L, a.T).total * a.dt);
          if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
          if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
        } else {
          __stcg(a.cur.step + e, step);
          __stcg(a.cur.dt_rem + e, dt_rem);
          __stcg(a.cur.ox + e, s_old[0][tid]); __stcg(a.cur.oy + e, s_old[1][tid]); __stcg(a.cur.oz + e, s_old[2][tid]);
          __stcg(a.cur.ev + e, L.nevent - ev0);
        }
      }
      file_excitons(a.q, done, cls, e, lane, lt_mask);
      if (!ring_push(a.ring[role_v ? 0 : 1], leave && !done, e, lane, lt_mask)) ok = false;
      const unsigned nd = __ballot_sync(kFullMask, done);
      if (nd && lane == 0) atomicAdd(a.remaining, 0ULL - (unsigned long long)__popc(nd));
      if (leave) have = false;
      try_fill = true;
    }
    mark(role_v, 4);
  }
  if (!ok) atomicOr(a.flags + FLAG_STUCK, 1);
  if (kInstr && lane == 0) {
    for (int v = 0; v < 2; ++v) {
      for (int k = 0; k < 6; ++k) atomicAdd(a.counters + CTR_PHASE + 6 * v + k, ph_ns[v][k]);
      atomicAdd(a.counters + CTR_ITER + v, ph_it[v]);
    }
  }

  const unsigned nr = __reduce_add_sync(kFullMask, L.nreinject);
  if (lane == 0 && nr) atomicAdd(a.counters + CTR_REINJECT, (unsigned long long)nr);
  if (kInstr) {
    const unsigned nc = __reduce_add_sync(kFullMask, L.ncross);
    const unsigned np = __reduce_add_sync(kFullMask, L.nprobe);
    const unsigned nf = __reduce_add_sync(kFullMask, L.nfast);
    const unsigned long long t_exit = global_ns();
    atomicAdd(a.counters + CTR_LANE_BUSY, it_busy);
    atomicAdd(a.counters + CTR_LANE_IDLE, it_idle);
    if (lane == 0) {
      if (nc) atomicAdd(a.counters + CTR_CROSS, (unsigned long long)nc);
      if (np) atomicAdd(a.counters + CTR_PROBE, (unsigned long long)np);
      if (nf) atomicAdd(a.counters + CTR_FAST, (unsigned long long)nf);
      atomicAdd(a.counters + CTR_WARP_NS, t_exit - t_enter);
      atomicMax(a.counters + CTR_T_FIRST_INV, ~t_enter);
      atomicMax(a.counters + CTR_T_LAST, t_exit);
      atomicAdd(a.counters + CTR_WARPS, 1ULL);
    }
  }
}

// file every exciton under its activity class (first launch after creation or after an upload of the population)
__global__ void __launch_bounds__(256) classify_kernel(const Tables T, const int32_t* site, int64_t P, double dt, ClassLists q) {
  const int64_t  e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int      lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool     mine = e < P;
  const int      cls = mine ? activity_class(load_hop(T.site + site[e]).total * dt) : 0;
  file_excitons(q, mine, cls, (uint32_t)e, lane, lt_mask);
}

// partial[s][j][c] = sum over the j-th slice of excitons of the (step, exciton) records, thread-strided then a
// shared-memory tree: the order is fixed by (P, kStageSplits) alone.
constexpr int kStageSplits = 16;
__global__ void __launch_bounds__(256) reduce_stage_kernel(const StageRec* stage, int64_t P, int nsteps, double* partial) {
  __shared__ double sh[256][4];
  const int         s = blockIdx.x, j = blockIdx.y;
  const int64_t     len = (P + kStageSplits - 1) / kStageSplits;
  const int64_t     q0 = (int64_t)j * len, q1 = (q0 + len < P) ? q0 + len : P;
  const double2*    row = reinterpret_cast<const double2*>(stage + (size_t)s * (size_t)P);
  double            v[4] = {0, 0, 0, 0};
  for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
    const double2 a = __ldcs(row + 2 * q), b = __ldcs(row + 2 * q + 1);
    v[0] += a.x;
    v[1] += a.y;
    v[2] += b.x;
    v[3] += b.y;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) sh[threadIdx.x][c] = v[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int c = 0; c < 4; ++c) sh[threadIdx.x][c] += sh[threadIdx.x + o][c];
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[((size_t)s * kStageSplits + j) * 4 + threadIdx.x] = sh[0][threadIdx.x];
}
__global__ void finish_sums_kernel(const double* partial, int nsteps, double* sums) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsteps * 4) return;
  const int s = k >> 2, c = k & 3;
  double    v = 0.0;
  for (int j = 0; j < kStageSplits; ++j) v += partial[((size_t)s * kStageSplits + j) * 4 + c];
  sums[k] = v;
}

__global__ void iota_kernel(uint32_t* v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

// ---- K3: the hop kernel, contact flavour ---------------------------------------------------------------------------------
// monte_carlo::step + save_population_profile + save_currents + repopulate_contacts (monte_carlo.h:343-355, 525-643,
// 443-491) for nsteps iterations of the loop at main.cpp:98-106.
//
// Excitons never interact, and the contact rules are per-exciton: after the metrics of a step, an exciton whose y lies in
// one of the two contact slabs is deleted, and exactly c1_pop + c2_pop fresh excitons are created on random contact
// sites.  So a launch covers many steps: work item q < P_alive is an exciton alive at the start of the launch, work
// item P_alive + s*C + j is the j-th exciton created after step s (stream id next_gid + s*C + j, first moved in step
// s+1).  Every work item runs until it is deleted or the launch ends; population / current counts are integers and are
// accumulated with integer atomics (shared memory per block, then global), so the result is order-independent.
struct ContactArgs {
  Tables              T;
  ExcitonArrays       S;      // capacity >= P_alive + nsteps * (c1_pop + c2_pop); S.gid is required
  DrawConfig          draws;
  uint8_t*            alive;  // [work items]
  int64_t             P_alive;
  int64_t             c1_pop, c2_pop;
  const int32_t *     c1_sites, *c2_sites;
  int32_t             n_c1, n_c2;
  uint64_t            next_gid;
  double              dt;
  int32_t             nsteps;
  int32_t             n_seg;
  int32_t             use_top;  // try the three widest entries of a row before searching it
  double              ymin, ymax, dy;
  unsigned long long* bins;  // [nsteps][2*n_seg-1]: population per slab, then net crossings per interface
  int32_t*            flags;
  unsigned long long* counters;
};

template <typename Draws, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) contact_kernel(const ContactArgs a) {
  extern __shared__ int hist[];  // [nsteps][2*n_seg-1]
  const int nb = 2 * a.n_seg - 1;
  for (int k = threadIdx.x; k < a.nsteps * nb; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  const int      lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t  C = a.c1_pop + a.c2_pop;
  const int64_t  W = a.P_alive + (int64_t)a.nsteps * C;
  const double   c1_hi = a.ymin + a.dy, c2_lo = a.ymin + double(a.n_seg - 1) * a.dy;  // monte_carlo.h:448-453
  int64_t        q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool           have = q < W;
  Lane           L{};
  Draws          D{};
  Cursor         c{};
  bool           dead = false;
  uint32_t       ev_total = 0;

  auto take = [&]() {
    const uint32_t nc = L.ncross, np = L.nprobe;
    dead = false;
    if (q < a.P_alive) {
      load_lane(L, a.S, a.T, q);
      init_draws(D, a.draws, a.S, q);
      c.step = 0;
    } else {  // born after step s on a contact site: particle ctor via repopulate (monte_carlo.h:477-486)
      const int64_t k = q - a.P_alive, s = k / C, j = k - s * C;
      a.S.gid[q] = a.next_gid + (uint64_t)k;
      init_draws(D, a.draws, a.S, q);
      if (j < a.c1_pop)
        create_exciton(L, a.T, D, a.c1_sites, a.n_c1);
      else
        create_exciton(L, a.T, D, a.c2_sites, a.n_c2);
      c.step = (int32_t)s + 1;
    }
    L.ncross = nc;
    L.nprobe = np;
    begin_step(c, L, a.dt);
  };
  if (have) take();

  while (__any_sync(kFullMask, have)) {
    bool finished = false;
    if (have) {
      if (c.step < a.nsteps && advance_contact(L, a.T, D, c, a.use_top != 0)) {
        int* h = hist + c.step * nb;
        atomicAdd(h + y_slab(L.py, a.ymin, a.dy, a.n_seg), 1);  // monte_carlo.h:566-573
        for (int k = 1; k < a.n_seg; ++k) {                      // monte_carlo.h:626-636
          const int x = interface_crossing(c.oy, L.py, a.ymin + a.dy * double(k));
          if (x) atomicAdd(h + a.n_seg + k - 1, x);
        }
        // repopulate_contacts: everything inside a contact slab is recycled (monte_carlo.h:463-470)
        dead = (L.py >= a.ymin && L.py <= c1_hi) || (L.py >= c2_lo && L.py <= a.ymax);
        ++c.step;
        begin_step(c, L, a.dt);
      }
      finished = dead || (c.step >= a.nsteps) || L.stuck;
    }
    const unsigned fm = __ballot_sync(kFullMask, finished);
    if (fm) {
      if (finished) {
        ev_total += L.nevent;
        materialize(L, a.T);
        store_lane(L, a.S, q);
        a.alive[q] = dead ? 0 : 1;
        if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
        if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
      }
      unsigned long long base = 0;
      const int          leader = __ffs(fm) - 1;
      if (lane == leader) base = atomicAdd(a.counters + CTR_QUEUE, (unsigned long long)__popc(fm));
      base = __shfl_sync(kFullMask, base, leader);
      if (finished) {
        q = (int64_t)base + __popc(fm & lt_mask);
        have = q < W;
        if (have) take();
      }
    }
  }
  const unsigned nc = __reduce_add_sync(kFullMask, L.ncross);
  const unsigned np = __reduce_add_sync(kFullMask, L.nprobe);
  const unsigned ne = __reduce_add_sync(kFullMask, ev_total);
  if (lane == 0) {
    if (nc) atomicAdd(a.counters + CTR_CROSS, (unsigned long long)nc);
    if (np) atomicAdd(a.counters + CTR_PROBE, (unsigned long long)np);
    if (ne) atomicAdd(a.counters + CTR_EVENTS, (unsigned long long)ne);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.nsteps * nb; k += blockDim.x)
    if (hist[k]) atomicAdd(a.bins + k, (unsigned long long)(long long)hist[k]);
}

// initial population of monte_carlo::create_particles (monte_carlo.h:274-316): exciton e belongs to the slab whose
// cumulative count first exceeds e and is created on a random site of that slab's (half-open) site list
struct ContactCreateArgs {
  Tables         T;
  ExcitonArrays  S;
  DrawConfig     draws;
  int64_t        P;
  const int64_t* count_off;  // [n_seg+1] exciton index offsets per slab
  const int64_t* site_off;   // [n_seg+1] offsets into slab_sites
  const int32_t* slab_sites;
  int32_t        n_seg;
  uint8_t*       alive;
  int32_t*       flags;
};
template <typename Draws>
__global__ void __launch_bounds__(256) create_contact_population_kernel(const ContactCreateArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.P) return;
  int slab = 0;
  while (slab + 1 < a.n_seg && e >= a.count_off[slab + 1]) ++slab;
  Lane  L{};
  Draws D{};
  a.S.gid[e] = a.draws.first_gid + (uint64_t)e;
  init_draws(D, a.draws, a.S, e);
  create_exciton(L, a.T, D, a.slab_sites + a.site_off[slab], (int32_t)(a.site_off[slab + 1] - a.site_off[slab]));
  store_lane(L, a.S, e);
  a.S.last_events[e] = 0;
  a.alive[e] = 1;
  if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
}

// compaction after a contact launch: survivor i of the new list is work item src[i] of the old one
__global__ void __launch_bounds__(256) gather_excitons_kernel(const ExcitonArrays from, const ExcitonArrays to, const uint32_t* src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s = src[i];
  to.px[i] = from.px[s]; to.py[i] = from.py[s]; to.pz[i] = from.pz[s];
  to.dx[i] = from.dx[s]; to.dy[i] = from.dy[s]; to.dz[i] = from.dz[s];
  to.ff[i] = from.ff[s];
  to.site[i] = from.site[s];
  to.heading[i] = from.heading[s];
  to.ndraw[i] = from.ndraw[s];
  to.last_events[i] = from.last_events[s];
  to.gid[i] = from.gid[s];
}

// ---- track_particle (monte_carlo.h:786-818) -------------------------------------------------------------------------------
// One exciton born on the first contact (repopulate with n_particle = 1, monte_carlo.h:797-799) is stepped until it
// enters the last slab; its position after every step is the row the reference writes to particle_path.N.dat.  A
// diagnostic: one thread, bounded by max_steps (the reference's loop is unbounded).
struct TrackArgs {
  Tables         T;
  DrawConfig     draws;  // Philox: (seed, first_gid) is the stream; replay: draws [replay_off[0], replay_off[1])
  const int32_t* c1_sites;
  int32_t        n_c1;
  double         dt, y_stop;
  int64_t        max_steps;
  double*        path;   // [max_steps][3]
  int64_t*       n_out;  // steps written, reached (0/1), events
  int32_t*       flags;
};
template <typename Draws>
__device__ __forceinline__ void init_track_draws(Draws& D, const DrawConfig& dc);
template <>
__device__ __forceinline__ void init_track_draws<PhiloxDraws>(PhiloxDraws& D, const DrawConfig& dc) {
  D.init(dc.seed, dc.first_gid);
}
template <>
__device__ __forceinline__ void init_track_draws<ReplayDraws>(ReplayDraws& D, const DrawConfig& dc) {
  D.init(dc.replay_draws, dc.replay_logs, dc.replay_off[0], dc.replay_off[1]);
}
template <typename Draws>
__global__ void track_kernel(const TrackArgs a) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Lane  L{};
  Draws D{};
  init_track_draws(D, a.draws);
  create_exciton(L, a.T, D, a.c1_sites, a.n_c1);
  int64_t n = 0;
  while (L.py < a.y_stop && n < a.max_steps && !L.stuck && !D.exhausted()) {  // monte_carlo.h:811
    Cursor c{};
    begin_step(c, L, a.dt);
    while (!advance_contact(L, a.T, D, c) && !L.stuck) {
    }
    a.path[3 * n + 0] = L.px;
    a.path[3 * n + 1] = L.py;
    a.path[3 * n + 2] = L.pz;
    ++n;
  }
  a.n_out[0] = n;
  a.n_out[1] = (L.py < a.y_stop) ? 0 : 1;
  a.n_out[2] = (int64_t)L.nevent;
  if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
  if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
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
  RowEntry*       row;        // [nnz]
  SiteRec*        site;       // rate fields of every record: total, 1/total, CSR row
  TopRec*         top;        // [N] the three widest entries of every row
  int32_t*        flags;
  unsigned long long* counters;
};

// Enumerate the candidates of site i in the reference's order: the 27-cell stencil with x outermost and z innermost
// (monte_carlo.h:402-411), sites of a cell in ascending list order (monte_carlo.h:389-395).
//
// Only about one candidate in six lies inside the cutoff sphere, and the work per accepted pair (acos, four nearest-grid
// searches) is ~15x the distance test.  The loop is therefore split in two: every lane first advances its own cursor
// to its next accepted candidate (cheap, divergent), then the lanes of the warp evaluate their pairs together.
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
  TopEntries top;
  top.clear();
  if (kFill) base = a.row_begin[i];
  int      c = -1;  // ordinal of the current stencil cell: ix = cx-1 + c/9, iy = cy-1 + (c/3)%3, iz = cz-1 + c%3
  int64_t  q = 0, q1 = 0;
  for (;;) {
    bool     found = false;
    SiteGeom s2{};
    int64_t  qf = 0;
    while (!found) {
      if (q >= q1) {  // next non-empty cell of the stencil, in the reference's order
        bool more = false;
        while (++c < 27) {
          const int ix = cx - 1 + c / 9, iy = cy - 1 + (c / 3) % 3, iz = cz - 1 + c % 3;
          if (!(ix > -1 && ix < a.nb[0] && iy > -1 && iy < a.nb[1] && iz > -1 && iz < a.nb[2])) continue;
          const int64_t b = (int64_t)ix + (int64_t)iy * a.nb[0] + (int64_t)iz * a.nb[0] * a.nb[1];
          q = a.cell_start[b];
          q1 = a.cell_start[b + 1];
          if (q < q1) {
            more = true;
            break;
          }
        }
        if (!more) break;
      }
      s2 = a.cell_geom[q];
      qf = q++;
      found = within_cutoff(s1, s2, a.radius);
    }
    if (!found) break;
    if (kFill) {
      const double rate = pair_rate(s1, s2, a.R, &guard);
      const double below = (d == 0) ? -1.0 : acc;
      acc = (d == 0) ? rate : acc + rate;  // scatterer.cpp:78-80, sequential
      RowEntry en;
      en.cum = acc;
      en.nbr = a.cell_sites[qf];
      top.add(rate, below, acc, en.nbr);
      en.pad = 0;
      a.row[base + d] = en;
    }
    ++d;
  }
  if (!kFill) {
    a.deg[i] = d;
  } else {
    a.site[i].total = acc;                  // scatterer.h:91  _max_rate = neighbors.back().first
    a.site[i].inv_total = d ? 1. / acc : 0.0;  // scatterer.h:92
    a.site[i].row_begin = (uint32_t)base;
    a.site[i].row_len = d;
    uint8_t g8[kGuideBuckets];
    struct CumView {
      const RowEntry* r;
      __device__ double operator[](uint32_t k) const { return r[k].cum; }
    };
    build_guide(CumView{a.row + base}, d, acc, g8);  // reads back this thread's own row
#pragma unroll
    for (int j = 0; j < kGuideBuckets; ++j) a.site[i].guide[j] = g8[j];
    top.store(a.top[i]);
    if (d == 0) atomicOr(a.flags + FLAG_EMPTY_ROW, 1);
    if (guard) atomicAdd(a.counters + CTR_GUARD, 1ULL);
  }
}

}  // namespace cntmc
