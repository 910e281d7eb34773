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
  uint64_t* gid;           // stream id where excitons are created and destroyed (contact mode); null = first_gid + index
};

struct DrawConfig {
  uint64_t       seed;
  uint64_t       first_gid;
  const int64_t* replay_off;  // non-null selects replay
  const int32_t* replay_draws;
  const double*  replay_logs;
};

__device__ __forceinline__ void load_lane(Lane& L, const ExcitonArrays& S, const Tables& T, int64_t e) {
  L.px = __ldcs(S.px + e);  // exciton state streams through once per launch
  L.py = __ldcs(S.py + e);
  L.pz = __ldcs(S.pz + e);
  L.dx = __ldcs(S.dx + e);
  L.dy = __ldcs(S.dy + e);
  L.dz = __ldcs(S.dz + e);
  L.ff = __ldcs(S.ff + e);
  L.site = __ldcs(S.site + e);
  L.heading_right = __ldcs(S.heading + e) != 0;
  L.ndraw = __ldcs(S.ndraw + e);
  L.nevent = 0;
  L.stuck = false;
  attach_site(L, T);
}
// Loads that are served by L2 and never by a (possibly stale) L1 line: for state another kernel wrote while this one was
// already running (overlap mode of the trap kernel, see hop_loop).  LDG.E.STRONG.GPU; no fence, no L1 invalidation.
__device__ __forceinline__ double ld_strong(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_strong(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t ld_strong(const int32_t* p) { return (int32_t)ld_strong(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint32_t ld_strong(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {  // MEMBAR.ALL.GPU + STG.STRONG: everything this thread stored before is visible first
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void load_lane_strong(Lane& L, const ExcitonArrays& S, const Tables& T, int64_t e) {
  L.px = ld_strong(S.px + e);
  L.py = ld_strong(S.py + e);
  L.pz = ld_strong(S.pz + e);
  L.dx = ld_strong(S.dx + e);
  L.dy = ld_strong(S.dy + e);
  L.dz = ld_strong(S.dz + e);
  L.ff = ld_strong(S.ff + e);
  L.site = ld_strong(S.site + e);
  L.heading_right = ld_strong(S.heading + e) != 0;
  L.ndraw = ld_strong(S.ndraw + e);
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

enum { FLAG_STUCK = 0, FLAG_REPLAY = 1, FLAG_EMPTY_ROW = 2, FLAG_BAD_LINKS = 3, FLAG_COUNT = 4 };
enum {
  CTR_REINJECT = 0, CTR_GUARD = 1, CTR_CROSS = 2, CTR_PROBE = 3, CTR_QUEUE = 4, CTR_EVENTS = 5, CTR_FAST = 6,
  CTR_WALK = 7,  // instrumented trap solver: events decided by the window walk
  // instrumented hop kernel only: warp residency (ns summed over warps), kernel span as seen from the device
  // (~first start and last exit, stored for atomicMax), number of warps, lane-iterations with / without an exciton
  CTR_WARP_NS = 8, CTR_T_FIRST_INV = 9, CTR_T_LAST = 10, CTR_WARPS = 11, CTR_LANE_BUSY = 12, CTR_LANE_IDLE = 13,
  CTR_COUNT = 16
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
  if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
}

// ---- K2: the hop kernel, Green-Kubo flavour -------------------------------------------------------------------------------
// Activity classes.  Exciton activity is extremely skewed and predictable from the state: Gamma(site) * dt = expected
// events per step if the exciton stays where it is.  On the C2 film 83 % of the sites have Gamma*dt < 0.125; 17 % sit
// in "traps" -- two to five neighbouring sites of one tube whose mutual rate is the table's maximum (8.78e13/s, so
// Gamma*dt = 8.8 per partner): an exciton there scatters 9 to 38 times per step until a flight carries it out.  1 % of
// the excitons make a third of all events (deep traps, two or more partners), another 7 % a second third.
// When a lane stores an exciton it files it under one of five classes; in the next launch "hot" blocks serve the
// active classes and "cold" blocks the quiet ones, so that warps mostly hold excitons that run the same branch of the
// loop, and class 4 (the deep traps) goes to the blocks that run the group solver (see hop_loop).
constexpr int kClasses = 5;
constexpr int kLists = kClasses + 2;  // + the excitons handed over during the launch: lanes -> trap solver, and back
constexpr int kDeferred = kClasses, kReturned = kClasses + 1;
__device__ __forceinline__ int activity_class(double ev, double deep_thr) {
  return ev >= deep_thr ? 4 : ev >= 8.0 ? 3 : ev >= 1.0 ? 2 : ev >= 0.125 ? 1 : 0;
}
struct ClassLists {
  const uint32_t*     list[kLists];           // excitons of each class, filed by the previous launch; [kDeferred], [kReturned]: this launch
  const uint32_t*     count;                  // [kClasses]
  const uint32_t*     hand_count;             // [2] sizes of list[kDeferred], list[kReturned]
  unsigned long long* head;                   // [kLists] next unassigned position of each list
  uint32_t*           next_list[kClasses];    // being filled for the next launch
  uint32_t*           next_count;             // [kClasses]
  uint32_t*           hand_list[2];           // being filled in this launch: [0] by kubo_kernel (deferred), [1] by deep_kernel (returned)
  uint32_t*           hand_to;                // [2] their sizes
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
// hand excitons over to the other kernel: which = 0 lanes -> trap solver (deferred), 1 trap solver -> lanes (returned)
// publish: the receiver is running already (overlap mode) and waits for its slot of the list to become non-zero; the entry is
// e + 1, stored with release semantics so that the exciton's state and cursor, stored by this thread before, are visible first
__device__ __forceinline__ void hand_over(const ClassLists& q, int which, bool mine, uint32_t e, int lane, unsigned lt_mask,
                                          bool publish = false) {
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (m) {
    uint32_t  base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(q.hand_to + which, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
      uint32_t* slot = q.hand_list[which] + (base + __popc(m & lt_mask));
      if (publish)
        st_release(slot, e + 1u);
      else
        *slot = e;
    }
  }
}
// hand an exciton to every lane that `want`s one, from the n_serve lists named in `serve`, in that order (call with the
// whole warp).  Returns false for lanes left without work.
// serve: list numbers, four bits each, first one in the low bits
__device__ __forceinline__ bool take_exciton(const ClassLists& q, bool want, uint32_t serve, int n_serve, int lane, unsigned lt_mask,
                                             int64_t& e, int& from) {
  bool     got = false;
  unsigned need = __ballot_sync(0xffffffffu, want);
  for (int k = 0; k < n_serve && need; ++k) {
    const int          c = (int)((serve >> (4 * k)) & 15u);
    const int          n = __popc(need), leader = __ffs(need) - 1;
    const long long    cnt = (long long)(c >= kClasses ? q.hand_count[c - kClasses] : q.count[c]);
    unsigned long long base = 0;
    if (lane == leader) base = (cnt > 0) ? atomicAdd(q.head + c, (unsigned long long)n) : (unsigned long long)cnt;
    base = __shfl_sync(0xffffffffu, base, leader);
    const long long avail = cnt - (long long)base;
    if (want && !got) {
      const int rank = __popc(need & lt_mask);
      if ((long long)rank < avail) {
        e = (int64_t)q.list[c][base + (unsigned long long)rank];
        from = c;
        got = true;
      }
    }
    need = __ballot_sync(0xffffffffu, want && !got);
  }
  return got;
}

#if defined(CNTMC_PROFILE_SEGMENTS)
constexpr int kWarpTimeCols = 12;  // + cycles per loop segment of lane 0
#else
constexpr int kWarpTimeCols = 4;
#endif
struct alignas(32) StageRec {  // one full 32-byte sector per (step, exciton)
  double dx2, dy2, dz2, events;
};
// where an exciton stood when a lane block handed it to the group solver in the middle of a launch
struct CursorArrays {
  double *  dt_rem, *ox, *oy, *oz;  // time left in the current step, particle::_old_pos of the step
  int32_t*  step;
  uint32_t* nevent;                 // events of the current step so far
};

struct KuboArgs {
  Tables              T;
  ExcitonArrays       S;
  CursorArrays        C;
  DrawConfig          draws;
  ClassLists          q;
  int32_t             round;        // 1: the lists of the launch; 2: what the other kernel handed back during round 1
  int32_t             yield_on;     // deep_kernel: an exciton that ends a time step outside a trap goes back to the lanes
  int32_t             hot_blocks;   // blocks [0, hot_blocks) serve the active classes first
  int64_t             n_sites;
  int32_t             top_entries;  // try the three widest entries of a row before searching it
  int32_t             burst;        // trap-lanes kernel: events a lane may run in a row inside one iteration of its warp (>= 1)
  int32_t             overlap;      // the lane kernel and the trap kernel of a launch run side by side (see hop_loop)
  int32_t             lane_grid;    // overlap: blocks of the lane kernel
  uint32_t*           sync;         // overlap: [0] = blocks of the lane kernel that have finished
  double              deep_thr;     // Gamma*dt from which an exciton belongs to the group solver (inf: never)
  double              deep_rate;    // the same as a rate: deep_thr / dt
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

// Draw source of the group solver: G lanes share one exciton, and lane j of the group prepares the two draws of the j-th
// event from now -- Philox2x32-10 and the logarithm of the free-flight draw, the two longest dependency chains of an
// event -- while its neighbours prepare the others.  The stream is counter-based, so this is the same sequence the
// one-lane source produces, fetched with a shuffle instead of computed in line; anything that does not fit the pattern
// (a re-injection draw, a zero draw) is served from the same batch or computed directly.
template <int G>
struct GroupDraws {
  PhiloxDraws base;
  uint32_t    nd0, r1, r2;
  double      lg;
  unsigned    gmask;
  int         gbase, j;
  bool        valid;
  __device__ __forceinline__ void init(uint64_t seed, uint64_t gid) {
    base.init(seed, gid);
    const int lane = threadIdx.x & 31;
    gbase = lane & ~(G - 1);
    j = lane - gbase;
    gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    valid = false;
    nd0 = 0;
  }
  __device__ __forceinline__ void refill(uint32_t ndraw) {
    nd0 = ndraw;
    uint32_t    n = ndraw + 2u * (uint32_t)j;
    PhiloxDraws t = base;
    t.have = false;
    r1 = (uint32_t)t.next(n);
    r2 = (uint32_t)t.next(n);
    lg = r2 ? fast_log_unit(div_by((double)r2, kRandMax, kInvRandMax)) : 0.0;
    valid = true;
  }
  __device__ __forceinline__ int32_t next(uint32_t& ndraw) {
    uint32_t off = ndraw - nd0;
    if (!valid || off >= 2u * G) {
      refill(ndraw);
      off = 0;
    }
    const uint32_t v = __shfl_sync(gmask, (off & 1u) ? r2 : r1, gbase + (int)(off >> 1));
    ++ndraw;
    return (int32_t)v;
  }
  __device__ __forceinline__ double log_ratio(int32_t r, uint32_t ndraw_after) const {
    const uint32_t off = ndraw_after - 1u - nd0;
    if (valid && off < 2u * G && (off & 1u)) return __shfl_sync(gmask, lg, gbase + (int)(off >> 1));
    return fast_log_unit(div_by((double)r, kRandMax, kInvRandMax));
  }
  __device__ __forceinline__ bool exhausted() const { return false; }
};
template <int G>
__device__ __forceinline__ void init_draws(GroupDraws<G>& D, const DrawConfig& dc, const ExcitonArrays& S, int64_t e) {
  D.init(dc.seed, S.gid ? S.gid[e] : dc.first_gid + (uint64_t)e);
}
template <typename Draws, int G>
struct LoopDraws {
  typedef GroupDraws<G> type;
};
template <typename Draws>
struct LoopDraws<Draws, 1> {
  typedef Draws type;
};

// The loop of the hop kernels.  Persistent warps; every group of G lanes owns one exciton at a time, carries it through
// the time steps of the launch, files it under its activity class for the next launch and takes another one.
//
// G = 1 (kubo_kernel): one exciton per lane.  The loop is flat: an iteration moves every busy lane forward by the end of
// one time step and / or one scattering event; lanes of a warp are in general in different time steps of different
// excitons.  Nothing in the loop needs a barrier or a floating-point atomic: when a lane ends a step it writes its
// squared displacement to the (step, exciton) record, and reduce_stage_kernel sums the records in a fixed order
// afterwards, so the ensemble sums do not depend on which lane processed which exciton, nor on any tuning option.
//   Deferral: an exciton that lands in a deep trap (Gamma*dt >= deep_thr) would occupy its lane for thousands of
//   sequential events while the other 31 idle at the end of the launch -- at 1e6 excitons that chain IS the launch
//   time.  The lane stores it with its cursor (step, time left, events of the step, start-of-step position) and
//   deep_kernel finishes its launch.
//
// G = 8 (deep_kernel, the trap solver): the G lanes of a group hold the same exciton and execute the same instructions
// on the same values, and they split what does not depend on the event chain:
//   * lane j prepares the two draws of the j-th event from now and the logarithm of its free-flight draw (GroupDraws);
//   * lane j keeps the record of site b + j of a window of G consecutive sites around the exciton in registers (a trap
//     is two to five neighbouring sites of one tube), and for each of the G prepared dice draws it evaluates what an
//     event on ITS site would do: dice = Gamma_j * r / RAND_MAX against the three widest entries of its row ->
//     destination, as an index into the window (4 bits per draw; 15 = not decided here);
//   * the chain itself is then walked from those registers with shuffles: flight time of the segment ahead against the
//     free-flight time (owner's q), the owner's 4-bit answer for this draw, the destination's 1/Gamma times the
//     prepared logarithm.  About 90 cycles and 25 instructions per event instead of ~2000 and 200.
// Whatever the walk cannot decide -- a flight that reaches the next site, a dice outside the top entries, a destination
// outside the window, the end of a time step -- leaves the exciton exactly where the generic path (the same code as
// G = 1, run redundantly by the G lanes) picks it up.  Same draws, same arithmetic, same order: the same bits.
// Only the first lane of a group stores anything.
//
// kInstr adds what only tests and the roofline bookkeeping need (site traces, probe / crossing counters).
template <typename Draws, bool kInstr, int G, bool kDefer, bool kTrap>
__device__ __forceinline__ void hop_loop(const KuboArgs& a, double (&s_delta)[3][128], double (&s_old)[3][128]) {
  typedef typename LoopDraws<Draws, G>::type DrawsT;
  const int      tid = threadIdx.x, lane = threadIdx.x & 31;
  const int      gbase = lane & ~(G - 1);
  const bool     leader = (lane == gbase);
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const bool     hot_role = (int)blockIdx.x < a.hot_blocks;
  constexpr bool deep_on = kDefer;  // compile-time: even switched off at run time, the hand-over code cost 2.5 % (profiles/round2_*)
  uint32_t       e = 0;
  Lane           L{};
  DrawsT         D{};
  double         dt_rem = 0.0;
  int32_t        step = 0;
  int32_t*       trace = nullptr;
  int32_t        trace_base = 0;  // events already in the exciton's trace when the current time step began
  int            from = 0;        // the list the current exciton came from
  uint32_t       n_walk = 0;      // (instrumented) events decided by the window walk
  unsigned long long t_enter = 0, it_busy = 0, it_idle = 0, t_dry = 0;
  int                iter = 0;
  if (kInstr) t_enter = global_ns();
#if defined(CNTMC_PROFILE_SEGMENTS)
  L.seg_t = clock64();
#endif
  // trap solver: this lane's site of the window
  int32_t  win_b = -1;
  double   w_total = 0, w_inv = 0, w_qr = 0, w_ql = 0;
  uint32_t w_iv0 = 0, w_iv1 = 0, w_iv2 = 0;  // top entries as intervals of the draw (hop_core.h SiteRec)
  int32_t  w_n0 = -1, w_n1 = -1, w_n2 = -1, w_left = -1, w_right = -1;
  uint32_t w_flags = 0;

  auto start = [&]() {
    const uint32_t nc = L.ncross, np = L.nprobe, nr = L.nreinject, nf = L.nfast;
    const bool     fresh = kTrap && G == 1 && a.overlap != 0;  // the state may have been written while this kernel was running
    if (fresh)
      load_lane_strong(L, a.S, a.T, (int64_t)e);
    else
      load_lane(L, a.S, a.T, (int64_t)e);
    L.ncross = nc;
    L.nprobe = np;
    L.nreinject = nr;
    L.nfast = nf;
    init_draws(D, a.draws, a.S, (int64_t)e);
    step = 0;
    dt_rem = a.dt;
    s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
    s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;  // _old_pos = _pos (particle.cpp:59)
    if (kTrap && from == kDeferred) {  // handed over in the middle of a time step
      if (fresh) {
        step = ld_strong(a.C.step + e);
        dt_rem = ld_strong(a.C.dt_rem + e);
        L.nevent = ld_strong(a.C.nevent + e);
        s_old[0][tid] = ld_strong(a.C.ox + e); s_old[1][tid] = ld_strong(a.C.oy + e); s_old[2][tid] = ld_strong(a.C.oz + e);
      } else {
        step = a.C.step[e];
        dt_rem = a.C.dt_rem[e];
        L.nevent = a.C.nevent[e];
        s_old[0][tid] = a.C.ox[e]; s_old[1][tid] = a.C.oy[e]; s_old[2][tid] = a.C.oz[e];
      }
    }
    if (!kTrap && from == kReturned) step = a.C.step[e];  // handed back at a step boundary
    if (kInstr && a.trace_sites) {  // the trace continues where the previous launch stopped
      trace_base = a.trace_counts[e];  // may exceed the capacity: events beyond it are counted, not recorded
      trace = (leader && trace_base < a.trace_cap) ? a.trace_sites + (int64_t)e * a.trace_cap + trace_base : nullptr;
    }
  };
  // A warp keeps to one role so that its lanes run the same branch of the loop most of the time (hot: scattering
  // events, cold: chain walks and step ends); it changes role only once, when all its lanes have run dry.
  // The lists a warp serves in its first and in its second role, four bits per list number:
  const uint32_t hot_lists = deep_on ? 0x23u : 0x234u, cold_lists = 0x10u;
  const int      n_hot = deep_on ? 2 : 3, n_roles = kTrap ? 1 : 2;
  for (int role = 0; role < n_roles; ++role) {
    const bool     serve_hot = (role == 0) == hot_role;
    // round 2 of a launch serves what the other kernel handed over during round 1
    const uint32_t serve = kTrap ? (a.round == 2 ? (uint32_t)kDeferred : (4u | ((uint32_t)kDeferred << 4)))
                                 : a.round == 2 ? (uint32_t)kReturned : serve_hot ? hot_lists : cold_lists;
    // Overlap mode (trap kernel, one exciton per lane): the lane kernel of this launch runs at the same time and is still
    // filling the deferred list.  A lane that finds class 4 exhausted claims the next SLOT of that list -- claimed or not,
    // slots fill in order -- and polls it until the lane kernel publishes an exciton there (hand_over) or has finished
    // without reaching it.
    const bool     overlap = kTrap && G == 1 && a.overlap != 0;
    const int      n_serve = kTrap ? ((a.round == 2 || overlap) ? 1 : 2) : a.round == 2 ? 1 : serve_hot ? n_hot : 2;
    bool           polling = false;
    uint32_t       slot = 0, poll_it = 0;
    auto claim_slot = [&](bool want) {  // call with the whole warp
      const unsigned m = __ballot_sync(kFullMask, want);
      if (m) {
        unsigned long long base = 0;
        const int          first = __ffs(m) - 1;
        if (lane == first) base = atomicAdd(a.q.head + kDeferred, (unsigned long long)__popc(m));
        base = __shfl_sync(kFullMask, base, first);
        if (want) {
          const unsigned long long mine = base + (unsigned long long)__popc(m & lt_mask);
          slot = (uint32_t)mine;
          polling = mine < (unsigned long long)a.P;  // an exciton is deferred at most once per launch: later slots stay empty
        }
      }
    };
    int64_t        e64 = 0;
    bool           have = take_exciton(a.q, leader, serve, n_serve, lane, lt_mask, e64, from);
    if (overlap) claim_slot(!have);
    if (G > 1) {
      have = __shfl_sync(kFullMask, have ? 1 : 0, gbase) != 0;
      e64 = __shfl_sync(kFullMask, e64, gbase);
      from = __shfl_sync(kFullMask, from, gbase);
    }
    if (have) {
      e = (uint32_t)e64;
      start();
    }
    while (__any_sync(kFullMask, have || polling)) {
      if (overlap) {
        if (polling && (poll_it++ & 3u) == 0u) {
          const uint32_t* entry = a.q.list[kDeferred] + slot;
          uint32_t        v = ld_strong(entry);
          // the lane kernel has finished: the list is final.  (The relaxed read comes first: an acquire load is followed by an
          // invalidation of the SM's whole L1, CCTL.IVALL, which the busy lanes of this SM would pay for at every empty poll.)
          if (v == 0u && ld_strong(a.sync) >= (uint32_t)a.lane_grid && ld_acquire(a.sync) >= (uint32_t)a.lane_grid) {
            v = ld_strong(entry);
            if (v == 0u) polling = false;
          }
          if (v != 0u) {
            e = v - 1u;
            from = kDeferred;
            polling = false;
            have = true;
            start();
          }
        }
        if (!__any_sync(kFullMask, have)) {
          __nanosleep(400);
          continue;
        }
      }
      bool finished = false, did_event = false, did_step = false;
      if (kInstr) {
        if (have) ++it_busy; else ++it_idle;
        ++iter;
      }
      bool walk_finished = false, walk_yield = false;
      if constexpr (G > 1) {
        // ---- trap solver: up to G events, and the step ends between them, from the window registers
        const bool can = have && !L.stuck && a.top_entries != 0 && a.T.dir != nullptr && a.n_sites >= G &&
                         ((L.ff <= dt_rem) || L.at_site);
        if (can) {
          if (win_b < 0 || L.site < win_b || L.site >= win_b + G) {  // (re)centre the window: lane j takes site b + j
            int64_t b = (int64_t)L.site - (G / 2 - 1);
            b = b < 0 ? 0 : (b > a.n_sites - G ? a.n_sites - G : b);
            win_b = (int32_t)b;
            const SiteRec*  rec = a.T.site + (win_b + (lane - gbase));
            const SiteChain c = load_chain(rec);
            const TopLoaded t = load_top(&rec->top);
            w_left = c.left; w_right = c.right; w_qr = c.q_right; w_ql = c.q_left;
            w_total = t.total; w_inv = c.inv_total;
            w_iv0 = t.iv0; w_iv1 = t.iv1; w_iv2 = t.iv2;
            w_n0 = t.nbr0; w_n1 = t.nbr1; w_n2 = t.nbr2;
            // what particle::fly does with the heading on this site (particle.cpp:20-34): bit 0 = heading after a start
            // to the right, bit 1 = after a start to the left, bit 2 = link-less or odd chain (generic path)
            w_flags = ((w_right > -1) ? 1u : 0u) | ((w_left > -1) ? 0u : 2u) | ((w_left < 0 && w_right < 0) ? 4u : 0u) |
                      ((w_left == w_right) ? 4u : 0u);
          }
          uint32_t off = L.ndraw - D.nd0;
          if (!D.valid || off >= 2u * G || (off & 1u)) {  // the prepared draws must start at an event
            D.refill(L.ndraw);
            off = 0;
          }
          // every lane: what would an event on MY site do with each of the prepared dice draws?
          uint32_t      tbl = 0;
          const int32_t my_site = win_b + (lane - gbase);
#pragma unroll
          for (int k = 0; k < G; ++k) {
            const uint32_t u = __shfl_sync(gmask, D.r1, gbase + k) >> kTopBlockShift;  // scatterer.cpp:17 through the draw intervals
            const bool    in0 = in_draw_blocks(w_iv0, u), in1 = in_draw_blocks(w_iv1, u), in2 = in_draw_blocks(w_iv2, u);
            const int32_t dest = in0 ? w_n0 : in1 ? w_n1 : w_n2;
            const int32_t di = dest - win_b;
            const bool    ok = (in0 || in1 || in2) && w_n0 != kEmptyRow && dest != my_site && di >= 0 && di < G;
            tbl |= (ok ? (uint32_t)di : 15u) << (4 * k);
          }
          const unsigned zero2 = __ballot_sync(gmask, D.r2 == 0u) >> gbase;  // a zero free-flight draw is redrawn (scatterer.h:76-78)
          int    s_idx = L.site - win_b, k = (int)(off >> 1);
          bool   heading = L.heading_right, at = L.at_site, moved = false;
          double ff = L.ff, px = L.px, py = L.py, pz = L.pz;  // the position counts only while `at` is false
          for (int guard = 0; guard < 4 * G; ++guard) {
            const uint32_t fl = __shfl_sync(gmask, w_flags, gbase + s_idx);
            const double   qr = __shfl_sync(gmask, w_qr, gbase + s_idx), ql = __shfl_sync(gmask, w_ql, gbase + s_idx);
            const bool     nh = heading ? (fl & 1u) != 0 : (fl & 2u) != 0;  // heading after the start of the flight
            if (fl & 4u) break;  // link-less or odd chain: generic path
            if (!(ff <= dt_rem)) {
              // -- the flight outlasts the time step: particle::step's fly(dt) tail and monte_carlo::kubo_step's loop body
              if (!at) break;                          // (a step end in mid-segment: generic path)
              const double q = nh ? qr : ql;
              if (q < dt_rem) break;                   // reaches the next site within the step: generic path
              const int32_t  site = win_b + s_idx;
              const SitePos  p0 = load_pos(a.T.pos + site);
              const Quad     u = load_quad(reinterpret_cast<const char*>(a.T.dir + site) + (nh ? 0 : 32));
              const double   kk = a.T.velocity * dt_rem;  // particle.cpp:48
              const double   nx = p0.x + u.a * kk, ny = p0.y + u.b * kk, nz = p0.z + u.c * kk;
              if (nx < a.T.rem_lo[0] || ny < a.T.rem_lo[1] || nz < a.T.rem_lo[2] || a.T.rem_hi[0] < nx || a.T.rem_hi[1] < ny ||
                  a.T.rem_hi[2] < nz)
                break;                                 // leaves the removal box: re-injection on the generic path
              heading = nh;
              ff -= dt_rem;                            // particle.cpp:79
              const double dx = s_delta[0][tid] + (nx - s_old[0][tid]), dy = s_delta[1][tid] + (ny - s_old[1][tid]),
                           dz = s_delta[2][tid] + (nz - s_old[2][tid]);  // particle.h:97
              if (leader) {
                double2* rec = reinterpret_cast<double2*>(a.stage + ((size_t)step * (size_t)a.P + (size_t)e));
                __stcs(rec, make_double2(dx * dx, dy * dy));
                __stcs(rec + 1, make_double2(dz * dz, (double)L.nevent));
              }
              s_delta[0][tid] = dx; s_delta[1][tid] = dy; s_delta[2][tid] = dz;
              s_old[0][tid] = nx; s_old[1][tid] = ny; s_old[2][tid] = nz;
              px = nx; py = ny; pz = nz;
              at = false;
              ++step;
              dt_rem = a.dt;
              if (kInstr) {
                trace_base += (int32_t)L.nevent;
                if (trace) trace = (trace_base < a.trace_cap) ? trace + L.nevent : nullptr;
              }
              L.nevent = 0;
              moved = true;
              if (step >= a.nsteps) {
                walk_finished = true;
                break;
              }
              if (a.yield_on && __shfl_sync(gmask, w_total, gbase + s_idx) < a.deep_rate) {  // not a trap: back to the lanes
                walk_yield = true;
                break;
              }
              continue;
            }
            // -- a scattering event with the k-th prepared pair of draws
            if (k >= G) break;
            double q;
            if (at) {
              q = nh ? qr : ql;
            } else {  // the flight starts between two sites: particle.cpp:39-41 on the actual position
              const int32_t next = __shfl_sync(gmask, nh ? w_right : w_left, gbase + s_idx);
              const SitePos n = load_pos(a.T.pos + next);
              q = div_by(norm3(px - n.x, py - n.y, pz - n.z), a.T.velocity, a.T.inv_velocity);
            }
            if (q < ff) break;  // the flight reaches the next site: generic path
            const uint32_t tb = __shfl_sync(gmask, tbl, gbase + s_idx);
            const uint32_t di = (tb >> (4 * k)) & 15u;
            if (di == 15u || ((zero2 >> k) & 1u)) break;
            // the flight stops on the way (its leg is never seen), the exciton hops to window site di
            heading = nh;
            dt_rem -= ff;  // particle.cpp:63
            const double inv = __shfl_sync(gmask, w_inv, gbase + (int)di), lgk = __shfl_sync(gmask, D.lg, gbase + k);
            ff = -inv * lgk;  // scatterer.h:79
            s_idx = (int)di;
            at = true;
            const uint32_t room = (kInstr && trace) ? (uint32_t)(a.trace_cap - trace_base) : 0u;
            if (kInstr && trace != nullptr && L.nevent < room) trace[L.nevent] = win_b + s_idx;
            ++L.nevent;
            L.ndraw += 2u;
            L.nprobe += 2u;
            ++L.nfast;
            if (kInstr) ++n_walk;
            ++k;
            moved = true;
          }
          if (moved) {
            L.site = win_b + s_idx;
            L.heading_right = heading;
            L.ff = ff;
            L.left = __shfl_sync(gmask, w_left, gbase + s_idx);
            L.right = __shfl_sync(gmask, w_right, gbase + s_idx);
            L.q_right = __shfl_sync(gmask, w_qr, gbase + s_idx);
            L.q_left = __shfl_sync(gmask, w_ql, gbase + s_idx);
            L.at_site = at;
            L.pos_valid = !at;
            if (!at) {
              L.px = px; L.py = py; L.pz = pz;
            }
          }
        }
      }
      // One iteration = up to two operations per lane: first the end of a time step for the lanes whose free flight
      // outlasts the step, then a scattering event for the lanes whose flight ends inside the (possibly new) step.
      const bool need_s = have && !walk_finished && !walk_yield && !(L.ff <= dt_rem);  // particle.cpp:62 false: the flight outlasts the step
      CNTMC_SEG(L, 6);  // loop head, refill
      if (need_s) {
        const double t = dt_rem;
        const Leg    leg = fly(L, a.T, t, true);
        L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
        after_flight_step_end(L, a.T, D, leg, t, s_old[0][tid], s_old[1][tid], s_old[2][tid]);
        if (!kInstr && L.nreinject) {  // counted at once (re-injections are rare): one register less across the loop
          if (leader) atomicAdd(a.counters + CTR_REINJECT, 1ULL);
          L.nreinject = 0;
        }
        if (leader) {
          // streaming stores: written once, read once by the reduction, must not evict the tables from L2
          double2* rec = reinterpret_cast<double2*>(a.stage + ((size_t)step * (size_t)a.P + (size_t)e));
          __stcs(rec, make_double2(L.dx * L.dx, L.dy * L.dy));  // std::pow(delta_pos, 2), monte_carlo.cpp:397-399
          __stcs(rec + 1, make_double2(L.dz * L.dz, (double)L.nevent));  // events of this time step
        }
        s_delta[0][tid] = L.dx; s_delta[1][tid] = L.dy; s_delta[2][tid] = L.dz;
        s_old[0][tid] = L.px; s_old[1][tid] = L.py; s_old[2][tid] = L.pz;
        ++step;
        dt_rem = a.dt;
        if (kInstr) {
          trace_base += (int32_t)L.nevent;
          if (trace) trace = (trace_base < a.trace_cap) ? trace + L.nevent : nullptr;
        }
        L.nevent = 0;
        did_step = true;
        finished = (step >= a.nsteps);
      }
      CNTMC_SEG(L, 5);  // step-end path (own, or waiting for the lanes that run it)
      const bool need_e = have && !finished && !walk_finished && !walk_yield && (L.ff <= dt_rem);
      if (need_e) {
        // trap-lanes kernel only (option trap_burst): up to `burst` events of the same time step in a row, without a visit to
        // the warp's step-end path in between.  15 % off the pass over the trapped excitons at 4; on the ordinary lane kernel
        // the same loop LOSES (2: -3 %, 8: -15 %, profiles/round2_trap_solver.txt), so there it is compiled out.
        int left = (kTrap && G == 1) ? a.burst : 1;
        do {
          const double t = L.ff;
          const Leg    leg = fly(L, a.T, t, false);
          dt_rem -= t;  // particle.cpp:63
          const uint32_t room = (kInstr && trace) ? (uint32_t)(a.trace_cap - trace_base) : 0u;
          after_flight_scatter(L, a.T, D, leg, kInstr ? trace : nullptr, room, a.top_entries != 0);
        } while (--left > 0 && (L.ff <= dt_rem) && !L.stuck && !(kDefer && site_total(L, a.T) >= a.deep_rate));
        did_event = true;
      }
      CNTMC_SEG(L, 7);  // waiting for the other lanes of the warp to finish their events
      // trap solver: an exciton that ended a time step on a site that is no trap goes back to the lanes
      bool yield = walk_yield;
      if (kTrap && a.yield_on && did_step && !did_event && !finished && !L.stuck && !(site_total(L, a.T) >= a.deep_rate)) yield = true;
      finished = have && (finished || walk_finished || L.stuck);
      // lanes: landed in a deep trap, the trap solver takes over
      bool defer = false;
      if constexpr (!kTrap && deep_on) defer = did_event && !finished && site_total(L, a.T) >= a.deep_rate;
      const bool release = finished || defer || yield;
      if (__any_sync(kFullMask, release)) {
        int cls = 0;
        if (release) {
          materialize(L, a.T);
          L.dx = s_delta[0][tid]; L.dy = s_delta[1][tid]; L.dz = s_delta[2][tid];
          if (leader) {
            store_lane(L, a.S, (int64_t)e);
            if (yield) a.C.step[e] = step;  // at a step boundary: the step index is the whole cursor
            if (defer) {
              a.C.step[e] = step;
              a.C.dt_rem[e] = dt_rem;
              a.C.nevent[e] = L.nevent;
              a.C.ox[e] = s_old[0][tid]; a.C.oy[e] = s_old[1][tid]; a.C.oz[e] = s_old[2][tid];
            }
            if (kInstr && a.trace_counts) a.trace_counts[e] = trace_base + ((defer || yield) ? 0 : (int32_t)L.nevent);
            if (L.stuck) atomicOr(a.flags + FLAG_STUCK, 1);
            if (D.exhausted()) atomicOr(a.flags + FLAG_REPLAY, 1);
          }
          if (finished) cls = activity_class(site_total(L, a.T) * a.dt, a.deep_thr);
        }
        file_excitons(a.q, finished && leader, cls, e, lane, lt_mask);
        if (!kTrap && deep_on) hand_over(a.q, 0, defer, e, lane, lt_mask, a.overlap != 0);
        if (kTrap && a.yield_on) hand_over(a.q, 1, yield && leader, e, lane, lt_mask);
        bool got = take_exciton(a.q, release && leader, serve, n_serve, lane, lt_mask, e64, from);
        if (G > 1) {
          got = __shfl_sync(kFullMask, got ? 1 : 0, gbase) != 0;
          e64 = __shfl_sync(kFullMask, e64, gbase);
          from = __shfl_sync(kFullMask, from, gbase);
        }
        if (release) {
          have = got;
          if (have) {
            e = (uint32_t)e64;
            start();
          }
        }
        if (overlap) claim_slot(release && !got);
        if (kInstr && t_dry == 0 && __any_sync(kFullMask, release && !got)) t_dry = global_ns();
      }
    }
  }

  if (!leader) L.nreinject = L.ncross = L.nprobe = L.nfast = 0;  // the lanes of a group counted the same things
  const unsigned nr = __reduce_add_sync(kFullMask, L.nreinject);
  if (lane == 0 && nr) atomicAdd(a.counters + CTR_REINJECT, (unsigned long long)nr);
  if (kInstr) {
    const unsigned nc = __reduce_add_sync(kFullMask, L.ncross);
    const unsigned np = __reduce_add_sync(kFullMask, L.nprobe);
    const unsigned nf = __reduce_add_sync(kFullMask, L.nfast);
    if (lane == 0 && nf) atomicAdd(a.counters + CTR_FAST, (unsigned long long)nf);
    const unsigned nw = __reduce_add_sync(kFullMask, leader ? n_walk : 0u);
    if (lane == 0 && nw) atomicAdd(a.counters + CTR_WALK, (unsigned long long)nw);
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
      if (a.warp_times && G == 1) {
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

constexpr int kGroup = 8;  // lanes per exciton in the trap solver

// kDefer: the lanes leave class 4 and whatever lands in a deep trap to deep_kernel (option deep_thr > 0)
template <typename Draws, int kMinBlocks, bool kInstr, bool kDefer>
__global__ void __launch_bounds__(128, kMinBlocks) kubo_kernel(const KuboArgs a) {
  // The displacement accumulator and the position at the start of the step are only touched when a time step ends;
  // they live in shared memory (one slot per thread) so that the event path does not carry 12 registers of them.
  __shared__ double s_delta[3][128], s_old[3][128];
  hop_loop<Draws, kInstr, 1, kDefer, false>(a, s_delta, s_old);
  if (kDefer && a.overlap) {  // tell the trap kernel, which runs beside this one, that this block will defer nothing more
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.sync, 1u);
    }
  }
}
// the trap solver: class 4 of the previous launch and the excitons kubo_kernel deferred in this one
template <bool kInstr>
__global__ void __launch_bounds__(128, 4) deep_kernel(const KuboArgs a) {
  __shared__ double s_delta[3][128], s_old[3][128];
  hop_loop<PhiloxDraws, kInstr, kGroup, false, true>(a, s_delta, s_old);
}
// the same lists served one exciton per lane by the generic loop: warps that hold trapped excitons only (experiment, option
// deep_group = 1)
template <bool kInstr>
__global__ void __launch_bounds__(128, 4) trap_lanes_kernel(const KuboArgs a) {
  __shared__ double s_delta[3][128], s_old[3][128];
  hop_loop<PhiloxDraws, kInstr, 1, false, true>(a, s_delta, s_old);
}


// file every exciton under its activity class (first launch after creation or after an upload of the population)
__global__ void __launch_bounds__(256) classify_kernel(const Tables T, const int32_t* site, int64_t P, double dt, double deep_thr, ClassLists q) {
  const int64_t  e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int      lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool     mine = e < P;
  const int      cls = mine ? activity_class(ro(&T.site[site[e]].top.total) * dt, deep_thr) : 0;
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

// ---- set-up on the device ---------------------------------------------------------------------------------------------------
// Everything a site contributes to the tables except its row: chain links and the flight times to its two chain neighbours
// (particle.cpp:39-41 for an exciton that sits on the site), its position record, the unit vectors towards the neighbours
// (particle.cpp:47), the segment time shared with its memory neighbour (NaN if they are not each other's chain neighbours),
// and its geometry for the table build.  One thread per site; IEEE sqrt and division, no contraction: the bits of the host
// functions of the same name in host_setup.cpp (which the CPU tests run).
struct SiteSetupArgs {
  const double * px, *py, *pz, *ox, *oy, *oz;  // post-trim site list, struct of arrays
  const int32_t *left, *right;
  int64_t        N;
  double         velocity;
  SiteRec*       site;
  PosRec*        pos;
  DirRec*        dir;
  double*        seg;   // [N], already offset by the padding
  SiteGeom*      geom;  // [N] site order
  int32_t*       flags;
};
__global__ void __launch_bounds__(256) site_records_kernel(const SiteSetupArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const double  x = a.px[i], y = a.py[i], z = a.pz[i];
  const int32_t l = a.left[i], r = a.right[i];
  // create_scatterers + trim_scats always produce symmetric links (monte_carlo.h:249-262, 730-772)
  if ((l > -1 && a.right[l] != (int32_t)i) || (r > -1 && a.left[r] != (int32_t)i) || (l > -1 && l == r)) atomicOr(a.flags + FLAG_BAD_LINKS, 1);
  SiteRec rec{};
  rec.left = l;
  rec.right = r;
  rec.q_left = rec.q_right = 0.0;
  DirRec d{0, 0, 0, 0, 0, 0, 0, 0};
  if (l > -1) {
    rec.q_left = segment_time(x, y, z, a.px[l], a.py[l], a.pz[l], a.velocity);
    const double wx = a.px[l] - x, wy = a.py[l] - y, wz = a.pz[l] - z, nn = norm3(wx, wy, wz), den = (nn > 0) ? nn : 1.0;
    d.lx = wx / den; d.ly = wy / den; d.lz = wz / den;
  }
  if (r > -1) {
    rec.q_right = segment_time(x, y, z, a.px[r], a.py[r], a.pz[r], a.velocity);
    const double wx = a.px[r] - x, wy = a.py[r] - y, wz = a.pz[r] - z, nn = norm3(wx, wy, wz), den = (nn > 0) ? nn : 1.0;
    d.rx = wx / den; d.ry = wy / den; d.rz = wz / den;
  }
  rec.top.nbr0 = rec.top.nbr1 = rec.top.nbr2 = -1;  // rates, row and top entries: the table build
  a.site[i] = rec;
  a.pos[i] = PosRec{x, y, z, 0.0};
  a.dir[i] = d;
  a.geom[i] = SiteGeom{x, y, z, a.ox[i], a.oy[i], a.oz[i]};
  // segment between sites i and i+1: usable by the run walk only if they are each other's chain neighbours and both sides
  // computed the same time
  double sg = __longlong_as_double(0x7ff8000000000000LL);
  if (i + 1 < a.N && r == (int32_t)(i + 1) && a.left[i + 1] == (int32_t)i) {
    const double back = segment_time(a.px[i + 1], a.py[i + 1], a.pz[i + 1], x, y, z, a.velocity);
    if (__double_as_longlong(back) == __double_as_longlong(rec.q_right)) sg = rec.q_right;
  }
  a.seg[i] = sg;
}
// bucket-ordered copy of the geometry: candidates of one cell are contiguous for the table build
__global__ void __launch_bounds__(256) gather_geom_kernel(const SiteGeom* geom, const int32_t* cell_sites, int64_t N, SiteGeom* cell_geom) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < N) cell_geom[q] = geom[cell_sites[q]];
}
__global__ void __launch_bounds__(256) widen_kernel(const uint32_t* deg, int64_t N, uint64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = deg[i];
  if (i == N) out[i] = 0;
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
  int32_t*        flags;
  unsigned long long* counters;
  int32_t*        guard_sites;  // sites whose row holds a pair near a theta midpoint (first guard_cap of them)
  int32_t         guard_cap;
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
    top.store(a.site[i].top, acc, d);  // with scatterer.h:91  _max_rate = neighbors.back().first
    if (d == 0) atomicOr(a.flags + FLAG_EMPTY_ROW, 1);
    if (guard) {
      const unsigned long long k = atomicAdd(a.counters + CTR_GUARD, 1ULL);
      if (k < (unsigned long long)a.guard_cap) a.guard_sites[k] = (int32_t)i;
    }
  }
}

// K1, fill pass, one WARP per site (round 2).  The one-thread-per-site kernel above spends 1.7e11 warp-instructions on C4 with
// 10 of 32 lanes active: every lane walks its own cells and meets its accepted pairs at its own pace.  Here the 32 lanes test
// 32 candidates of the current cell per instruction, the accepted ones are queued in candidate order in shared memory, and
// every time 32 have gathered the warp evaluates their rates together (scatterer.cpp:44-63, all lanes converged), walks the
// sequential FP64 prefix sum over them in order (scatterer.cpp:78-80: the same additions in the same order, one lane-value at
// a time through a shuffle), writes the 32 entries as one coalesced 512-byte store and merges them into the row's three
// widest entries (order: rate descending, row position ascending -- what TopEntries::add keeps).  The search guide is found
// by eight lanes with a binary search each (first k with cum[k] > dice_min(j), else d-1: build_guide's scan, independently per
// bucket, valid because cum is non-decreasing).  Same candidates, same order, same arithmetic: the same bits.
struct TopKey {
  double  rate, lo, hi;
  int32_t nbr, pos;  // pos: position in the row (ties keep the earlier entry); INT32_MAX = empty
};
__device__ __forceinline__ bool top_before(double ra, int32_t pa, double rb, int32_t pb) { return ra > rb || (ra == rb && pa < pb); }
__global__ void __launch_bounds__(128) csr_fill_warp_kernel(const CsrArgs a) {
  __shared__ int32_t s_queue[4][64];
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t  i = (int64_t)blockIdx.x * 4 + w;
  if (i >= a.N) return;
  int32_t*       queue = s_queue[w];
  const SiteGeom s1 = a.geom[i];
  const int      cx = cell_coord(s1.px, a.lo[0], a.radius), cy = cell_coord(s1.py, a.lo[1], a.radius),
            cz = cell_coord(s1.pz, a.lo[2], a.radius);
  const uint64_t base = a.row_begin[i];
  const uint32_t want = (uint32_t)(a.row_begin[i + 1] - base);
  uint32_t d = 0;
  double   acc = 0.0;
  bool     guard = false;
  int      qcount = 0;
  TopKey   top[kTopEntries];  // uniform over the warp
#pragma unroll
  for (int j = 0; j < kTopEntries; ++j) top[j] = TopKey{-1.0, 0.0, 0.0, -1, 0x7fffffff};

  auto process = [&](int m) {  // the first m queued candidates (m <= 32), in order
    double  rate = 0.0;
    int32_t nbr = -1;
    bool    g = false;
    if (lane < m) {
      const int32_t  qq = queue[lane];
      const SiteGeom s2 = a.cell_geom[qq];
      rate = pair_rate(s1, s2, a.R, &g);
      nbr = a.cell_sites[qq];
    }
    guard |= __any_sync(kFullMask, g);
    double cum = 0.0, below = 0.0;
    for (int l = 0; l < m; ++l) {  // scatterer.cpp:78-80, sequential
      const double r = __shfl_sync(kFullMask, rate, l);
      const double b = (d + l == 0) ? -1.0 : acc;
      acc = (d + l == 0) ? r : acc + r;
      if (lane == l) {
        cum = acc;
        below = b;
      }
    }
    if (lane < m) {
      RowEntry en;
      en.cum = cum;
      en.nbr = nbr;
      en.pad = 0;
      a.row[base + d + lane] = en;
    }
    // the three widest of { current top } and { this batch }: three rounds of a warp-wide arg-best
    bool taken = !(lane < m);
    TopKey nt[kTopEntries];
    int    used = 0;  // how many of the old top entries have been consumed (they are sorted)
#pragma unroll
    for (int j = 0; j < kTopEntries; ++j) {
      double  br = taken ? -2.0 : rate;
      int32_t bp = taken ? 0x7fffffff : (int32_t)(d + lane);
      int     bl = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double  r2 = __shfl_xor_sync(kFullMask, br, o);
        const int32_t p2 = __shfl_xor_sync(kFullMask, bp, o);
        const int     l2 = __shfl_xor_sync(kFullMask, bl, o);
        if (top_before(r2, p2, br, bp)) {
          br = r2;
          bp = p2;
          bl = l2;
        }
      }
      // best of the batch (uniform: br, bp, bl) against the next unused old entry
      const bool have_new = br > -2.0 && bp != 0x7fffffff;
      const bool old_wins = used < kTopEntries && top[used].pos != 0x7fffffff &&
                            (!have_new || top_before(top[used].rate, top[used].pos, br, bp));
      if (old_wins) {
        nt[j] = top[used];
        ++used;
      } else if (have_new) {
        nt[j].rate = br;
        nt[j].pos = bp;
        nt[j].lo = __shfl_sync(kFullMask, below, bl);
        nt[j].hi = __shfl_sync(kFullMask, cum, bl);
        nt[j].nbr = __shfl_sync(kFullMask, nbr, bl);
        if (lane == bl) taken = true;
      } else {
        nt[j] = TopKey{-1.0, 0.0, 0.0, -1, 0x7fffffff};
      }
    }
#pragma unroll
    for (int j = 0; j < kTopEntries; ++j) top[j] = nt[j];
    d += (uint32_t)m;
  };

  for (int c = 0; c < 27; ++c) {  // the reference's stencil order: x outermost, z innermost (monte_carlo.h:402-411)
    const int ix = cx - 1 + c / 9, iy = cy - 1 + (c / 3) % 3, iz = cz - 1 + c % 3;
    if (!(ix > -1 && ix < a.nb[0] && iy > -1 && iy < a.nb[1] && iz > -1 && iz < a.nb[2])) continue;
    const int64_t b = (int64_t)ix + (int64_t)iy * a.nb[0] + (int64_t)iz * a.nb[0] * a.nb[1];
    const int64_t q0 = a.cell_start[b], q1 = a.cell_start[b + 1];
    for (int64_t q = q0; q < q1; q += 32) {
      const bool in = q + lane < q1;
      bool       ok = false;
      if (in) ok = within_cutoff(s1, a.cell_geom[q + lane], a.radius);
      const unsigned m = __ballot_sync(kFullMask, ok);
      if (ok) queue[qcount + __popc(m & lt_mask)] = (int32_t)(q + lane);
      qcount += __popc(m);
      __syncwarp();
      if (qcount >= 32) {
        process(32);
        const int32_t keep = (lane + 32 < qcount) ? queue[lane + 32] : 0;
        __syncwarp();
        if (lane + 32 < qcount) queue[lane] = keep;
        qcount -= 32;
        __syncwarp();
      }
    }
  }
  if (qcount > 0) process(qcount);

  if (d != want) atomicOr(a.flags + FLAG_BAD_LINKS, 2);  // the two passes must agree on every row's length
  // search guide: lane j < 8 finds the answer for the smallest dice of bucket j (hop_core.h build_guide)
  __syncwarp();  // the row written by the other lanes is read below
  const uint32_t sc = guide_scale(d);
  uint32_t       gk = 0;
  if (lane < kGuideBuckets && d > 0) {
    const double dm = guide_dice_min(acc, lane);
    uint32_t     lo = 0, hi = d - 1;  // first k with cum[k] > dm, else d-1
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (a.row[base + mid].cum > dm)
        hi = mid;
      else
        lo = mid + 1;
    }
    gk = lo >> sc;
  }
  __syncwarp();
  uint32_t g8[kGuideBuckets];
#pragma unroll
  for (int j = 0; j < kGuideBuckets; ++j) g8[j] = __shfl_sync(kFullMask, gk, j);
  // top entries as intervals of the draw (hop_core.h SiteRec): lane 2j bisects for the lower end of entry j, lane 2j+1 for its upper end
  uint32_t rk = 0;
  if (lane < 2 * kTopEntries) {
    const int    j = lane >> 1;
    const double x = (lane & 1) ? (j == 0 ? top[0].hi : j == 1 ? top[1].hi : top[2].hi) : (j == 0 ? top[0].lo : j == 1 ? top[1].lo : top[2].lo);
    rk = first_draw_reaching(acc, x);
  }
  uint32_t r6[2 * kTopEntries];
#pragma unroll
  for (int j = 0; j < 2 * kTopEntries; ++j) r6[j] = __shfl_sync(kFullMask, rk, j);
  if (lane == 0) {
    SiteRec& r = a.site[i];
    r.inv_total = d ? 1. / acc : 0.0;   // scatterer.h:92
    r.row_begin = (uint32_t)base;
    r.row_len = d;
#pragma unroll
    for (int j = 0; j < kGuideBuckets; ++j) r.guide[j] = (uint8_t)g8[j];
    r.top.iv0 = draw_blocks(r6[0], r6[1]);
    r.top.iv1 = draw_blocks(r6[2], r6[3]);
    r.top.iv2 = draw_blocks(r6[4], r6[5]);
    r.top.nbr0 = d ? top[0].nbr : kEmptyRow; r.top.nbr1 = top[1].nbr; r.top.nbr2 = top[2].nbr;
    r.top.total = acc;                   // scatterer.h:91  _max_rate = neighbors.back().first
    if (d == 0) atomicOr(a.flags + FLAG_EMPTY_ROW, 1);
    if (guard) {
      const unsigned long long k = atomicAdd(a.counters + CTR_GUARD, 1ULL);
      if (k < (unsigned long long)a.guard_cap) a.guard_sites[k] = (int32_t)i;
    }
  }
}

}  // namespace cntmc
