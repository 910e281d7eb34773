// cntmc_api.cu -- the handle behind include/cntmc.h: host orchestration of the CUDA kernels in kernels.cuh.
//
// Reference citations (file:line) are into /root/reference/src.  There is no CPU execution path in this file: every
// entry point that computes anything launches kernels, and fails with CNTMC_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <memory>
#include <type_traits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cntmc.h"
#include "host_setup.h"
#include "kernels.cuh"

using namespace cntmc;

namespace {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StateError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ReplayError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_CHECK(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      throw CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                                     \
  } while (0)

template <typename T>
struct DevBuf {
  T*     p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void swap(DevBuf& o) {
    std::swap(p, o.p);
    std::swap(n, o.n);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) CUDA_CHECK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

thread_local std::string g_create_error;

}  // namespace

// one copy of the exciton population on the device (contact mode keeps two and compacts from one into the other)
struct ExcitonStore {
  DevBuf<double>   px, py, pz, dx, dy, dz, ff;
  DevBuf<int32_t>  site;
  DevBuf<uint8_t>  heading;
  DevBuf<uint32_t> ndraw;
  DevBuf<uint64_t> gid;
  void alloc(size_t n, bool with_gid) {
    px.alloc(n); py.alloc(n); pz.alloc(n);
    dx.alloc(n); dy.alloc(n); dz.alloc(n);
    ff.alloc(n);
    site.alloc(n); heading.alloc(n); ndraw.alloc(n);
    if (with_gid) gid.alloc(n);
  }
  void swap(ExcitonStore& o) {
    px.swap(o.px); py.swap(o.py); pz.swap(o.pz);
    dx.swap(o.dx); dy.swap(o.dy); dz.swap(o.dz);
    ff.swap(o.ff);
    site.swap(o.site); heading.swap(o.heading); ndraw.swap(o.ndraw); gid.swap(o.gid);
  }
  ExcitonArrays view(bool with_gid) const {
    ExcitonArrays S{};
    S.px = px.p; S.py = py.p; S.pz = pz.p;
    S.dx = dx.p; S.dy = dy.p; S.dz = dz.p;
    S.ff = ff.p;
    S.site = site.p;
    S.heading = heading.p;
    S.ndraw = ndraw.p;
    S.gid = with_gid ? gid.p : nullptr;
    return S;
  }
};

struct cntmc_handle {
  mutable std::string err;
  json::Value         block;
  Params              prm;
  int                 device = -1;
  cudaStream_t        stream = nullptr;

  // host-side set-up state
  Mesh      mesh;
  bool      have_mesh = false;
  HostTable table;
  Sites     sites;
  Domain    dom{};
  Buckets   buckets;
  Injection inj;
  bool      initialised = false;
  bool      contact_mode = false;
  int       n_seg = 0;
  std::vector<double>  area;
  std::vector<int32_t> c1_sites, c2_sites;
  int64_t              c1_pop = 0, c2_pop = 0;

  // host mirrors of device tables, fetched when somebody asks (a 5e6-site film: 40 MB and 640 MB of downloads)
  mutable std::vector<uint64_t> row_ptr;  // [N+1]
  mutable std::vector<double>   max_rate, inv_max_rate;
  uint64_t              nnz = 0;
  DevBuf<uint64_t>      d_row_begin;  // [N+1]
  int64_t               midpoint_guards = 0, midpoint_repairs = 0, midpoint_changed = 0;
  int64_t               opt_guard_ppb = 1;
  int64_t               opt_csr_warp = 1;  // table build, fill pass: one warp per row  // theta within this many 1e-9 grid pitches of a midpoint flags its row
  double                csr_seconds = 0;

  // device tables
  DevBuf<SiteRec> d_site;
  DevBuf<double> d_seg;
  DevBuf<DirRec> d_dir;
  DevBuf<PosRec>  d_pos;
  DevBuf<RowEntry> d_row;
  DevBuf<int32_t>  d_inject, d_c1, d_c2;
  DevBuf<double>  d_theta, d_z, d_a1, d_a2, d_rates;
  Tables          T{};

  // excitons
  int64_t          P = 0, capacity = 0;
  ExcitonStore     ex, ex_spare;
  DevBuf<uint32_t> e_keys_out, e_iota, e_perm;
  DevBuf<uint8_t>  d_alive;
  uint64_t         next_gid = 0;
  DevBuf<char>     sort_tmp;
  DrawConfig       draws{};
  bool             replay = false;
  DevBuf<int64_t>  r_off;
  DevBuf<int32_t>  r_draws;
  DevBuf<double>   r_logs;
  int64_t          replay_ids = 0;  // contact replay: ids covered by the draw lists

  // reductions, diagnostics
  DevBuf<double>             d_partial, d_sums;
  DevBuf<StageRec>           d_stage;
  DevBuf<uint32_t>           d_list[2][kClasses], d_list_count[2], d_hand_list[2], d_hand_count;  // hand-over lists: deferred, returned
  DevBuf<unsigned long long> d_list_head;
  DevBuf<double>             d_cur_dt, d_cur_ox, d_cur_oy, d_cur_oz;  // cursors of the excitons deferred to the group solver
  DevBuf<int32_t>            d_cur_step;
  DevBuf<uint32_t>           d_cur_nevent;
  int                        cur_list = 0;
  bool                       have_lists = false;
  double                     lists_deep_thr = 0;  // the class boundary the lists were filed with
  DevBuf<int32_t>            d_flags;
  DevBuf<unsigned long long> d_counters;
  DevBuf<int32_t>            d_trace_sites, d_trace_counts;
  int32_t                    trace_cap = 0;

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // overlap mode of the trap kernel: its own stream, and the events that order it against the handle's stream
  cudaStream_t     trap_stream = nullptr;
  cudaEvent_t      ev_fork = nullptr, ev_join = nullptr;
  DevBuf<uint32_t> d_overlap_sync;
  double      last_ms = 0;
  int64_t     last_launches = 0;

  // tuning
  int64_t opt_chunk = 64;     // time steps per launch
  int64_t opt_dirs = 1;         // last legs that leave from a site use stored unit vectors
  int64_t opt_runs = 1;         // chain walks over memory-consecutive sites read segment times instead of chasing records
  int64_t opt_top_entries = 1;  // the three widest entries of a row are tried before the row is searched
  int64_t opt_hot_pct = 25;   // share of the lane blocks that serve the most active classes first
  int64_t opt_gid_base = 0;   // contact mode: stream ids start at opt_gid_base * 2^56 (cntmc_multi gives every GPU its own range)
  int64_t opt_deep_thr = 0;   // Gamma*dt from which an exciton belongs to the trap solver (0: no trap solver, the default:
                              // parity-green but slower on every workload measured, profiles/round2_trap_solver.txt)
  int64_t opt_deep_blocks = 4;  // blocks per SM of the trap solver's launch
  int64_t opt_deep_group = 8;   // lanes per exciton in the trap solver (8: window walk; 1: the generic loop over trapped excitons only)
  int64_t opt_trap_burst = 1;   // events per lane and warp iteration of the trap-lanes kernel (deep_group = 1; see hop_loop)
  int64_t opt_deep_overlap = 0;  // deep_group = 1 only: the trap kernel runs beside the lane kernel of the same launch and takes
                                 // the deferred excitons while they arrive (see hop_loop)
  int64_t opt_overlap_trap_blocks = 1;  // blocks per SM of the trap kernel in overlap mode; the lane kernel takes the rest of five
  int64_t opt_deep_rounds = 2;  // 2: the trap solver hands excitons that left their trap back to the lanes once per launch
  int64_t opt_occupancy = 0;   // resident 128-thread blocks per SM the hop kernel is compiled for (4 to 8); 0 = by table size, see occupancy_of
  int64_t l2_bytes = -1;
  int64_t opt_stage_mb = 0;  // cap on the (step, exciton) staging buffer in MiB; shortens the launches if needed.
                             // 0 = a third of the device memory that is free when the buffer is first sized
  int     sm_count = 0;
  int64_t opt_stats = 0;         // count cumulative-rate probes and chain crossings (roofline bookkeeping)
  int64_t opt_time_kernels = 0;  // CUDA events around every hop-kernel launch (bench.py's roofline figure)
  std::vector<cudaEvent_t> kernel_events, deep_events;
  double  deep_ms = 0;  // of kernel_ms, the part spent in the trap solver's launches
  double  kernel_ms = 0;
  int64_t kernel_launches = 0;

  // cntmc_kubo_step_host_state: the host-resident population is stepped in slices, each on its own stream with its own
  // exciton buffers and lists but the parent's tables, so that the copies of one slice overlap the kernels of the others
  std::vector<std::unique_ptr<cntmc_handle>> slices;
  bool    is_slice = false;
  int64_t last_chunk = 0;  // time steps per launch the last step call actually used (option chunk_steps capped by the staging budget)
  int64_t grid_share = 1;  // this handle's launches use 1/grid_share of the resident block slots
  int64_t opt_host_slices = 4;
  int64_t opt_slice_share = 0;  // a slice's launches take 1/this of the block slots (0: 1/number of slices, the best measured:
                                // profiles/round2_e2e_slices.txt)

  double  time = 0;  // monte_carlo::_time (never initialised by the reference, monte_carlo.h:45; starts at 0 here)
  int64_t hops = 0, reinjections = 0, crossings = 0, probes = 0;

  ~cntmc_handle() {
    slices.clear();
    if (ev0) cudaEventDestroy(ev0);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (trap_stream) cudaStreamDestroy(trap_stream);
    if (ev1) cudaEventDestroy(ev1);
    if (is_slice && stream) cudaStreamDestroy(stream);
  }

  ExcitonArrays arrays() { return ex.view(contact_mode); }
  // contact mode: room for `cap` work items in both copies, keeping the first `keep` excitons of the live copy
  void reserve_contact(int64_t cap, int64_t keep);
  void          alloc_excitons(int64_t cap) {
    ex.alloc((size_t)cap, contact_mode);
    e_keys_out.alloc((size_t)cap);
    e_iota.alloc((size_t)cap);
    e_perm.alloc((size_t)cap);
    capacity = cap;
  }
};

template <typename T>
static void grow_keep(DevBuf<T>& b, size_t cap, size_t keep, cudaStream_t st) {
  if (cap <= b.n && b.p) return;
  DevBuf<T> nb;
  nb.alloc(cap);
  if (keep && b.p) CUDA_CHECK(cudaMemcpyAsync(nb.p, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  std::swap(b.p, nb.p);
  std::swap(b.n, nb.n);
}
void cntmc_handle::reserve_contact(int64_t cap, int64_t keep) {
  if (cap <= capacity) return;
  const size_t c = (size_t)(cap + cap / 4 + 1024), k = (size_t)keep;
  grow_keep(ex.px, c, k, stream); grow_keep(ex.py, c, k, stream); grow_keep(ex.pz, c, k, stream);
  grow_keep(ex.dx, c, k, stream); grow_keep(ex.dy, c, k, stream); grow_keep(ex.dz, c, k, stream);
  grow_keep(ex.ff, c, k, stream);
  grow_keep(ex.site, c, k, stream); grow_keep(ex.heading, c, k, stream); grow_keep(ex.ndraw, c, k, stream);
  grow_keep(ex.gid, c, k, stream);
  ex_spare.alloc(c, true);
  d_alive.alloc(c);
  e_iota.alloc(c);
  e_perm.alloc(c);
  e_keys_out.alloc(c);
  capacity = (int64_t)c;
}

namespace {

void use_device(const cntmc_t* h);
// host mirrors of the row offsets / the rates, fetched on first use
void need_row_ptr(const cntmc_t* h) {
  if (!h->row_ptr.empty()) return;
  use_device(h);
  h->row_ptr.resize((size_t)h->sites.N + 1);
  CUDA_CHECK(cudaMemcpy(h->row_ptr.data(), h->d_row_begin.p, h->row_ptr.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
}
void need_rates(const cntmc_t* h) {
  if (!h->max_rate.empty()) return;
  use_device(h);
  const size_t         N = (size_t)h->sites.N;
  std::vector<SiteRec> rec(N);
  CUDA_CHECK(cudaMemcpy(rec.data(), h->d_site.p, N * sizeof(SiteRec), cudaMemcpyDeviceToHost));
  h->max_rate.resize(N);
  h->inv_max_rate.resize(N);
  for (size_t i = 0; i < N; ++i) {
    h->max_rate[i] = rec[i].top.total;
    h->inv_max_rate[i] = rec[i].inv_total;
  }
}

void use_device(const cntmc_t* h) {
  if (h->device >= 0) CUDA_CHECK(cudaSetDevice(h->device));
}
// Resident blocks per SM of the hop kernels.  More warps hide more of the dependent-gather latency and cost registers: 7 blocks
// (70 registers, no spills) against 8 (64 registers, 16 bytes of spills in a cold branch) measured on one B200, hops/s:
//   tables in HBM (C4)                    1.03e10 / 1.10e10      Green-Kubo, tables in L2: 1e6 excitons 7.57e9 / 7.55e9,
//   contacts, tables in L2 (C5)           1.53e10 / 1.49e10        2e6 8.74e9 / 8.87e9, 6e6 9.61e9 / 9.89e9, 1e8 6.93e9 / 7.16e9
// so: 8 once the rows come from HBM, or for a Green-Kubo population of two million and more; 7 otherwise.
int64_t occupancy_of(cntmc_t* h) {
  if (h->opt_occupancy != 0) return h->opt_occupancy;
  if (h->l2_bytes < 0) {
    int dev = 0, l2 = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    h->l2_bytes = l2;
  }
  const int64_t tables = (int64_t)h->nnz * (int64_t)sizeof(RowEntry) +
                         (int64_t)h->sites.N * (int64_t)(sizeof(SiteRec) + sizeof(PosRec) + sizeof(DirRec) + sizeof(double));
  if (tables > h->l2_bytes) return 8;
  return (!h->contact_mode && h->P >= 2000000) ? 8 : 7;
}

void require(bool ok, const char* msg) {
  if (!ok) throw std::invalid_argument(msg);
}

std::string expand_home(std::string path) {  // prepare_directory.hpp:12-16
  if (!path.empty() && path[0] == '~') {
    const char* home = std::getenv("HOME");
    path = std::string(home ? home : "") + path.substr(1);
  }
  return path;
}

void check_flags(cntmc_t* h) {
  int32_t flags[FLAG_COUNT];
  h->d_flags.download(flags, FLAG_COUNT, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (flags[FLAG_REPLAY] || flags[FLAG_STUCK]) {  // reported once: a fresh population can use the handle again
    CUDA_CHECK(cudaMemsetAsync(h->d_flags.p, 0, 2 * sizeof(int32_t), h->stream));
    static_assert(FLAG_STUCK == 0 && FLAG_REPLAY == 1, "the two run-time flags are the first two words");
  }
  if (flags[FLAG_REPLAY]) throw ReplayError("a replayed draw list ran out before the end of the run");
  if (flags[FLAG_STUCK]) throw StateError("an exciton exceeded the chain-walk guard (coincident chain sites?) or met a chain-walk operand outside [2^-400, 2^401) m");
}

// monte_carlo::create_scattering_table, "davoody" branch (monte_carlo.cpp:32-49): tube physics for the input's tubes, then
// the table of transfers from the first tube to itself, on this handle's GPU (cntmc_davoody.cu)
void davoody_table(cntmc_t* h) {
  struct TubeHandle {
    cntmc_tube_t* t = nullptr;
    ~TubeHandle() { cntmc_tube_destroy(t); }
  };
  struct TransferHandle {
    cntmc_transfer_t* x = nullptr;
    ~TransferHandle() { cntmc_transfer_destroy(x); }
  };
  std::vector<std::unique_ptr<TubeHandle>> tubes;
  for (const auto& spec : h->prm.tubes) {  // the reference solves every listed tube, then uses the first
    tubes.emplace_back(new TubeHandle);
    tubes.back()->t = cntmc_tube_create(spec[0], spec[1], spec[2]);
    if (!tubes.back()->t) throw std::invalid_argument(std::string("\"cnts\": ") + cntmc_davoody_last_error());
  }
  HostTable      t = make_table_axes(h->prm);
  TransferHandle x;
  // exciton_transfer(cnt1, cnt2) fixes 300 K and 4 meV (exciton_transfer.h:41-42)
  x.x = cntmc_transfer_create(tubes[0]->t, tubes[0]->t, 300, 4.e-3 * (1.6 * std::pow(10, -19.0)), h->device);
  if (!x.x) throw CudaError(cntmc_davoody_last_error());
  const int32_t dims[4] = {(int32_t)t.theta.size(), (int32_t)t.z.size(), (int32_t)t.a1.size(), (int32_t)t.a2.size()};
  t.rates.resize(t.theta.size() * t.z.size() * t.a1.size() * t.a2.size());
  if (cntmc_transfer_table(x.x, dims, t.theta.data(), t.z.data(), t.a1.data(), t.a2.data(), t.rates.data()) != CNTMC_OK)
    throw CudaError(cntmc_davoody_last_error());
  h->table = std::move(t);
}

// ---- set-up common to kubo_init and init -------------------------------------------------------------------------------
// monte_carlo.cpp:262-298 / monte_carlo.h:166-186: table, scatterers, trim, domain, buckets, set_max_rate
void common_init(cntmc_t* h) {
  require(h->have_mesh || !h->prm.mesh_dir.empty(), "no mesh: call cntmc_load_mesh / cntmc_set_mesh first");
  if (!h->have_mesh) {
    h->mesh = load_mesh(expand_home(h->prm.mesh_dir));
    h->have_mesh = true;
  }
  if (h->table.empty() && h->prm.rate_type == "davoody" && !h->prm.tubes.empty()) davoody_table(h);
  if (h->table.empty()) h->table = make_rate_table(h->prm);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw CudaError("no CUDA device: this engine has no CPU execution path");
  use_device(h);
  h->sites = create_sites(h->mesh);
  trim_sites(h->sites, h->prm.xlim, h->prm.ylim, h->prm.zlim);
  h->dom = find_domain(h->sites);
  const double  R = h->prm.max_hopping_radius;
  const int64_t N = h->sites.N;
  require(R > 0, "\"max hopping radius [m]\" must be positive");
  h->buckets = build_buckets(h->sites, h->dom, R);

  if (!h->ev0) {
    CUDA_CHECK(cudaEventCreate(&h->ev0));
    CUDA_CHECK(cudaEventCreate(&h->ev1));
  }
  cudaStream_t st = h->stream;
  h->d_flags.alloc(FLAG_COUNT);
  h->d_counters.alloc(CTR_COUNT);
  CUDA_CHECK(cudaMemsetAsync(h->d_flags.p, 0, FLAG_COUNT * sizeof(int32_t), st));
  CUDA_CHECK(cudaMemsetAsync(h->d_counters.p, 0, CTR_COUNT * sizeof(unsigned long long), st));

  // rate table
  const HostTable& t = h->table;
  h->d_theta.upload(t.theta, st);
  h->d_z.upload(t.z, st);
  h->d_a1.upload(t.a1, st);
  h->d_a2.upload(t.a2, st);
  h->d_rates.upload(t.rates, st);

  // site list (struct of arrays, 56 bytes per site) and the bucket order go up once; records, geometry copies and the row
  // offsets are made on the device
  DevBuf<double>   d_soa[6];
  DevBuf<int32_t>  d_left, d_right, d_cell_sites;
  DevBuf<SiteGeom> d_geom, d_cell_geom;
  DevBuf<int64_t>  d_cell_start;
  DevBuf<uint32_t> d_deg;
  for (int c = 0; c < 3; ++c) {
    d_soa[c].upload(h->sites.pos[c], st);
    d_soa[3 + c].upload(h->sites.orient[c], st);
  }
  d_left.upload(h->sites.left, st);
  d_right.upload(h->sites.right, st);
  d_cell_sites.upload(h->buckets.sites, st);
  d_cell_start.upload(h->buckets.start, st);
  d_deg.alloc((size_t)N);
  d_geom.alloc((size_t)N);
  d_cell_geom.alloc((size_t)N);
  h->d_site.alloc((size_t)N);
  h->d_pos.alloc((size_t)N);
  h->d_dir.alloc((size_t)N);
  h->d_seg.alloc((size_t)N + 2 * kSegPad);
  CUDA_CHECK(cudaMemsetAsync(h->d_seg.p, 0xff, ((size_t)N + 2 * kSegPad) * sizeof(double), st));  // all-ones = a NaN: the padding
  {
    SiteSetupArgs sa{};
    sa.px = d_soa[0].p; sa.py = d_soa[1].p; sa.pz = d_soa[2].p;
    sa.ox = d_soa[3].p; sa.oy = d_soa[4].p; sa.oz = d_soa[5].p;
    sa.left = d_left.p;
    sa.right = d_right.p;
    sa.N = N;
    sa.velocity = h->prm.velocity;
    sa.site = h->d_site.p;
    sa.pos = h->d_pos.p;
    sa.dir = h->d_dir.p;
    sa.seg = h->d_seg.p + kSegPad;
    sa.geom = d_geom.p;
    sa.flags = h->d_flags.p;
    site_records_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(sa);
    CUDA_CHECK(cudaGetLastError());
    gather_geom_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_geom.p, d_cell_sites.p, N, d_cell_geom.p);
    CUDA_CHECK(cudaGetLastError());
  }

  CsrArgs a{};
  a.geom = d_geom.p;
  a.cell_geom = d_cell_geom.p;
  a.cell_sites = d_cell_sites.p;
  a.cell_start = d_cell_start.p;
  for (int c = 0; c < 3; ++c) {
    a.nb[c] = h->buckets.n[c];
    a.lo[c] = h->dom.lo[c];
  }
  a.radius = R;
  a.N = N;
  a.R.theta = h->d_theta.p; a.R.z = h->d_z.p; a.R.a1 = h->d_a1.p; a.R.a2 = h->d_a2.p; a.R.rates = h->d_rates.p;
  a.R.n_theta = (int32_t)t.theta.size(); a.R.n_z = (int32_t)t.z.size();
  a.R.n_a1 = (int32_t)t.a1.size(); a.R.n_a2 = (int32_t)t.a2.size();
  a.R.guard_tol = 1e-9 * (double)h->opt_guard_ppb;
  DevBuf<int32_t> d_guard_sites;
  const int32_t   guard_cap = 1 << 20;
  d_guard_sites.alloc((size_t)guard_cap);
  a.guard_sites = d_guard_sites.p;
  a.guard_cap = guard_cap;
  grid_hint(t.theta.data(), a.R.n_theta, &a.R.start[0], &a.R.inv_step[0]);
  grid_hint(t.z.data(), a.R.n_z, &a.R.start[1], &a.R.inv_step[1]);
  grid_hint(t.a1.data(), a.R.n_a1, &a.R.start[2], &a.R.inv_step[2]);
  grid_hint(t.a2.data(), a.R.n_a2, &a.R.start[3], &a.R.inv_step[3]);
  a.deg = d_deg.p;
  a.site = h->d_site.p;
  a.flags = h->d_flags.p;
  a.counters = h->d_counters.p;

  const int      block = 128;
  const unsigned grid = (unsigned)((N + block - 1) / block);
  CUDA_CHECK(cudaEventRecord(h->ev0, st));
  csr_rows_kernel<false><<<grid, block, 0, st>>>(a);
  CUDA_CHECK(cudaGetLastError());
  // row offsets: exclusive scan of the degrees on the device (stable in site index by construction); only nnz comes back
  h->d_row_begin.alloc((size_t)N + 1);
  widen_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(d_deg.p, N, h->d_row_begin.p);
  CUDA_CHECK(cudaGetLastError());
  {
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, h->d_row_begin.p, h->d_row_begin.p, (int)(N + 1), st));
    h->sort_tmp.alloc(bytes);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(h->sort_tmp.p, bytes, h->d_row_begin.p, h->d_row_begin.p, (int)(N + 1), st));
  }
  uint64_t nnz = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nnz, h->d_row_begin.p + N, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (nnz >= (1ull << 32)) throw std::invalid_argument("neighbour table has >= 2^32 entries; not supported by this build");
  h->nnz = nnz;
  h->row_ptr.clear();
  h->max_rate.clear();
  h->inv_max_rate.clear();
  h->d_row.alloc((size_t)nnz);
  a.row_begin = h->d_row_begin.p;
  a.row = h->d_row.p;
  if (h->opt_csr_warp)  // one warp per row (default); 0: the one-thread-per-row kernel, kept as the cross-check
    csr_fill_warp_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(a);
  else
    csr_rows_kernel<true><<<grid, block, 0, st>>>(a);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(h->ev1, st));
  int32_t            flags[FLAG_COUNT];
  unsigned long long ctrs[CTR_COUNT];
  h->d_flags.download(flags, FLAG_COUNT, st);
  h->d_counters.download(ctrs, CTR_COUNT, st);
  CUDA_CHECK(cudaStreamSynchronize(st));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->csr_seconds = ms * 1e-3;
  h->midpoint_guards = (int64_t)ctrs[CTR_GUARD];
  h->midpoint_repairs = h->midpoint_changed = 0;
  if (h->midpoint_guards > 0) {  // expected: none.  The flagged rows again, on the host, with glibc's acos; patched in place.
    if (h->midpoint_guards > guard_cap) throw StateError("too many rows with a theta next to a grid midpoint to repair");
    std::vector<int32_t> gs((size_t)h->midpoint_guards);
    d_guard_sites.download(gs.data(), gs.size(), st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    std::sort(gs.begin(), gs.end());
    RateTable RH = a.R;  // the same table through host pointers
    RH.theta = t.theta.data(); RH.z = t.z.data(); RH.a1 = t.a1.data(); RH.a2 = t.a2.data(); RH.rates = t.rates.data();
    auto geom_of = [&](int64_t k) {
      return SiteGeom{h->sites.pos[0][(size_t)k], h->sites.pos[1][(size_t)k], h->sites.pos[2][(size_t)k],
                      h->sites.orient[0][(size_t)k], h->sites.orient[1][(size_t)k], h->sites.orient[2][(size_t)k]};
    };
    std::vector<RowEntry> row, old;
    for (int32_t i : gs) {
      const SiteGeom s1 = geom_of(i);
      const int      cx = cell_coord(s1.px, a.lo[0], R), cy = cell_coord(s1.py, a.lo[1], R), cz = cell_coord(s1.pz, a.lo[2], R);
      row.clear();
      TopEntries top;
      top.clear();
      double acc = 0.0;
      for (int c = 0; c < 27; ++c) {  // the kernel's enumeration: monte_carlo.h:402-411, cells in list order
        const int ix = cx - 1 + c / 9, iy = cy - 1 + (c / 3) % 3, iz = cz - 1 + c % 3;
        if (!(ix > -1 && ix < a.nb[0] && iy > -1 && iy < a.nb[1] && iz > -1 && iz < a.nb[2])) continue;
        const int64_t b = (int64_t)ix + (int64_t)iy * a.nb[0] + (int64_t)iz * a.nb[0] * a.nb[1];
        for (int64_t q = h->buckets.start[(size_t)b]; q < h->buckets.start[(size_t)b + 1]; ++q) {
          const int32_t  j = h->buckets.sites[(size_t)q];
          const SiteGeom s2 = geom_of(j);
          if (!within_cutoff(s1, s2, R)) continue;
          const double rate = pair_rate(s1, s2, RH, nullptr);  // host build: acos is glibc's
          const double below = row.empty() ? -1.0 : acc;
          acc = row.empty() ? rate : acc + rate;
          top.add(rate, below, acc, j);
          row.push_back(RowEntry{acc, j, 0});
        }
      }
      uint64_t be[2];
      CUDA_CHECK(cudaMemcpyAsync(be, h->d_row_begin.p + i, sizeof be, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (be[1] - be[0] != row.size()) throw StateError("host and device disagree on the length of a row");
      old.resize(row.size());
      CUDA_CHECK(cudaMemcpyAsync(old.data(), h->d_row.p + be[0], row.size() * sizeof(RowEntry), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      ++h->midpoint_repairs;
      if (memcmp(old.data(), row.data(), row.size() * sizeof(RowEntry)) == 0) continue;  // glibc agrees: nothing to patch
      ++h->midpoint_changed;
      SiteRec rec;
      CUDA_CHECK(cudaMemcpyAsync(&rec, h->d_site.p + i, sizeof rec, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      rec.inv_total = row.empty() ? 0.0 : 1. / acc;
      struct CumView {
        const RowEntry* r;
        __host__ __device__ double operator[](uint32_t k) const { return r[k].cum; }  // build_guide is a host-device template
      };
      build_guide(CumView{row.data()}, (uint32_t)row.size(), acc, rec.guide);
      top.store(rec.top, acc, (uint32_t)row.size());  // with scatterer.h:91 _max_rate
      CUDA_CHECK(cudaMemcpyAsync(h->d_row.p + be[0], row.data(), row.size() * sizeof(RowEntry), cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(h->d_site.p + i, &rec, sizeof rec, cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
    }
  }
  if (flags[FLAG_BAD_LINKS] & 2) throw StateError("table build: count pass and fill pass disagree on the length of a row");
  if (flags[FLAG_BAD_LINKS]) throw std::invalid_argument("chain links of the site list are not symmetric");
  if (flags[FLAG_EMPTY_ROW])
    throw StateError("a site has no neighbour inside the hopping radius (undefined behaviour in the reference, scatterer.h:91)");

  h->T.site = h->d_site.p;
  h->T.seg = h->opt_runs ? h->d_seg.p + kSegPad : nullptr;
  h->T.dir = h->opt_dirs ? h->d_dir.p : nullptr;
  h->T.pos = h->d_pos.p;
  h->T.row = h->d_row.p;
  h->T.velocity = h->prm.velocity;
  h->T.inv_velocity = 1.0 / h->prm.velocity;
  h->time = 0;
  h->hops = h->reinjections = 0;
}

template <typename Draws>
void launch_create(cntmc_t* h, int64_t P, const int32_t* d_list, int32_t n_list) {
  CreateArgs a{};
  a.T = h->T;
  a.S = h->arrays();
  a.draws = h->draws;
  a.P = P;
  a.site_list = d_list;
  a.n_list = n_list;
  a.flags = h->d_flags.p;
  const int block = 256;
  create_excitons_kernel<Draws><<<(unsigned)((P + block - 1) / block), block, 0, h->stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
}

void create_common(cntmc_t* h, int64_t P) {
  require(h->initialised && !h->contact_mode, "call cntmc_kubo_init first");
  require(P > 0, "number of particles must be positive");
  require(!h->inj.sites.empty(), "the injection region holds no site");
  use_device(h);
  h->alloc_excitons(P);
  if (P != h->P) h->trace_cap = 0;  // the trace buffers were sized for the previous population
  h->P = P;
  h->have_lists = false;
  h->time = 0;
  h->hops = h->reinjections = 0;
  h->crossings = h->probes = 0;
  CUDA_CHECK(cudaMemsetAsync(h->d_counters.p, 0, CTR_COUNT * sizeof(unsigned long long), h->stream));
  if (h->replay)
    launch_create<ReplayDraws>(h, P, h->d_inject.p, (int32_t)h->inj.sites.size());
  else
    launch_create<PhiloxDraws>(h, P, h->d_inject.p, (int32_t)h->inj.sites.size());
  check_flags(h);
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

// ---- self-test of the call-free square root and divisions of the chain walk (hop_core.h sqrt_walk / div3_walk) --------------
// counts[0]: square roots that differ from sqrt(), [1]: quotients that differ from '/', [2]: in-range operands flagged,
// [3]: out-of-range operands NOT flagged.  Operands: full-range mantissas, exponents over the accepted range, with the
// numerators of a quotient at most the divisor in magnitude (components of a vector over its norm) or exactly zero.
__device__ __forceinline__ uint64_t mix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double make_double(uint64_t mant, int exp2, bool neg) {  // (1 + mant / 2^52) * 2^exp2
  return __longlong_as_double((long long)(((uint64_t)neg << 63) | ((uint64_t)(exp2 + 1023) << 52) | (mant & 0xfffffffffffffull)));
}
__global__ void __launch_bounds__(256) walk_arith_kernel(int64_t per_thread, uint64_t seed, unsigned long long* counts) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x + 1));
  unsigned long long bad_sqrt = 0, bad_div = 0, false_flag = 0, missed = 0;
  for (int64_t it = 0; it < per_thread; ++it) {
    const uint64_t a = mix64(s), b = mix64(s), c = mix64(s), d = mix64(s), e = mix64(s);
    // square root: the sum of squares of a separation, any exponent of the accepted range; every 16th a perfect square
    double x = make_double(a, (int)(b % 799u) - 399, false);
    if ((b >> 40 & 15u) == 0u) {
      const double r = make_double(a & 0xffffffc000000ull, ((int)(b % 799u) - 399) / 2, false);  // 26 significant bits: r * r is exact
      x = r * r;
    }
    bool flag = false;
    const double got = sqrt_walk(x, flag), want = sqrt(x);
    bad_sqrt += __double_as_longlong(got) != __double_as_longlong(want);
    false_flag += flag;
    // divisions: a vector over its norm
    const int    ed = (int)(c % 711u) - 320;  // the smallest numerator below is 72 binades under the divisor
    const double den = make_double(c >> 10, ed, false);
    double       w[3];
    const uint64_t m[3] = {d, e, a ^ (c << 7)};
    for (int k = 0; k < 3; ++k) {
      const int down = (int)((m[k] >> 54) % 9u);           // 0..8: how far below the divisor, in steps of 8 binades
      w[k] = make_double(m[k], ed - 1 - 8 * down - (int)(m[k] >> 60 & 7u), (m[k] >> 53 & 1u) != 0);
      if ((m[k] >> 48 & 31u) == 0u) w[k] = (m[k] >> 53 & 1u) ? -0.0 : 0.0;
      if ((m[k] >> 48 & 31u) == 1u) w[k] = (m[k] >> 53 & 1u) ? -den : den;
    }
    double ux, uy, uz;
    flag = false;
    div3_walk_dev(w[0], w[1], w[2], den, ux, uy, uz, flag);
    bad_div += __double_as_longlong(ux) != __double_as_longlong(w[0] / den);
    bad_div += __double_as_longlong(uy) != __double_as_longlong(w[1] / den);
    bad_div += __double_as_longlong(uz) != __double_as_longlong(w[2] / den);
    false_flag += flag;
    // outside the accepted range: must be flagged
    if ((it & 255) == 0) {
      const int    out = (a & 1u) ? 402 + (int)(b % 600u) : -401 - (int)(b % 600u);
      const double xo = make_double(a, out, false);
      flag = false;
      (void)sqrt_walk(xo, flag);
      missed += !flag;
      flag = false;
      div3_walk_dev(w[0], w[1], w[2], xo, ux, uy, uz, flag);
      missed += !flag;
      flag = false;
      div3_walk_dev(w[0], make_double(d, -401 - (int)(b % 600u), false), w[2], den, ux, uy, uz, flag);
      missed += !flag;
    }
  }
  bad_sqrt = __reduce_add_sync(kFullMask, (unsigned)bad_sqrt);
  bad_div = __reduce_add_sync(kFullMask, (unsigned)bad_div);
  false_flag = __reduce_add_sync(kFullMask, (unsigned)false_flag);
  missed = __reduce_add_sync(kFullMask, (unsigned)missed);
  if ((threadIdx.x & 31) == 0) {
    if (bad_sqrt) atomicAdd(counts + 0, bad_sqrt);
    if (bad_div) atomicAdd(counts + 1, bad_div);
    if (false_flag) atomicAdd(counts + 2, false_flag);
    if (missed) atomicAdd(counts + 3, missed);
  }
}


template <typename Draws, bool kInstr>
void launch_kubo_i(cntmc_t* h, const KuboArgs& a, unsigned grid, cudaStream_t st, bool defer) {
  if constexpr (std::is_same<Draws, PhiloxDraws>::value) {
    if (defer) {  // with the trap solver (one occupancy only: an opt-in path)
      kubo_kernel<Draws, 5, kInstr, true><<<grid, 128, 0, st>>>(a);
      return;
    }
  }
  switch (occupancy_of(h)) {
    case 4: kubo_kernel<Draws, 4, kInstr, false><<<grid, 128, 0, st>>>(a); break;
    case 6: kubo_kernel<Draws, 6, kInstr, false><<<grid, 128, 0, st>>>(a); break;
    case 7: kubo_kernel<Draws, 7, kInstr, false><<<grid, 128, 0, st>>>(a); break;
    case 8: kubo_kernel<Draws, 8, kInstr, false><<<grid, 128, 0, st>>>(a); break;
    default: kubo_kernel<Draws, 5, kInstr, false><<<grid, 128, 0, st>>>(a); break;
  }
}
template <typename Draws>
void launch_kubo(cntmc_t* h, const KuboArgs& a, unsigned grid, cudaStream_t st, bool defer) {
  // the instrumented variant (site traces, probe / crossing counters) runs only when somebody asked for its output
  if (h->trace_cap > 0 || h->opt_stats) {
    // the residency counters describe the last instrumented launch only
    CUDA_CHECK(cudaMemsetAsync(h->d_counters.p + CTR_WARP_NS, 0, (CTR_COUNT - CTR_WARP_NS) * sizeof(unsigned long long), st));
    const char* dump = getenv("CNTMC_DEBUG_WARP_TIMES");  // diagnostics: per-warp timestamps of this launch -> file
    if (dump && *dump) {
      KuboArgs                   b = a;
      DevBuf<unsigned long long> d_times;
      const size_t               n = (size_t)grid * 4 * kWarpTimeCols;
      d_times.alloc(n);
      CUDA_CHECK(cudaMemsetAsync(d_times.p, 0, n * sizeof(unsigned long long), st));
      b.warp_times = d_times.p;
      launch_kubo_i<Draws, true>(h, b, grid, st, defer);
      std::vector<unsigned long long> host(n);
      d_times.download(host.data(), n, st);
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (FILE* f = fopen(dump, "wb")) {
        fwrite(host.data(), sizeof(unsigned long long), n, f);
        fclose(f);
      }
      return;
    }
    launch_kubo_i<Draws, true>(h, a, grid, st, defer);
  } else
    launch_kubo_i<Draws, false>(h, a, grid, st, defer);
}

// nsteps x kubo_step on the device; sums -> dev_sums[nsteps][4]
void kubo_step_device(cntmc_t* h, double dt, int64_t nsteps, double* dev_sums) {
  require(h->initialised && !h->contact_mode, "call cntmc_kubo_init first");
  require(h->P > 0, "no excitons: call cntmc_kubo_create_particles first");
  require(nsteps > 0, "nsteps must be positive");
  use_device(h);
  cudaStream_t st = h->stream;
  if (h->sm_count == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  // time steps per launch: the option, capped so that the (step, exciton) staging buffer stays within its budget
  if (h->opt_stage_mb <= 0) {
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    h->opt_stage_mb = std::max<int64_t>(256, (int64_t)((free_b + h->d_stage.n * sizeof(StageRec)) / 3) >> 20);
  }
  const int64_t by_budget = std::max<int64_t>(1, (h->opt_stage_mb << 20) / ((int64_t)sizeof(StageRec) * h->P));
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>({h->opt_chunk, by_budget, nsteps}));
  // persistent grid: as many 128-thread blocks as the SMs hold at the compiled occupancy, never more than needed
  const int64_t  want = (h->P + 127) / 128;
  // (a slice of a host-resident population takes its share of the block slots: the slices' launches run side by side)
  // (CNTMC_DBG_BLOCKS_PER_SM: fewer resident blocks than the kernel was compiled for -- separates the price of a variant's spills
  // from the gain of its extra blocks, profiles/round2_trap_solver.txt section 15)
  const char*    dbg_bps = getenv("CNTMC_DBG_BLOCKS_PER_SM");
  const int64_t  bps = dbg_bps ? atoll(dbg_bps) : occupancy_of(h);
  const unsigned grid = (unsigned)std::min<int64_t>(want, std::max<int64_t>(1, (int64_t)h->sm_count * bps / h->grid_share));
  h->last_chunk = chunk;
  h->d_stage.alloc((size_t)chunk * (size_t)h->P);
  h->d_partial.alloc((size_t)chunk * kStageSplits * 4);
  // activity-class lists, double-buffered: [cur] is read by a launch, [1-cur] is filled by it
  for (int b = 0; b < 2; ++b) {
    for (int c = 0; c < kClasses; ++c) h->d_list[b][c].alloc((size_t)h->P);
    h->d_list_count[b].alloc(kClasses);
  }
  h->d_list_head.alloc(kLists);
  // group solver (counter-based streams only: a replayed stream has nothing to prepare in parallel)
  const bool   deep = h->opt_deep_thr > 0 && !h->replay;
  const double deep_thr = deep ? (double)h->opt_deep_thr : INFINITY;
  h->d_hand_list[0].alloc((size_t)h->P);
  h->d_hand_list[1].alloc((size_t)h->P);
  h->d_hand_count.alloc(2);
  if (deep) {
    const bool fresh = h->d_cur_step.n < (size_t)h->P || !h->d_cur_step.p;
    h->d_cur_dt.alloc((size_t)h->P); h->d_cur_ox.alloc((size_t)h->P); h->d_cur_oy.alloc((size_t)h->P); h->d_cur_oz.alloc((size_t)h->P);
    h->d_cur_step.alloc((size_t)h->P);
    h->d_cur_nevent.alloc((size_t)h->P);
    if (fresh) {
      // Every cursor is written by the lane that defers its exciton before anybody reads it.  The arrays are zeroed once all the
      // same: in overlap mode the reader is another kernel running at the same time, ordered by release / acquire on the list
      // entry, and compute-sanitizer's initcheck (whose shadow memory is not part of that ordering) otherwise reports a handful
      // of first-ever cursor reads per run, a different handful every time, with bit-identical results.
      CUDA_CHECK(cudaMemsetAsync(h->d_cur_step.p, 0, (size_t)h->P * sizeof(int32_t), st));
      CUDA_CHECK(cudaMemsetAsync(h->d_cur_nevent.p, 0, (size_t)h->P * sizeof(uint32_t), st));
      for (DevBuf<double>* b : {&h->d_cur_dt, &h->d_cur_ox, &h->d_cur_oy, &h->d_cur_oz})
        CUDA_CHECK(cudaMemsetAsync(b->p, 0, (size_t)h->P * sizeof(double), st));
    }
  }
  auto lists = [&](int read_buf) {
    ClassLists q{};
    for (int c = 0; c < kClasses; ++c) {
      q.list[c] = h->d_list[read_buf][c].p;
      q.next_list[c] = h->d_list[1 - read_buf][c].p;
    }
    q.list[kDeferred] = h->d_hand_list[0].p;
    q.list[kReturned] = h->d_hand_list[1].p;
    q.hand_count = h->d_hand_count.p;
    q.hand_list[0] = h->d_hand_list[0].p;
    q.hand_list[1] = h->d_hand_list[1].p;
    q.hand_to = h->d_hand_count.p;
    q.count = h->d_list_count[read_buf].p;
    q.next_count = h->d_list_count[1 - read_buf].p;
    q.head = h->d_list_head.p;
    return q;
  };
  h->last_launches = 0;
  CUDA_CHECK(cudaEventRecord(h->ev0, st));
  if (!h->have_lists || h->lists_deep_thr != deep_thr) {  // population just created or uploaded: file everything once
    CUDA_CHECK(cudaMemsetAsync(h->d_list_count[h->cur_list].p, 0, kClasses * sizeof(uint32_t), st));
    classify_kernel<<<(unsigned)((h->P + 255) / 256), 256, 0, st>>>(h->T, h->ex.site.p, h->P, dt, deep_thr, lists(1 - h->cur_list));
    CUDA_CHECK(cudaGetLastError());
    h->have_lists = true;
    h->lists_deep_thr = deep_thr;
    h->last_launches += 1;
  }
  for (int64_t done = 0; done < nsteps; done += chunk) {
    const int n = (int)std::min(chunk, nsteps - done);
    CUDA_CHECK(cudaMemsetAsync(h->d_list_head.p, 0, kLists * sizeof(unsigned long long), st));
    CUDA_CHECK(cudaMemsetAsync(h->d_list_count[1 - h->cur_list].p, 0, kClasses * sizeof(uint32_t), st));
    CUDA_CHECK(cudaMemsetAsync(h->d_hand_count.p, 0, 2 * sizeof(uint32_t), st));
    KuboArgs a{};
    a.T = h->T;
    a.S = h->arrays();
    a.C.dt_rem = h->d_cur_dt.p; a.C.ox = h->d_cur_ox.p; a.C.oy = h->d_cur_oy.p; a.C.oz = h->d_cur_oz.p;
    a.C.step = h->d_cur_step.p;
    a.C.nevent = h->d_cur_nevent.p;
    a.draws = h->draws;
    a.q = lists(h->cur_list);
    a.round = 1;
    a.yield_on = (deep && h->opt_deep_rounds > 1) ? 1 : 0;
    a.n_sites = h->sites.N;
    a.hot_blocks = (int32_t)((int64_t)grid * h->opt_hot_pct / 100);
    a.top_entries = (int32_t)h->opt_top_entries;
    a.burst = 1;
    a.deep_thr = deep_thr;
    a.deep_rate = a.deep_thr / dt;
    a.P = h->P;
    a.dt = dt;
    a.nsteps = n;
    a.stage = h->d_stage.p;
    a.trace_sites = h->trace_cap ? h->d_trace_sites.p : nullptr;
    a.trace_counts = h->trace_cap ? h->d_trace_counts.p : nullptr;
    a.trace_cap = h->trace_cap;
    a.flags = h->d_flags.p;
    a.counters = h->d_counters.p;
    cudaEvent_t k0 = nullptr, k1 = nullptr;
    if (h->opt_time_kernels) {
      CUDA_CHECK(cudaEventCreate(&k0));
      CUDA_CHECK(cudaEventCreate(&k1));
      h->kernel_events.push_back(k0);
      h->kernel_events.push_back(k1);
      CUDA_CHECK(cudaEventRecord(k0, st));
    }
    // Round 1: the lanes serve classes 0-3 and defer what lands in a deep trap; the trap solver serves class 4 and the
    // deferred, and hands back what ends a time step outside a trap.  Round 2: the lanes finish the returned (deferring
    // again), the trap solver finishes whatever is left and keeps it.  Kernels whose lists are empty leave at once.
    const bool overlap = deep && h->opt_deep_overlap != 0 && h->opt_deep_group == 1;
    const int  rounds = (deep && !overlap) ? (int)std::max<int64_t>(1, std::min<int64_t>(2, h->opt_deep_rounds)) : 1;
    if (overlap) {
      // One launch = the lane kernel on the handle's stream and the trap kernel on a stream of its own, side by side.  The
      // lane kernel never waits for the trap kernel, so the pair cannot deadlock whatever order the blocks are placed in; the
      // trap kernel leaves when the lane kernel has finished and the deferred list is empty.
      if (!h->trap_stream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&h->trap_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      }
      h->d_overlap_sync.alloc(32);
      CUDA_CHECK(cudaMemsetAsync(h->d_overlap_sync.p, 0, 32 * sizeof(uint32_t), st));
      CUDA_CHECK(cudaMemsetAsync(h->d_hand_list[0].p, 0, (size_t)h->P * sizeof(uint32_t), st));  // empty slots read 0
      const int64_t  tb = std::max<int64_t>(1, std::min<int64_t>(4, h->opt_overlap_trap_blocks));
      const unsigned lane_grid = (unsigned)std::min<int64_t>(want, std::max<int64_t>(1, (int64_t)h->sm_count * (5 - tb)));
      const unsigned trap_grid = (unsigned)((int64_t)h->sm_count * tb);
      a.yield_on = 0;
      a.overlap = 1;
      a.lane_grid = (int32_t)lane_grid;
      a.sync = h->d_overlap_sync.p;
      a.hot_blocks = (int32_t)((int64_t)lane_grid * h->opt_hot_pct / 100);
      CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
      CUDA_CHECK(cudaStreamWaitEvent(h->trap_stream, h->ev_fork, 0));
      launch_kubo<PhiloxDraws>(h, a, lane_grid, st, true);
      CUDA_CHECK(cudaGetLastError());
      KuboArgs b = a;
      b.burst = (int32_t)h->opt_trap_burst;
      if (h->trace_cap > 0 || h->opt_stats)
        trap_lanes_kernel<true><<<trap_grid, 128, 0, h->trap_stream>>>(b);
      else
        trap_lanes_kernel<false><<<trap_grid, 128, 0, h->trap_stream>>>(b);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaEventRecord(h->ev_join, h->trap_stream));
      CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
      h->last_launches += 2;
    }
    for (int round = 1; round <= rounds && !overlap; ++round) {
      a.round = round;
      a.yield_on = (deep && round < rounds) ? 1 : 0;
      if (round == 2) {  // the deferred list starts again (the trap solver has emptied it)
        CUDA_CHECK(cudaMemsetAsync(h->d_hand_count.p, 0, sizeof(uint32_t), st));
        CUDA_CHECK(cudaMemsetAsync(h->d_list_head.p + kDeferred, 0, sizeof(unsigned long long), st));
      }
      if (h->replay)
        launch_kubo<ReplayDraws>(h, a, grid, st, false);
      else
        launch_kubo<PhiloxDraws>(h, a, grid, st, deep);
      CUDA_CHECK(cudaGetLastError());
      h->last_launches += 1;
      if (deep) {
        cudaEvent_t d0 = nullptr, d1 = nullptr;
        if (h->opt_time_kernels) {
          CUDA_CHECK(cudaEventCreate(&d0));
          CUDA_CHECK(cudaEventCreate(&d1));
          h->deep_events.push_back(d0);
          h->deep_events.push_back(d1);
          CUDA_CHECK(cudaEventRecord(d0, st));
        }
        const unsigned dgrid = (unsigned)((int64_t)h->sm_count * h->opt_deep_blocks);
        const bool instr = h->trace_cap > 0 || h->opt_stats;
        if (h->opt_deep_group == 1) {
          const unsigned lgrid = (unsigned)((int64_t)h->sm_count * 5);
          KuboArgs       b = a;
          b.burst = (int32_t)h->opt_trap_burst;
          if (instr)
            trap_lanes_kernel<true><<<lgrid, 128, 0, st>>>(b);
          else
            trap_lanes_kernel<false><<<lgrid, 128, 0, st>>>(b);
        } else if (instr)
          deep_kernel<true><<<dgrid, 128, 0, st>>>(a);
        else
          deep_kernel<false><<<dgrid, 128, 0, st>>>(a);
        CUDA_CHECK(cudaGetLastError());
        if (d1) CUDA_CHECK(cudaEventRecord(d1, st));
        h->last_launches += 1;
      }
    }
    if (k1) CUDA_CHECK(cudaEventRecord(k1, st));
    h->cur_list = 1 - h->cur_list;
    reduce_stage_kernel<<<dim3((unsigned)n, kStageSplits), 256, 0, st>>>(h->d_stage.p, h->P, n, h->d_partial.p);
    CUDA_CHECK(cudaGetLastError());
    finish_sums_kernel<<<(n * 4 + 127) / 128, 128, 0, st>>>(h->d_partial.p, n, dev_sums + done * 4);
    CUDA_CHECK(cudaGetLastError());
    h->last_launches += 2;
  }
  CUDA_CHECK(cudaEventRecord(h->ev1, st));
  for (int64_t s = 0; s < nsteps; ++s) h->time += dt;  // monte_carlo.cpp:341, one addition per step
}

void finish_step_host(cntmc_t* h, int64_t nsteps, double* msd_out) {
  std::vector<double> sums((size_t)nsteps * 4);
  h->d_sums.download(sums.data(), sums.size(), h->stream);
  unsigned long long ctrs[CTR_COUNT];
  h->d_counters.download(ctrs, CTR_COUNT, h->stream);
  check_flags(h);  // synchronises
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_ms = ms;
  h->deep_ms = 0;
  for (size_t k = 0; k + 1 < h->deep_events.size(); k += 2) {
    float dms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&dms, h->deep_events[k], h->deep_events[k + 1]));
    h->deep_ms += dms;
  }
  for (cudaEvent_t ev : h->deep_events) cudaEventDestroy(ev);
  h->deep_events.clear();
  h->kernel_ms = 0;
  h->kernel_launches = (int64_t)h->kernel_events.size() / 2;
  for (size_t k = 0; k + 1 < h->kernel_events.size(); k += 2) {
    float kms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&kms, h->kernel_events[k], h->kernel_events[k + 1]));
    h->kernel_ms += kms;
    cudaEventDestroy(h->kernel_events[k]);
    cudaEventDestroy(h->kernel_events[k + 1]);
  }
  h->kernel_events.clear();
  h->reinjections = (int64_t)ctrs[CTR_REINJECT];
  h->crossings = (int64_t)ctrs[CTR_CROSS];
  h->probes = (int64_t)ctrs[CTR_PROBE];
  for (int64_t s = 0; s < nsteps; ++s) {
    h->hops += (int64_t)sums[(size_t)s * 4 + 3];
    if (msd_out)
      for (int c = 0; c < 3; ++c) msd_out[s * 3 + c] = sums[(size_t)s * 4 + c] / double(h->P);  // monte_carlo.cpp:402-404
  }
}

template <typename F>
int guarded(const cntmc_t* h, F&& body) {
  try {
    body();
    return CNTMC_OK;
  } catch (const CudaError& e) {
    (h ? h->err : g_create_error) = e.what();
    return CNTMC_ERR_CUDA;
  } catch (const StateError& e) {
    (h ? h->err : g_create_error) = e.what();
    return CNTMC_ERR_STATE;
  } catch (const ReplayError& e) {
    (h ? h->err : g_create_error) = e.what();
    return CNTMC_ERR_REPLAY;
  } catch (const std::exception& e) {
    (h ? h->err : g_create_error) = e.what();
    return CNTMC_ERR_INVALID;
  }
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char* cntmc_version(void) { return "cntmc-b200 0.1 (sm_100a)"; }

const char* cntmc_last_error(const cntmc_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cntmc_create(const char* json_text, cntmc_t** out) {
  if (out) *out = nullptr;
  return guarded(nullptr, [&] {
    require(json_text != nullptr && out != nullptr, "null argument");
    std::unique_ptr<cntmc_t> h(new cntmc_t);
    const json::Value      doc = json::parse(json_text);
    h->block = mc_block(doc);
    h->prm = parse_params(h->block);
    if (const json::Value* cnts = doc.find("cnts") ? doc.find("cnts") : h->block.find("cnts")) h->prm.tubes = parse_tubes(*cnts);
    // the kernels divide by the velocity with div_by (hop_core.h), which is exact for every divisor but these
    if (div_by_unsafe(h->prm.velocity))
      throw std::invalid_argument("\"exciton velocity [m/s]\": a value whose binary significand is all ones is not supported");
    *out = h.release();
  });
}

void cntmc_destroy(cntmc_t* h) {
  if (!h) return;
  if (h->device >= 0) cudaSetDevice(h->device);
  delete h;
}

int cntmc_set_device(cntmc_t* h, int device) {
  return guarded(h, [&] {
    require(!h->initialised, "cntmc_set_device must precede initialisation");
    h->device = device;
  });
}
int cntmc_set_stream(cntmc_t* h, void* cuda_stream) {
  return guarded(h, [&] { h->stream = (cudaStream_t)cuda_stream; });
}

int cntmc_load_mesh(cntmc_t* h, const char* dir) {
  return guarded(h, [&] {
    const std::string d = dir ? std::string(dir) : h->prm.mesh_dir;
    require(!d.empty(), "no mesh directory given");
    h->mesh = load_mesh(expand_home(d));
    h->have_mesh = true;
  });
}

int cntmc_set_mesh(cntmc_t* h, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient) {
  return guarded(h, [&] {
    require(n_tubes > 0 && n_cols > 0 && pos_nm && orient, "bad mesh arguments");
    const size_t N = (size_t)(n_tubes * n_cols);
    h->mesh.n_tubes = n_tubes;
    h->mesh.n_cols = n_cols;
    for (int c = 0; c < 3; ++c) {
      h->mesh.pos[c].assign(pos_nm + c * N, pos_nm + (c + 1) * N);
      h->mesh.orient[c].assign(orient + c * N, orient + (c + 1) * N);
    }
    h->have_mesh = true;
  });
}

int cntmc_set_rate_table(cntmc_t* h, const int32_t dims[4], const double* theta, const double* z, const double* a1,
                         const double* a2, const double* rates) {
  return guarded(h, [&] {
    require(dims && theta && z && a1 && a2 && rates, "null argument");
    require(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[3] > 0, "table dimensions must be positive");
    require(!h->initialised, "cntmc_set_rate_table must precede initialisation");
    h->table.theta.assign(theta, theta + dims[0]);
    h->table.z.assign(z, z + dims[1]);
    h->table.a1.assign(a1, a1 + dims[2]);
    h->table.a2.assign(a2, a2 + dims[3]);
    h->table.rates.assign(rates, rates + (size_t)dims[0] * dims[1] * dims[2] * dims[3]);
  });
}

int cntmc_get_rate_table_dims(const cntmc_t* h, int32_t dims[4]) {
  return guarded(h, [&] {
    require(!h->table.empty(), "no rate table yet");
    dims[0] = (int32_t)h->table.theta.size();
    dims[1] = (int32_t)h->table.z.size();
    dims[2] = (int32_t)h->table.a1.size();
    dims[3] = (int32_t)h->table.a2.size();
  });
}
int cntmc_get_rate_table(const cntmc_t* h, double* theta, double* z, double* a1, double* a2, double* rates) {
  return guarded(h, [&] {
    require(!h->table.empty(), "no rate table yet");
    const HostTable& t = h->table;
    if (theta) std::copy(t.theta.begin(), t.theta.end(), theta);
    if (z) std::copy(t.z.begin(), t.z.end(), z);
    if (a1) std::copy(t.a1.begin(), t.a1.end(), a1);
    if (a2) std::copy(t.a2.begin(), t.a2.end(), a2);
    if (rates) std::copy(t.rates.begin(), t.rates.end(), rates);
  });
}

int cntmc_save_rate_table(const cntmc_t* h, const char* dir) {
  return guarded(h, [&] {
    require(dir != nullptr, "null directory");
    require(!h->table.empty(), "no rate table yet");
    save_rate_table(h->table, expand_home(dir));
  });
}
int cntmc_load_rate_table(cntmc_t* h, const char* dir) {
  return guarded(h, [&] {
    require(dir != nullptr, "null directory");
    require(!h->initialised, "cntmc_load_rate_table must precede initialisation");
    h->table = load_rate_table(expand_home(dir));
  });
}

int cntmc_kubo_init(cntmc_t* h) {
  return guarded(h, [&] {
    h->contact_mode = false;
    common_init(h);
    h->inj = injection_region(h->sites, h->dom, h->prm.n_sections);
    h->d_inject.upload(h->inj.sites, h->stream);
    h->T.inject = h->d_inject.p;
    h->T.n_inject = (int32_t)h->inj.sites.size();
    for (int c = 0; c < 3; ++c) {
      h->T.rem_lo[c] = h->inj.rem_lo[c];
      h->T.rem_hi[c] = h->inj.rem_hi[c];
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->initialised = true;
  });
}

int cntmc_kubo_create_particles(cntmc_t* h, int64_t n_particles, uint64_t seed, uint64_t first_global_id) {
  return guarded(h, [&] {
    h->replay = false;
    h->draws = DrawConfig{};
    h->draws.seed = seed;
    h->draws.first_gid = first_global_id;
    create_common(h, n_particles > 0 ? n_particles : h->prm.n_particles);
  });
}

int cntmc_kubo_create_particles_replay(cntmc_t* h, int64_t P, const int64_t* offsets, const int32_t* draws_in,
                                       const double* logs) {
  return guarded(h, [&] {
    require(P > 0 && offsets && draws_in, "bad replay arguments");
    use_device(h);
    const size_t total = (size_t)offsets[P];
    h->r_off.upload(offsets, (size_t)P + 1, h->stream);
    h->r_draws.upload(draws_in, total, h->stream);
    if (logs) h->r_logs.upload(logs, total, h->stream);
    h->replay = true;
    h->draws = DrawConfig{};
    h->draws.replay_off = h->r_off.p;
    h->draws.replay_draws = h->r_draws.p;
    h->draws.replay_logs = logs ? h->r_logs.p : nullptr;
    create_common(h, P);
  });
}

int cntmc_kubo_step(cntmc_t* h, double dt, int64_t nsteps, double* msd_out) {
  return guarded(h, [&] {
    require(nsteps > 0, "nsteps must be positive");
    use_device(h);
    h->d_sums.alloc((size_t)nsteps * 4);
    kubo_step_device(h, dt, nsteps, h->d_sums.p);
    finish_step_host(h, nsteps, msd_out);
  });
}

int cntmc_kubo_step_dev(cntmc_t* h, double dt, int64_t nsteps, double* dev_sums) {
  return guarded(h, [&] {
    require(dev_sums != nullptr, "null device buffer");
    kubo_step_device(h, dt, nsteps, dev_sums);
  });
}

// a slice of the parent's simulation: same tables (borrowed pointers), own stream, exciton buffers, lists and staging
static cntmc_handle* make_slice(cntmc_t* h) {
  std::unique_ptr<cntmc_handle> s(new cntmc_handle);
  s->is_slice = true;
  s->prm = h->prm;
  s->device = h->device;
  CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreate(&s->ev0));
  CUDA_CHECK(cudaEventCreate(&s->ev1));
  s->sites.N = h->sites.N;
  s->T = h->T;
  s->initialised = true;
  s->contact_mode = false;
  s->sm_count = h->sm_count;
  s->d_flags.alloc(FLAG_COUNT);
  s->d_counters.alloc(CTR_COUNT);
  CUDA_CHECK(cudaMemsetAsync(s->d_flags.p, 0, FLAG_COUNT * sizeof(int32_t), s->stream));
  CUDA_CHECK(cudaMemsetAsync(s->d_counters.p, 0, CTR_COUNT * sizeof(unsigned long long), s->stream));
  h->slices.emplace_back(std::move(s));
  return h->slices.back().get();
}

int cntmc_kubo_step_host_state(cntmc_t* h, double dt, int64_t nsteps, int64_t P, int32_t* site, double* pos, double* delta,
                               double* ff, uint8_t* heading, uint32_t* ndraw, double* msd_out) {
  return guarded(h, [&] {
    require(h->initialised && !h->contact_mode, "call cntmc_kubo_init first");
    require(P > 0 && site && pos && delta && ff && heading && ndraw, "bad host-state arguments");
    require(nsteps > 0, "nsteps must be positive");
    use_device(h);
    cudaStream_t st = h->stream;
    // the kernels index the site tables with what the caller hands in: refuse what would read out of bounds
    const int64_t N = h->sites.N;
    unsigned bad_site = 0, bad_ff = 0;  // branch-free so that the compiler vectorises it (0.1 ms per 1e6 excitons)
    for (int64_t i = 0; i < P; ++i) {
      bad_site |= (unsigned)((uint32_t)site[i] >= (uint32_t)N);
      bad_ff |= (unsigned)!(ff[i] - ff[i] == 0.0);
    }
    if (bad_site) throw std::invalid_argument("host state: site index out of range");
    if (bad_ff) throw std::invalid_argument("host state: free-flight time is not finite");
    const size_t n = (size_t)P;
    const int    K = (int)std::max<int64_t>(1, std::min<int64_t>(h->opt_host_slices, P / 65536));
    if (K > 1 && h->trace_cap == 0 && !h->opt_stats && !h->opt_time_kernels) {
      // ---- sliced: slice k = excitons [off_k, off_k + n_k) on its own stream; H2D of slice k+1 and D2H of slice k-1 overlap
      // the kernels of slice k.  The streams are keyed by global id, so slicing changes no trajectory; the ensemble rows are
      // the sum of the slices' rows (summation order differs from the unsliced call in the last bits only).
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (h->sm_count == 0) {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
      }
      while ((int)h->slices.size() < K) make_slice(h);
      h->P = P;  // (the blocks per SM follow the whole population, occupancy_of)
      std::vector<std::vector<double>> rows((size_t)K, std::vector<double>((size_t)nsteps * 4));
      for (int k = 0; k < K; ++k) {
        cntmc_handle* s = h->slices[(size_t)k].get();
        const size_t  off = n * (size_t)k / (size_t)K, cnt = n * (size_t)(k + 1) / (size_t)K - off;
        s->T = h->T;
        s->opt_chunk = h->opt_chunk; s->opt_hot_pct = h->opt_hot_pct; s->opt_occupancy = occupancy_of(h);
        s->opt_top_entries = h->opt_top_entries; s->opt_deep_thr = h->opt_deep_thr; s->opt_deep_blocks = h->opt_deep_blocks;
        s->opt_deep_rounds = h->opt_deep_rounds; s->opt_trap_burst = h->opt_trap_burst;
        s->opt_deep_overlap = h->opt_deep_overlap; s->opt_overlap_trap_blocks = h->opt_overlap_trap_blocks;
        if (s->opt_stage_mb <= 0 && h->opt_stage_mb > 0) s->opt_stage_mb = std::max<int64_t>(64, h->opt_stage_mb / K);
        s->grid_share = std::max<int64_t>(1, std::min<int64_t>(K, h->opt_slice_share > 0 ? h->opt_slice_share : K));
        s->replay = h->replay;
        s->draws = h->draws;
        s->draws.first_gid = h->draws.first_gid + (uint64_t)off;
        cudaStream_t ss = s->stream;
        if ((int64_t)cnt > s->capacity) s->alloc_excitons((int64_t)cnt);
        s->have_lists = false;
        s->P = (int64_t)cnt;
        s->ex.site.upload(site + off, cnt, ss);
        s->ex.px.upload(pos + off, cnt, ss); s->ex.py.upload(pos + n + off, cnt, ss); s->ex.pz.upload(pos + 2 * n + off, cnt, ss);
        s->ex.dx.upload(delta + off, cnt, ss); s->ex.dy.upload(delta + n + off, cnt, ss); s->ex.dz.upload(delta + 2 * n + off, cnt, ss);
        s->ex.ff.upload(ff + off, cnt, ss);
        s->ex.heading.upload(heading + off, cnt, ss);
        s->ex.ndraw.upload(ndraw + off, cnt, ss);
        s->d_sums.alloc((size_t)nsteps * 4);
        kubo_step_device(s, dt, nsteps, s->d_sums.p);
        s->ex.site.download(site + off, cnt, ss);
        s->ex.px.download(pos + off, cnt, ss); s->ex.py.download(pos + n + off, cnt, ss); s->ex.pz.download(pos + 2 * n + off, cnt, ss);
        s->ex.dx.download(delta + off, cnt, ss); s->ex.dy.download(delta + n + off, cnt, ss); s->ex.dz.download(delta + 2 * n + off, cnt, ss);
        s->ex.ff.download(ff + off, cnt, ss);
        s->ex.heading.download(heading + off, cnt, ss);
        s->ex.ndraw.download(ndraw + off, cnt, ss);
      }  // (nothing above may block the host: a copy to pageable memory would serialise the slices)
      double ms = 0;
      int64_t launches = 0;
      for (int k = 0; k < K; ++k) {
        cntmc_handle*      s = h->slices[(size_t)k].get();
        unsigned long long ctrs[CTR_COUNT];
        s->d_sums.download(rows[(size_t)k].data(), rows[(size_t)k].size(), s->stream);
        s->d_counters.download(ctrs, CTR_COUNT, s->stream);
        check_flags(s);  // synchronises the slice's stream
        float sms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&sms, s->ev0, s->ev1));
        ms = std::max<double>(ms, sms);
        launches += s->last_launches;
        h->reinjections += (int64_t)ctrs[CTR_REINJECT];
        CUDA_CHECK(cudaMemsetAsync(s->d_counters.p, 0, CTR_COUNT * sizeof(unsigned long long), s->stream));
      }
      h->last_ms = ms;
      h->last_launches = launches;
      h->P = P;
      h->have_lists = false;
      for (int64_t q = 0; q < nsteps; ++q) {
        h->time += dt;  // monte_carlo.cpp:341, one addition per step
        double v[4] = {0, 0, 0, 0};
        for (int k = 0; k < K; ++k)
          for (int c = 0; c < 4; ++c) v[c] += rows[(size_t)k][(size_t)q * 4 + c];
        h->hops += (int64_t)v[3];
        if (msd_out)
          for (int c = 0; c < 3; ++c) msd_out[q * 3 + c] = v[c] / double(P);  // monte_carlo.cpp:402-404
      }
      return;
    }
    if (P > h->capacity) h->alloc_excitons(P);
    if (P != h->P) h->trace_cap = 0;  // the trace buffers were sized for the previous population
    h->have_lists = false;  // the uploaded population has not been filed under activity classes yet
    h->P = P;
    h->ex.site.upload(site, n, st);
    h->ex.px.upload(pos, n, st); h->ex.py.upload(pos + n, n, st); h->ex.pz.upload(pos + 2 * n, n, st);
    h->ex.dx.upload(delta, n, st); h->ex.dy.upload(delta + n, n, st); h->ex.dz.upload(delta + 2 * n, n, st);
    h->ex.ff.upload(ff, n, st);
    h->ex.heading.upload(heading, n, st);
    h->ex.ndraw.upload(ndraw, n, st);
    h->d_sums.alloc((size_t)nsteps * 4);
    kubo_step_device(h, dt, nsteps, h->d_sums.p);
    h->ex.site.download(site, n, st);
    h->ex.px.download(pos, n, st); h->ex.py.download(pos + n, n, st); h->ex.pz.download(pos + 2 * n, n, st);
    h->ex.dx.download(delta, n, st); h->ex.dy.download(delta + n, n, st); h->ex.dz.download(delta + 2 * n, n, st);
    h->ex.ff.download(ff, n, st);
    h->ex.heading.download(heading, n, st);
    h->ex.ndraw.download(ndraw, n, st);
    finish_step_host(h, nsteps, msd_out);
  });
}

double  cntmc_time(const cntmc_t* h) { return h->time; }
double  cntmc_kubo_max_time(const cntmc_t* h) { return h->prm.max_time; }
double  cntmc_time_step(const cntmc_t* h) { return h->prm.time_step; }
int64_t cntmc_number_of_particles(const cntmc_t* h) { return h->P; }
int64_t cntmc_hops(const cntmc_t* h) { return h->hops; }
int64_t cntmc_reinjections(const cntmc_t* h) { return h->reinjections; }
int64_t cntmc_crossings(const cntmc_t* h) { return h->crossings; }
int64_t cntmc_probes(const cntmc_t* h) { return h->probes; }

// ---- contact flavour ----------------------------------------------------------------------------------------------------
static int contact_init(cntmc_t* h, int64_t c1_pop, int64_t c2_pop, uint64_t seed, int64_t capacity, int64_t n_ids,
                        const int64_t* offsets, const int32_t* draws_in, const double* logs) {
  return guarded(h, [&] {
    require(h->prm.n_seg >= 2, "\"number of segments\" must be at least 2 in contact mode");
    require(c1_pop >= 0 && c2_pop >= 0, "contact populations must be non-negative");
    h->contact_mode = true;
    common_init(h);
    const int n_seg = h->n_seg = h->prm.n_seg;
    cudaStream_t st = h->stream;
    h->area = slab_areas(h->sites, h->dom, n_seg);                      // monte_carlo.h:184
    h->c1_sites = contact_sites(h->sites, h->dom, n_seg, 1);            // monte_carlo.h:188
    h->c2_sites = contact_sites(h->sites, h->dom, n_seg, n_seg);        // monte_carlo.h:189
    if ((c1_pop > 0 && h->c1_sites.empty()) || (c2_pop > 0 && h->c2_sites.empty()))
      throw StateError("a contact with a non-zero population holds no site (rand() % 0 in the reference, monte_carlo.h:477)");
    h->d_c1.upload(h->c1_sites, st);
    h->d_c2.upload(h->c2_sites, st);
    h->c1_pop = c1_pop;
    h->c2_pop = c2_pop;
    h->replay = offsets != nullptr;
    h->draws = DrawConfig{};
    h->draws.seed = seed;
    h->draws.first_gid = (uint64_t)h->opt_gid_base << 56;  // several handles of one simulation keep their stream ids apart
    if (h->replay) {  // recorded draws per exciton id (ids in order of birth, as the kernels number them)
      require(n_ids > 0 && draws_in != nullptr, "bad replay arguments");
      use_device(h);
      const size_t total = (size_t)offsets[n_ids];
      h->r_off.upload(offsets, (size_t)n_ids + 1, st);
      h->r_draws.upload(draws_in, total, st);
      if (logs) h->r_logs.upload(logs, total, st);
      h->draws.first_gid = 0;
      h->draws.replay_off = h->r_off.p;
      h->draws.replay_draws = h->r_draws.p;
      h->draws.replay_logs = logs ? h->r_logs.p : nullptr;
      h->replay_ids = n_ids;
    }
    // create_particles (monte_carlo.h:274-316): linear profile over the slabs, sites from the half-open slab lists
    const double         dp = double(c2_pop - c1_pop) / (double(n_seg) - 1);
    std::vector<int64_t> count_off((size_t)n_seg + 1, 0), site_off((size_t)n_seg + 1, 0);
    std::vector<int32_t> slab_sites;
    for (int i = 0; i < n_seg; ++i) {
      const int64_t n_particle = (int64_t)std::round(double(c1_pop) + double(i) * dp);
      const auto    list = slab_sites_half_open(h->sites, h->dom, n_seg, i);
      if (n_particle > 0 && list.empty())
        throw StateError("a slab that must receive excitons holds no site (rand() % 0 in the reference, monte_carlo.h:305)");
      slab_sites.insert(slab_sites.end(), list.begin(), list.end());
      site_off[(size_t)i + 1] = (int64_t)slab_sites.size();
      count_off[(size_t)i + 1] = count_off[(size_t)i] + std::max<int64_t>(n_particle, 0);
    }
    const int64_t P0 = count_off[(size_t)n_seg];
    h->capacity = 0;
    h->reserve_contact(std::max<int64_t>(capacity, P0 + 1), 0);
    h->P = P0;
    h->next_gid = h->draws.first_gid + (uint64_t)P0;
    h->time = 0;
    h->hops = h->reinjections = h->crossings = h->probes = 0;
    if (P0 > 0) {
      DevBuf<int64_t> d_count_off, d_site_off;
      DevBuf<int32_t> d_slab_sites;
      d_count_off.upload(count_off, st);
      d_site_off.upload(site_off, st);
      d_slab_sites.upload(slab_sites, st);
      ContactCreateArgs a{};
      a.T = h->T;
      a.S = h->arrays();
      a.draws = h->draws;
      a.P = P0;
      a.count_off = d_count_off.p;
      a.site_off = d_site_off.p;
      a.slab_sites = d_slab_sites.p;
      a.n_seg = n_seg;
      a.alive = h->d_alive.p;
      a.flags = h->d_flags.p;
      if (h->replay)
        create_contact_population_kernel<ReplayDraws><<<(unsigned)((P0 + 255) / 256), 256, 0, st>>>(a);
      else
        create_contact_population_kernel<PhiloxDraws><<<(unsigned)((P0 + 255) / 256), 256, 0, st>>>(a);
      CUDA_CHECK(cudaGetLastError());
      check_flags(h);  // synchronises before the temporaries go away
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    h->initialised = true;
  });
}

int cntmc_init(cntmc_t* h, int64_t c1_pop, int64_t c2_pop, uint64_t seed, int64_t capacity) {
  return contact_init(h, c1_pop, c2_pop, seed, capacity, 0, nullptr, nullptr, nullptr);
}
int cntmc_init_replay(cntmc_t* h, int64_t c1_pop, int64_t c2_pop, int64_t n_ids, const int64_t* offsets, const int32_t* draws,
                      const double* logs) {
  if (!offsets || !draws || n_ids <= 0) {
    h->err = "bad replay arguments";
    return CNTMC_ERR_INVALID;
  }
  return contact_init(h, c1_pop, c2_pop, 0, 0, n_ids, offsets, draws, logs);
}

// nsteps iterations of { step ; metrics ; repopulate_contacts } -> dev_bins[nsteps][2*n_seg-1] (device, 64-bit counts)
static void contact_step_device(cntmc_t* h, double dt, int64_t nsteps, unsigned long long* dev_bins) {
  require(h->initialised && h->contact_mode, "call cntmc_init first");
  require(nsteps > 0, "nsteps must be positive");
  use_device(h);
  cudaStream_t st = h->stream;
  if (h->sm_count == 0) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int     nb = 2 * h->n_seg - 1;
  const int64_t C = h->c1_pop + h->c2_pop;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>({h->opt_chunk, nsteps, (int64_t)(40000 / (nb * 4))}));
  DevBuf<unsigned long long> d_count;
  d_count.alloc(1);
  h->last_launches = 0;
  CUDA_CHECK(cudaEventRecord(h->ev0, st));
  CUDA_CHECK(cudaMemsetAsync(dev_bins, 0, (size_t)nsteps * nb * sizeof(unsigned long long), st));
  for (int64_t done = 0; done < nsteps; done += chunk) {
    const int     n = (int)std::min(chunk, nsteps - done);
    const int64_t W = h->P + (int64_t)n * C;
    if (W == 0) {
      for (int s = 0; s < n; ++s) h->time += dt;
      continue;
    }
    require(W < (1ll << 32), "more than 2^32 excitons in one contact launch: lower chunk_steps");  // work items are 32-bit indices
    h->reserve_contact(W, h->P);
    const unsigned grid = (unsigned)std::min<int64_t>((W + 127) / 128, (int64_t)h->sm_count * occupancy_of(h));
    set_u64_kernel<<<1, 1, 0, st>>>(h->d_counters.p + CTR_QUEUE, (unsigned long long)grid * 128ull);
    ContactArgs a{};
    a.T = h->T;
    a.S = h->arrays();
    a.draws = h->draws;
    a.alive = h->d_alive.p;
    a.P_alive = h->P;
    a.c1_pop = h->c1_pop;
    a.c2_pop = h->c2_pop;
    a.c1_sites = h->d_c1.p;
    a.c2_sites = h->d_c2.p;
    a.n_c1 = (int32_t)h->c1_sites.size();
    a.n_c2 = (int32_t)h->c2_sites.size();
    a.next_gid = h->next_gid;
    a.dt = dt;
    a.nsteps = n;
    a.n_seg = h->n_seg;
    a.use_top = h->opt_top_entries ? 1 : 0;
    a.ymin = h->dom.lo[1];
    a.ymax = h->dom.hi[1];
    a.dy = (h->dom.hi[1] - h->dom.lo[1]) / double(h->n_seg);  // monte_carlo.h:446
    a.bins = dev_bins + done * nb;
    a.flags = h->d_flags.p;
    a.counters = h->d_counters.p;
    const size_t smem = (size_t)n * nb * sizeof(int);
    if (h->replay) {  // the reference's own rand() stream, split per exciton id: every id born in this call needs its list
      if ((int64_t)(h->next_gid + (uint64_t)((int64_t)n * C)) > h->replay_ids)
        throw ReplayError("the replayed draw lists do not cover the excitons this call creates");
      contact_kernel<ReplayDraws, 5><<<grid, 128, smem, st>>>(a);
    } else {
      switch (occupancy_of(h)) {
        case 4: contact_kernel<PhiloxDraws, 4><<<grid, 128, smem, st>>>(a); break;
        case 6: contact_kernel<PhiloxDraws, 6><<<grid, 128, smem, st>>>(a); break;
        case 7: contact_kernel<PhiloxDraws, 7><<<grid, 128, smem, st>>>(a); break;
        case 8: contact_kernel<PhiloxDraws, 8><<<grid, 128, smem, st>>>(a); break;
        default: contact_kernel<PhiloxDraws, 5><<<grid, 128, smem, st>>>(a); break;
      }
    }
    CUDA_CHECK(cudaGetLastError());
    // survivors, in work-item order, move to the front of the spare copy
    iota_kernel<<<(unsigned)((W + 255) / 256), 256, 0, st>>>(h->e_iota.p, W);
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, h->e_iota.p, h->d_alive.p, h->e_perm.p, d_count.p, (int64_t)W, st);
    h->sort_tmp.alloc(bytes);
    CUDA_CHECK(cub::DeviceSelect::Flagged(h->sort_tmp.p, bytes, h->e_iota.p, h->d_alive.p, h->e_perm.p, d_count.p, (int64_t)W, st));
    unsigned long long survivors = 0;
    d_count.download(&survivors, 1, st);
    check_flags(h);  // synchronises
    if (survivors > 0) {
      gather_excitons_kernel<<<(unsigned)((survivors + 255) / 256), 256, 0, st>>>(h->ex.view(true), h->ex_spare.view(true), h->e_perm.p,
                                                                                (int64_t)survivors);
      CUDA_CHECK(cudaGetLastError());
    }
    h->ex.swap(h->ex_spare);
    h->P = (int64_t)survivors;
    h->next_gid += (uint64_t)((int64_t)n * C);
    h->last_launches += 4;
    for (int s = 0; s < n; ++s) h->time += dt;  // monte_carlo.h:354
  }
  CUDA_CHECK(cudaEventRecord(h->ev1, st));
}

int cntmc_step_dev(cntmc_t* h, double dt, int64_t nsteps, int64_t* dev_bins) {
  return guarded(h, [&] {
    require(dev_bins != nullptr, "null device buffer");
    contact_step_device(h, dt, nsteps, reinterpret_cast<unsigned long long*>(dev_bins));
  });
}

int cntmc_step(cntmc_t* h, double dt, int64_t nsteps, int64_t* pop_out, int64_t* curr_out) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    use_device(h);
    const int                  nb = 2 * h->n_seg - 1;
    DevBuf<unsigned long long> d_bins;
    d_bins.alloc((size_t)nsteps * nb);
    contact_step_device(h, dt, nsteps, d_bins.p);
    std::vector<unsigned long long> bins((size_t)nsteps * nb);
    d_bins.download(bins.data(), bins.size(), h->stream);
    unsigned long long ctrs[CTR_COUNT];
    h->d_counters.download(ctrs, CTR_COUNT, h->stream);
    check_flags(h);
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->hops = (int64_t)ctrs[CTR_EVENTS];
    h->crossings = (int64_t)ctrs[CTR_CROSS];
    h->probes = (int64_t)ctrs[CTR_PROBE];
    for (int64_t s = 0; s < nsteps; ++s) {
      for (int i = 0; i < h->n_seg; ++i)
        if (pop_out) pop_out[s * h->n_seg + i] = (int64_t)bins[(size_t)s * nb + i];
      for (int i = 0; i + 1 < h->n_seg; ++i)
        if (curr_out) curr_out[s * (h->n_seg - 1) + i] = (int64_t)bins[(size_t)s * nb + h->n_seg + i];
    }
  });
}

int cntmc_get_area(const cntmc_t* h, double* area) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    std::copy(h->area.begin(), h->area.end(), area);
  });
}
int cntmc_num_contact_sites(const cntmc_t* h, int which, int64_t* n) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    require(which == 1 || which == 2, "contact must be 1 or 2");
    *n = (int64_t)(which == 1 ? h->c1_sites.size() : h->c2_sites.size());
  });
}
int cntmc_get_contact_sites(const cntmc_t* h, int which, int32_t* ids) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    require(which == 1 || which == 2, "contact must be 1 or 2");
    const auto& l = which == 1 ? h->c1_sites : h->c2_sites;
    std::copy(l.begin(), l.end(), ids);
  });
}
int cntmc_number_of_segments(const cntmc_t* h) { return h->prm.n_seg; }
int cntmc_get_scatterer_statistics(const cntmc_t* h, int64_t* pop) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    require(pop != nullptr, "null argument");
    const auto c = slab_site_counts(h->sites, h->dom, h->n_seg);
    std::copy(c.begin(), c.end(), pop);
  });
}

int cntmc_track_particle(cntmc_t* h, double dt, uint64_t seed, uint64_t global_id, int64_t n_replay,
                         const int32_t* replay_draws, const double* replay_logs, int64_t max_steps, double* path,
                         int64_t* n_steps, int32_t* reached) {
  return guarded(h, [&] {
    require(h->initialised && h->contact_mode, "call cntmc_init first");
    require(dt > 0 && max_steps > 0 && path && n_steps, "bad track_particle arguments");
    require(n_replay == 0 || replay_draws != nullptr, "bad replay arguments");
    if (h->c1_sites.empty()) throw StateError("the first contact holds no site (rand() % 0 in the reference, monte_carlo.h:477)");
    use_device(h);
    cudaStream_t     st = h->stream;
    DevBuf<double>   d_path, d_logs;
    DevBuf<int64_t>  d_out, d_off;
    DevBuf<int32_t>  d_draws;
    d_path.alloc((size_t)max_steps * 3);
    d_out.alloc(3);
    TrackArgs a{};
    a.T = h->T;
    a.draws.seed = seed;
    a.draws.first_gid = global_id;
    if (n_replay > 0) {
      const int64_t off[2] = {0, n_replay};
      d_off.upload(off, 2, st);
      d_draws.upload(replay_draws, (size_t)n_replay, st);
      if (replay_logs) d_logs.upload(replay_logs, (size_t)n_replay, st);
      a.draws.replay_off = d_off.p;
      a.draws.replay_draws = d_draws.p;
      a.draws.replay_logs = replay_logs ? d_logs.p : nullptr;
    }
    a.c1_sites = h->d_c1.p;
    a.n_c1 = (int32_t)h->c1_sites.size();
    a.dt = dt;
    const double ymin = h->dom.lo[1], ymax = h->dom.hi[1];
    const double dy = (ymax - ymin) / double(h->n_seg);
    a.y_stop = ymin + double(h->n_seg - 1) * dy;  // monte_carlo.h:809
    a.max_steps = max_steps;
    a.path = d_path.p;
    a.n_out = d_out.p;
    a.flags = h->d_flags.p;
    if (n_replay > 0)
      track_kernel<ReplayDraws><<<1, 32, 0, st>>>(a);
    else
      track_kernel<PhiloxDraws><<<1, 32, 0, st>>>(a);
    CUDA_CHECK(cudaGetLastError());
    int64_t out[3] = {0, 0, 0};
    d_out.download(out, 3, st);
    check_flags(h);  // synchronises
    d_path.download(path, (size_t)out[0] * 3, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    *n_steps = out[0];
    if (reached) *reached = (int32_t)out[1];
  });
}

// ---- read-back -----------------------------------------------------------------------------------------------------------
int cntmc_num_sites(const cntmc_t* h, int64_t* n) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    *n = h->sites.N;
  });
}

int cntmc_get_sites(const cntmc_t* h, double* pos, double* orient, int32_t* left, int32_t* right, double* max_rate,
                    double* inv_max_rate) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    const size_t N = (size_t)h->sites.N;
    for (int c = 0; c < 3; ++c) {
      if (pos) std::copy(h->sites.pos[c].begin(), h->sites.pos[c].end(), pos + c * N);
      if (orient) std::copy(h->sites.orient[c].begin(), h->sites.orient[c].end(), orient + c * N);
    }
    if (left) std::copy(h->sites.left.begin(), h->sites.left.end(), left);
    if (right) std::copy(h->sites.right.begin(), h->sites.right.end(), right);
    if (max_rate || inv_max_rate) need_rates(h);
    if (max_rate) std::copy(h->max_rate.begin(), h->max_rate.end(), max_rate);
    if (inv_max_rate) std::copy(h->inv_max_rate.begin(), h->inv_max_rate.end(), inv_max_rate);
  });
}

int cntmc_get_domain(const cntmc_t* h, double d[6]) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    for (int c = 0; c < 3; ++c) {
      d[c] = h->dom.lo[c];
      d[3 + c] = h->dom.hi[c];
    }
  });
}
int cntmc_get_removal_domain(const cntmc_t* h, double d[6]) {
  return guarded(h, [&] {
    require(h->initialised && !h->contact_mode, "not initialised in Green-Kubo mode");
    for (int c = 0; c < 3; ++c) {
      d[c] = h->inj.rem_lo[c];
      d[3 + c] = h->inj.rem_hi[c];
    }
  });
}
int cntmc_num_inject(const cntmc_t* h, int64_t* n) {
  return guarded(h, [&] {
    require(h->initialised && !h->contact_mode, "not initialised in Green-Kubo mode");
    *n = (int64_t)h->inj.sites.size();
  });
}
int cntmc_get_inject(const cntmc_t* h, int32_t* ids) {
  return guarded(h, [&] {
    require(h->initialised && !h->contact_mode, "not initialised in Green-Kubo mode");
    std::copy(h->inj.sites.begin(), h->inj.sites.end(), ids);
  });
}

int cntmc_csr_nnz(const cntmc_t* h, int64_t* nnz) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    *nnz = (int64_t)h->nnz;
  });
}
int cntmc_get_csr(const cntmc_t* h, int64_t* row_ptr, int32_t* nbr, double* cum) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    use_device(h);
    const size_t nnz = (size_t)h->nnz;
    if (row_ptr) {
      need_row_ptr(h);
      for (size_t i = 0; i < h->row_ptr.size(); ++i) row_ptr[i] = (int64_t)h->row_ptr[i];
    }
    if (nbr || cum) {
      std::vector<RowEntry> rows(nnz);
      h->d_row.download(rows.data(), nnz, h->stream);
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      for (size_t k = 0; k < nnz; ++k) {
        if (nbr) nbr[k] = rows[k].nbr;
        if (cum) cum[k] = rows[k].cum;
      }
    }
  });
}
int cntmc_get_csr_row(const cntmc_t* h, int64_t site, int64_t cap, int32_t* nbr, double* cum, int64_t* len) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    require(site >= 0 && site < h->sites.N && len != nullptr, "bad row request");
    use_device(h);
    uint64_t be[2];  // the row's bounds straight from the device: no 40 MB mirror for a handful of rows
    CUDA_CHECK(cudaMemcpy(be, h->d_row_begin.p + site, sizeof be, cudaMemcpyDeviceToHost));
    const uint64_t b = be[0], d = be[1] - b;
    *len = (int64_t)d;
    const size_t n = (size_t)std::min<int64_t>((int64_t)d, cap);
    if (n && (nbr || cum)) {
      std::vector<RowEntry> row(n);
      CUDA_CHECK(cudaMemcpyAsync(row.data(), h->d_row.p + b, n * sizeof(RowEntry), cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      for (size_t k = 0; k < n; ++k) {
        if (nbr) nbr[k] = row[k].nbr;
        if (cum) cum[k] = row[k].cum;
      }
    }
  });
}
int64_t cntmc_csr_midpoint_guards(const cntmc_t* h) { return h->midpoint_guards; }
double  cntmc_csr_build_seconds(const cntmc_t* h) { return h->csr_seconds; }

int cntmc_get_particles(const cntmc_t* h, int32_t* site, double* pos, double* delta, double* ff, uint8_t* heading,
                        uint32_t* ndraw) {
  return guarded(h, [&] {
    require(h->P > 0, "no excitons");
    use_device(h);
    const size_t n = (size_t)h->P;
    cudaStream_t st = h->stream;
    if (site) h->ex.site.download(site, n, st);
    if (pos) {
      h->ex.px.download(pos, n, st);
      h->ex.py.download(pos + n, n, st);
      h->ex.pz.download(pos + 2 * n, n, st);
    }
    if (delta) {
      h->ex.dx.download(delta, n, st);
      h->ex.dy.download(delta + n, n, st);
      h->ex.dz.download(delta + 2 * n, n, st);
    }
    if (ff) h->ex.ff.download(ff, n, st);
    if (heading) h->ex.heading.download(heading, n, st);
    if (ndraw) h->ex.ndraw.download(ndraw, n, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int cntmc_get_gids(const cntmc_t* h, uint64_t* gid) {
  return guarded(h, [&] {
    require(h->contact_mode && h->P > 0 && gid != nullptr, "exciton ids exist in contact mode only");
    use_device(h);
    h->ex.gid.download(gid, (size_t)h->P, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}

int cntmc_trace_enable(cntmc_t* h, int32_t cap) {
  return guarded(h, [&] {
    require(h->P > 0, "no excitons");
    use_device(h);
    h->trace_cap = cap;
    if (cap > 0) {
      h->d_trace_sites.alloc((size_t)h->P * cap);
      h->d_trace_counts.alloc((size_t)h->P);
      CUDA_CHECK(cudaMemsetAsync(h->d_trace_counts.p, 0, (size_t)h->P * sizeof(int32_t), h->stream));
      // unused slots read back as -1 (cntmc_trace_get copies the whole [P][cap] array)
      CUDA_CHECK(cudaMemsetAsync(h->d_trace_sites.p, 0xff, (size_t)h->P * cap * sizeof(int32_t), h->stream));
    }
  });
}
int cntmc_trace_get(const cntmc_t* h, int32_t* counts, int32_t* sites_out) {
  return guarded(h, [&] {
    require(h->trace_cap > 0, "tracing is off");
    use_device(h);
    if (counts) h->d_trace_counts.download(counts, (size_t)h->P, h->stream);
    if (sites_out) h->d_trace_sites.download(sites_out, (size_t)h->P * h->trace_cap, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}

int cntmc_set_option(cntmc_t* h, const char* name, int64_t value) {
  return guarded(h, [&] {
    const std::string k = name ? name : "";
    if (k == "chunk_steps") {
      require(value >= 1 && value <= 16384, "chunk_steps must be in [1, 16384]");
      h->opt_chunk = value;
    } else if (k == "gid_base_shift56") {
      require(value >= 0 && value < 128, "gid_base_shift56 must be in [0, 128)");
      h->opt_gid_base = value;
    } else if (k == "deep_thr") {
      require(value >= 0 && value <= 1000000, "deep_thr must be in [0, 1e6] (0 = no group solver)");
      h->opt_deep_thr = value;
    } else if (k == "slice_share") {
      require(value >= 0 && value <= 16, "slice_share must be in [0, 16]");
      h->opt_slice_share = value;
    } else if (k == "csr_warp") {
      require(!h->initialised, "csr_warp must precede initialisation");
      h->opt_csr_warp = value ? 1 : 0;
    } else if (k == "guard_ppb") {
      require(value >= 1 && value <= 1000000000, "guard_ppb must be in [1, 1e9]");
      require(!h->initialised, "guard_ppb must precede initialisation");
      h->opt_guard_ppb = value;
    } else if (k == "host_slices") {
      require(value >= 1 && value <= 16, "host_slices must be in [1, 16]");
      h->opt_host_slices = value;
    } else if (k == "deep_group") {
      require(value == 1 || value == 8, "deep_group must be 1 or 8");
      h->opt_deep_group = value;
    } else if (k == "trap_burst") {
      require(value >= 1 && value <= 4096, "trap_burst must be in [1, 4096]");
      h->opt_trap_burst = value;
    } else if (k == "deep_overlap") {
      require(value == 0 || value == 1, "deep_overlap must be 0 or 1");
      h->opt_deep_overlap = value;
    } else if (k == "overlap_trap_blocks") {
      require(value >= 1 && value <= 4, "overlap_trap_blocks must be in [1, 4]");
      h->opt_overlap_trap_blocks = value;
    } else if (k == "deep_rounds") {
      require(value >= 1 && value <= 2, "deep_rounds must be 1 or 2");
      h->opt_deep_rounds = value;
    } else if (k == "deep_blocks") {
      require(value >= 1 && value <= 4, "deep_blocks must be 1 to 4 blocks per SM");
      h->opt_deep_blocks = value;
    } else if (k == "hot_pct") {
      require(value >= 0 && value <= 100, "hot_pct must be in [0, 100]");
      h->opt_hot_pct = value;
    } else if (k == "occupancy") {
      require(value == 0 || (value >= 4 && value <= 8), "occupancy must be 4 to 8 blocks per SM, or 0 (by table size)");
      h->opt_occupancy = value;
    } else if (k == "dirs") {
      h->opt_dirs = value ? 1 : 0;
      if (h->initialised) h->T.dir = h->opt_dirs ? h->d_dir.p : nullptr;
    } else if (k == "runs") {
      h->opt_runs = value ? 1 : 0;
      if (h->initialised) h->T.seg = h->opt_runs ? h->d_seg.p + kSegPad : nullptr;
    } else if (k == "top_entries") {
      h->opt_top_entries = value ? 1 : 0;
    } else if (k == "stage_mb") {
      require(value >= 0, "stage_mb must not be negative (0 = a third of the free device memory)");
      h->opt_stage_mb = value;
    } else if (k == "stats") {
      h->opt_stats = value ? 1 : 0;
    } else if (k == "time_kernels") {
      h->opt_time_kernels = value ? 1 : 0;
    } else {
      throw std::invalid_argument("unknown option \"" + k + "\"");
    }
  });
}
int64_t cntmc_get_option(const cntmc_t* h, const char* name) {
  const std::string k = name ? name : "";
  if (k == "chunk_steps") return h->opt_chunk;
  if (k == "occupancy") return h->initialised ? occupancy_of(const_cast<cntmc_t*>(h)) : h->opt_occupancy;  // 0 = by table size, resolved by init
  if (k == "hot_pct") return h->opt_hot_pct;
  if (k == "deep_thr") return h->opt_deep_thr;
  if (k == "deep_blocks") return h->opt_deep_blocks;
  if (k == "deep_rounds") return h->opt_deep_rounds;
  if (k == "deep_overlap") return h->opt_deep_overlap;
  if (k == "overlap_trap_blocks") return h->opt_overlap_trap_blocks;
  if (k == "trap_burst") return h->opt_trap_burst;
  if (k == "deep_group") return h->opt_deep_group;
  if (k == "host_slices") return h->opt_host_slices;
  if (k == "guard_ppb") return h->opt_guard_ppb;
  if (k == "csr_warp") return h->opt_csr_warp;
  if (k == "dbg_last_chunk") return h->last_chunk;
  if (k == "dbg_midpoint_repairs") return h->midpoint_repairs;
  if (k == "dbg_midpoint_changed") return h->midpoint_changed;
  if (k == "top_entries") return h->opt_top_entries;
  if (k == "runs") return h->opt_runs;
  if (k == "dirs") return h->opt_dirs;
  if (k == "stage_mb") return h->opt_stage_mb;
  if (k.rfind("dbg_", 0) == 0) {  // raw device counters of the last instrumented hop-kernel launch (option "stats")
    unsigned long long ctrs[CTR_COUNT];
    if (!h->d_counters.p || cudaMemcpy(ctrs, h->d_counters.p, sizeof ctrs, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (k == "dbg_warp_ns") return (int64_t)ctrs[CTR_WARP_NS];
    if (k == "dbg_span_ns") return (int64_t)(ctrs[CTR_T_LAST] - ~ctrs[CTR_T_FIRST_INV]);
    if (k == "dbg_warps") return (int64_t)ctrs[CTR_WARPS];
    if (k == "dbg_lane_busy") return (int64_t)ctrs[CTR_LANE_BUSY];
    if (k == "dbg_lane_idle") return (int64_t)ctrs[CTR_LANE_IDLE];
    if (k == "dbg_top_events") return (int64_t)ctrs[CTR_FAST];
    if (k == "dbg_walk_events") return (int64_t)ctrs[CTR_WALK];
    if (k == "dbg_deep_us") return (int64_t)(h->deep_ms * 1e3);
    if (k == "dbg_returned") {
      uint32_t v2[2] = {0, 0};
      if (!h->d_hand_count.p || cudaMemcpy(v2, h->d_hand_count.p, sizeof v2, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
      return v2[1];
    }
    if (k == "dbg_deferred" || k == "dbg_class4") {  // sizes of the trap solver's two lists in the last launch
      uint32_t v[kClasses] = {0};
      if (k == "dbg_deferred") {  // of the last round
        if (!h->d_hand_count.p || cudaMemcpy(v, h->d_hand_count.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return v[0];
      }
      if (!h->d_list_count[1 - h->cur_list].p ||
          cudaMemcpy(v, h->d_list_count[1 - h->cur_list].p, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
      return v[4];
    }
  }
  return -1;
}
double  cntmc_last_step_ms(const cntmc_t* h) { return h->last_ms; }
int64_t cntmc_last_step_launches(const cntmc_t* h) { return h->last_launches; }
int cntmc_sync(cntmc_t* h) {
  return guarded(h, [&] {
    require(h->initialised, "not initialised");
    use_device(h);
    unsigned long long ctrs[CTR_COUNT];
    h->d_counters.download(ctrs, CTR_COUNT, h->stream);
    check_flags(h);  // synchronises the stream, raises what the asynchronous calls could not report
    if (h->contact_mode) h->hops = (int64_t)ctrs[CTR_EVENTS];
    h->reinjections = (int64_t)ctrs[CTR_REINJECT];
    h->crossings = (int64_t)ctrs[CTR_CROSS];
    h->probes = (int64_t)ctrs[CTR_PROBE];
  });
}
int cntmc_dbg_walk_arith(int device, int64_t n, uint64_t seed, int64_t counts[4]) {
  try {
    if (counts == nullptr || n < 0) throw std::invalid_argument("bad argument");
    CUDA_CHECK(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
    const int     blocks = 148 * 8, threads = 256;
    const int64_t per_thread = (n + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads);
    walk_arith_kernel<<<blocks, threads>>>(per_thread, seed, d);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    unsigned long long hc[4] = {0, 0, 0, 0};
    if (err == cudaSuccess) err = cudaMemcpy(hc, d, sizeof hc, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (err != cudaSuccess) throw std::runtime_error(cudaGetErrorString(err));
    for (int k = 0; k < 4; ++k) counts[k] = (int64_t)hc[k];
    return CNTMC_OK;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return CNTMC_ERR_CUDA;
  }
}
double  cntmc_last_kernel_ms(const cntmc_t* h) { return h->kernel_ms; }
int64_t cntmc_last_kernel_launches(const cntmc_t* h) { return h->kernel_launches; }

}  // extern "C"
