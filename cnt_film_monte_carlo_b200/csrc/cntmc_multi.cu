// cntmc_multi.cu -- one simulation on several GPUs of one box, behind the same C ABI (include/cntmc.h, cntmc_multi_*).
//
// The reference is one process with an OpenMP team over the particle list (monte_carlo.cpp:320-338, monte_carlo.h:
// 345-351).  Here one host thread drives one cntmc handle per GPU: the read-only tables are replicated (every GPU builds
// its own neighbour table), the exciton population is split by global id -- shard r of W owns [first_r, first_r + n_r),
// and an exciton's counter-based stream is keyed by its global id, so every trajectory is the same bits for any W --
// and the one exchange per call is an NCCL all-reduce (sum) of the per-step rows [sum dx^2, sum dy^2, sum dz^2, hops]
// (Green-Kubo; the serial sum of monte_carlo.cpp:396-400) or of the integer population / current bins (contacts;
// monte_carlo.h:566-573, 626-636), issued on every device's own stream behind its kernels.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): libcntmc.so has no link-time dependency on it, a process that
// already carries an NCCL (PyTorch bundles one) shares that copy, and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cntmc.h"

namespace {

// the few NCCL entry points used, with the library's own (stable, C) signatures
typedef struct ncclComm* ncclComm_t;
enum { ncclSuccessV = 0 };
enum { ncclInt64V = 4, ncclFloat64V = 8 };  // ncclDataType_t: ncclInt64 = 4, ncclDouble = 8
enum { ncclSumV = 0 };                      // ncclRedOp_t
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
Nccl& nccl() {
  static Nccl n;
  if (n.lib) return n;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) throw std::runtime_error(std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"));
  auto sym = [&](const char* name) {
    void* p = dlsym(lib, name);
    if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
    return p;
  };
  n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
  n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
  n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
  n.lib = lib;
  return n;
}

thread_local std::string g_multi_create_error;

}  // namespace

struct cntmc_multi {
  mutable std::string       err;
  std::vector<int>          dev;
  std::vector<cntmc_t*>     h;
  std::vector<cudaStream_t> stream;
  std::vector<ncclComm_t>   comm;
  std::vector<void*>        d_red;  // per device: the rows the all-reduce works on
  size_t                    red_bytes = 0;
  std::vector<int64_t>      first, count;  // shard of the population per device
  int64_t                   total = 0, hops = 0;
  int                       n_seg = 0;
  bool                      contacts = false;

  ~cntmc_multi() {
    for (size_t i = 0; i < h.size(); ++i) {
      cudaSetDevice(dev[i]);
      if (i < d_red.size() && d_red[i]) cudaFree(d_red[i]);
      if (i < comm.size() && comm[i]) nccl().CommDestroy(comm[i]);
      if (h[i]) cntmc_destroy(h[i]);
      if (i < stream.size() && stream[i]) cudaStreamDestroy(stream[i]);
    }
  }
  void cuda(cudaError_t e, const char* what) const {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
  void nc(int r, const char* what) const {
    if (r != ncclSuccessV) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(r));
  }
  void each(int rc, size_t i) const {
    if (rc != CNTMC_OK) throw std::runtime_error("GPU " + std::to_string(dev[i]) + ": " + cntmc_last_error(h[i]));
  }
  void reserve_rows(size_t bytes) {
    if (bytes <= red_bytes) return;
    for (size_t i = 0; i < h.size(); ++i) {
      cuda(cudaSetDevice(dev[i]), "cudaSetDevice");
      if (d_red[i]) cudaFree(d_red[i]);
      d_red[i] = nullptr;
      cuda(cudaMalloc(&d_red[i], bytes), "cudaMalloc");
    }
    red_bytes = bytes;
  }
  // sum the first `n` elements of every device's rows over all devices, in place, behind the kernels already queued
  void all_reduce(size_t n, int dtype) {
    nc(nccl().GroupStart(), "ncclGroupStart");
    for (size_t i = 0; i < h.size(); ++i) nc(nccl().AllReduce(d_red[i], d_red[i], n, dtype, ncclSumV, comm[i], stream[i]), "ncclAllReduce");
    nc(nccl().GroupEnd(), "ncclGroupEnd");
  }
};

namespace {
template <typename F>
int guarded(const cntmc_multi_t* m, F&& body) {
  try {
    body();
    return CNTMC_OK;
  } catch (const std::invalid_argument& e) {
    (m ? m->err : g_multi_create_error) = e.what();
    return CNTMC_ERR_INVALID;
  } catch (const std::exception& e) {
    (m ? m->err : g_multi_create_error) = e.what();
    return CNTMC_ERR_CUDA;
  }
}
void shard(int64_t total, int r, int w, int64_t& first, int64_t& count) {
  const int64_t base = total / w, extra = total % w;
  count = base + (r < extra ? 1 : 0);
  first = r * base + (r < extra ? r : extra);
}
}  // namespace

extern "C" {

const char* cntmc_multi_last_error(const cntmc_multi_t* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int cntmc_multi_create(const char* json_text, int n_devices, const int* devices, cntmc_multi_t** out) {
  if (out) *out = nullptr;
  return guarded(nullptr, [&] {
    if (!json_text || !out || n_devices < 1) throw std::invalid_argument("bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device: this engine has no CPU execution path");
    std::unique_ptr<cntmc_multi_t> m(new cntmc_multi_t);
    for (int i = 0; i < n_devices; ++i) {
      const int d = devices ? devices[i] : i;
      if (d < 0 || d >= ndev) throw std::invalid_argument("device index out of range");
      for (int k : m->dev)
        if (k == d) throw std::invalid_argument("a device is listed twice");
      m->dev.push_back(d);
    }
    m->h.assign((size_t)n_devices, nullptr);
    m->stream.assign((size_t)n_devices, nullptr);
    m->comm.assign((size_t)n_devices, nullptr);
    m->d_red.assign((size_t)n_devices, nullptr);
    m->first.assign((size_t)n_devices, 0);
    m->count.assign((size_t)n_devices, 0);
    for (size_t i = 0; i < m->h.size(); ++i) {
      if (cntmc_create(json_text, &m->h[i]) != CNTMC_OK) throw std::invalid_argument(cntmc_last_error(nullptr));
      m->cuda(cudaSetDevice(m->dev[i]), "cudaSetDevice");
      m->cuda(cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking), "cudaStreamCreate");
      m->each(cntmc_set_device(m->h[i], m->dev[i]), i);
      m->each(cntmc_set_stream(m->h[i], m->stream[i]), i);
    }
    m->nc(nccl().CommInitAll(m->comm.data(), n_devices, m->dev.data()), "ncclCommInitAll");
    *out = m.release();
  });
}

void cntmc_multi_destroy(cntmc_multi_t* m) { delete m; }

int      cntmc_multi_num_devices(const cntmc_multi_t* m) { return (int)m->h.size(); }
cntmc_t* cntmc_multi_handle(cntmc_multi_t* m, int i) { return (i >= 0 && i < (int)m->h.size()) ? m->h[(size_t)i] : nullptr; }
int      cntmc_multi_nccl_version(void) {
  try {
    int v = 0;
    nccl().GetVersion(&v);
    return v;
  } catch (...) {
    return -1;
  }
}

int cntmc_multi_load_mesh(cntmc_multi_t* m, const char* dir) {
  return guarded(m, [&] {
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_load_mesh(m->h[i], dir), i);
  });
}
int cntmc_multi_set_mesh(cntmc_multi_t* m, int64_t n_tubes, int64_t n_cols, const double* pos_nm, const double* orient) {
  return guarded(m, [&] {
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_set_mesh(m->h[i], n_tubes, n_cols, pos_nm, orient), i);
  });
}
int cntmc_multi_set_option(cntmc_multi_t* m, const char* name, int64_t value) {
  return guarded(m, [&] {
    for (size_t i = 0; i < m->h.size(); ++i)
      if (cntmc_set_option(m->h[i], name, value) != CNTMC_OK) throw std::invalid_argument(cntmc_last_error(m->h[i]));
  });
}

int cntmc_multi_kubo_init(cntmc_multi_t* m) {
  return guarded(m, [&] {
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_kubo_init(m->h[i]), i);
    m->contacts = false;
  });
}

int cntmc_multi_kubo_create_particles(cntmc_multi_t* m, int64_t n_particles, uint64_t seed) {
  return guarded(m, [&] {
    // n_particles <= 0: "number of particles for kubo simulation" of the JSON -- the handles know it
    int64_t total = n_particles;
    if (total <= 0) {
      m->each(cntmc_kubo_create_particles(m->h[0], 0, seed, 0), 0);  // resolves the JSON value
      total = cntmc_number_of_particles(m->h[0]);
    }
    const int w = (int)m->h.size();
    if (total < w) throw std::invalid_argument("fewer excitons than GPUs");
    for (int r = 0; r < w; ++r) {
      shard(total, r, w, m->first[(size_t)r], m->count[(size_t)r]);
      m->each(cntmc_kubo_create_particles(m->h[(size_t)r], m->count[(size_t)r], seed, (uint64_t)m->first[(size_t)r]), (size_t)r);
    }
    m->total = total;
    m->hops = 0;
  });
}

int cntmc_multi_kubo_step(cntmc_multi_t* m, double dt, int64_t nsteps, double* msd_out) {
  return guarded(m, [&] {
    if (nsteps <= 0) throw std::invalid_argument("nsteps must be positive");
    if (m->total <= 0 || m->contacts) throw std::invalid_argument("call cntmc_multi_kubo_create_particles first");
    const size_t n = (size_t)nsteps * 4;
    m->reserve_rows(n * sizeof(double));
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_kubo_step_dev(m->h[i], dt, nsteps, static_cast<double*>(m->d_red[i])), i);
    m->all_reduce(n, ncclFloat64V);
    std::vector<double> rows(n);
    m->cuda(cudaSetDevice(m->dev[0]), "cudaSetDevice");
    m->cuda(cudaMemcpyAsync(rows.data(), m->d_red[0], n * sizeof(double), cudaMemcpyDeviceToHost, m->stream[0]), "cudaMemcpyAsync");
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_sync(m->h[i]), i);  // waits, and raises what the kernels flagged
    for (int64_t s = 0; s < nsteps; ++s) {
      m->hops += (int64_t)rows[(size_t)s * 4 + 3];
      if (msd_out)
        for (int c = 0; c < 3; ++c) msd_out[s * 3 + c] = rows[(size_t)s * 4 + c] / double(m->total);  // monte_carlo.cpp:402-404
    }
  });
}

int cntmc_multi_get_particles(const cntmc_multi_t* m, int32_t* site, double* pos, double* delta, double* ff, uint8_t* heading,
                              uint32_t* ndraw) {
  return guarded(m, [&] {
    if (m->total <= 0) throw std::invalid_argument("no excitons");
    const size_t P = (size_t)m->total;
    for (size_t i = 0; i < m->h.size(); ++i) {
      const size_t        n = (size_t)cntmc_number_of_particles(m->h[i]), at = m->contacts ? 0 : (size_t)m->first[i];
      std::vector<double> p(pos ? 3 * n : 0), d(delta ? 3 * n : 0);
      m->each(cntmc_get_particles(m->h[i], site ? site + at : nullptr, pos ? p.data() : nullptr, delta ? d.data() : nullptr,
                                  ff ? ff + at : nullptr, heading ? heading + at : nullptr, ndraw ? ndraw + at : nullptr), i);
      for (int c = 0; c < 3; ++c) {
        if (pos) std::memcpy(pos + c * P + at, p.data() + c * n, n * sizeof(double));
        if (delta) std::memcpy(delta + c * P + at, d.data() + c * n, n * sizeof(double));
      }
    }
  });
}

int64_t cntmc_multi_number_of_particles(const cntmc_multi_t* m) {
  if (!m->contacts) return m->total;
  int64_t n = 0;
  for (cntmc_t* h : m->h) n += cntmc_number_of_particles(h);
  return n;
}
int64_t cntmc_multi_hops(const cntmc_multi_t* m) {
  if (!m->contacts) return m->hops;
  int64_t n = 0;
  for (cntmc_t* h : m->h) n += cntmc_hops(h);
  return n;
}
double cntmc_multi_time(const cntmc_multi_t* m) { return cntmc_time(m->h[0]); }

// ---- contact flavour: the contact populations are split over the GPUs, every GPU runs its own excitons on a full copy of the film
int cntmc_multi_init(cntmc_multi_t* m, int64_t c1_pop, int64_t c2_pop, uint64_t seed) {
  return guarded(m, [&] {
    const int w = (int)m->h.size();
    for (int r = 0; r < w; ++r) {
      int64_t f1, n1, f2, n2;
      shard(c1_pop, r, w, f1, n1);
      shard(c2_pop, r, w, f2, n2);
      // stream ids of different GPUs never meet: GPU r numbers its excitons from r * 2^56
      m->each(cntmc_set_option(m->h[(size_t)r], "gid_base_shift56", r), (size_t)r);
      m->each(cntmc_init(m->h[(size_t)r], n1, n2, seed, 0), (size_t)r);
    }
    m->n_seg = cntmc_number_of_segments(m->h[0]);
    m->contacts = true;
    m->total = 1;
  });
}

int cntmc_multi_step(cntmc_multi_t* m, double dt, int64_t nsteps, int64_t* pop_out, int64_t* curr_out) {
  return guarded(m, [&] {
    if (!m->contacts) throw std::invalid_argument("call cntmc_multi_init first");
    if (nsteps <= 0) throw std::invalid_argument("nsteps must be positive");
    const int    nb = 2 * m->n_seg - 1;
    const size_t n = (size_t)nsteps * (size_t)nb;
    m->reserve_rows(n * sizeof(int64_t));
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_step_dev(m->h[i], dt, nsteps, static_cast<int64_t*>(m->d_red[i])), i);
    m->all_reduce(n, ncclInt64V);
    std::vector<int64_t> bins(n);
    m->cuda(cudaSetDevice(m->dev[0]), "cudaSetDevice");
    m->cuda(cudaMemcpyAsync(bins.data(), m->d_red[0], n * sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream[0]), "cudaMemcpyAsync");
    for (size_t i = 0; i < m->h.size(); ++i) m->each(cntmc_sync(m->h[i]), i);
    for (int64_t s = 0; s < nsteps; ++s) {
      for (int i = 0; i < m->n_seg; ++i)
        if (pop_out) pop_out[s * m->n_seg + i] = bins[(size_t)s * nb + i];
      for (int i = 0; i + 1 < m->n_seg; ++i)
        if (curr_out) curr_out[s * (m->n_seg - 1) + i] = bins[(size_t)s * nb + m->n_seg + i];
    }
  });
}

}  // extern "C"
