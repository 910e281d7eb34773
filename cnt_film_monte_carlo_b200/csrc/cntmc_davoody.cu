// cntmc_davoody.cu -- C ABI of the "davoody" rate-table builder (include/cntmc.h, section "davoody rate table").
//
// Replaces, for "rate type":"davoody", the reference's monte_carlo::create_scattering_table (monte_carlo.cpp:24-49) and
// monte_carlo::create_davoody_scatt_table (monte_carlo.cpp:64-153): tube physics on the host once per chirality
// (davoody_tube.h), placement-independent transfer factors on the host once per tube pair (davoody_transfer.h), and one
// thread block per table entry on the GPU (davoody_kernels.cuh).  No CPU path computes rates: without a CUDA device
// cntmc_transfer_create fails.
#include <cuda_runtime.h>

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cntmc.h"
#include "davoody_kernels.cuh"
#include "davoody_transfer.h"

using namespace cntmc::davoody;

struct cntmc_tube {
  std::unique_ptr<Tube> tube;
  double                build_seconds = 0;
};

namespace {
thread_local std::string g_error;

struct CudaFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define DV_CUDA(expr)                                                                                                \
  do {                                                                                                               \
    cudaError_t _e = (expr);                                                                                         \
    if (_e != cudaSuccess) throw CudaFailure(std::string(#expr) + " failed: " + cudaGetErrorString(_e));             \
  } while (0)

template <typename T>
struct Dev {
  T* p = nullptr;
  Dev() {}
  Dev(const Dev&) = delete;
  Dev& operator=(const Dev&) = delete;
  ~Dev() {
    if (p) cudaFree(p);
  }
  void put(const T* h, size_t n) {
    if (p) cudaFree(p);
    p = nullptr;
    DV_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) DV_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  void put(const std::vector<T>& v) { put(v.data(), v.size()); }
  void reserve(size_t n) {
    if (p) cudaFree(p);
    p = nullptr;
    DV_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  }
};

template <typename F>
int guarded(F&& body) {
  try {
    body();
    return CNTMC_OK;
  } catch (const CudaFailure& e) {
    g_error = e.what();
    return CNTMC_ERR_CUDA;
  } catch (const std::exception& e) {
    g_error = e.what();
    return CNTMC_ERR_INVALID;
  }
}
void require(bool ok, const char* what) {
  if (!ok) throw std::invalid_argument(what);
}
}  // namespace

struct cntmc_transfer {
  std::unique_ptr<Transfer> host;
  int                       device = 0;
  int                       chunk = 4, threads = 128, Kd_pad = 0;
  size_t                    smem = 0;
  Dev<double>               d_x, d_y, d_z, a_x, a_y, a_z, boltzmann, lorentz;
  Dev<double2>              d_phase, a_phase, Q;
  Dev<int2>                 pair_k;
  double                    last_kernel_ms = 0;
  long long                 launches = 0;
};

// Measured FP64 pipe peak of the device, the denominator of the table kernel's roofline: 8 independent fused multiply-add
// chains per thread, enough resident warps to fill every SM, no memory traffic.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0 - 1e-9, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double v = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (v == 12345.678) out[0] = v;  // keeps the chains alive
}

#include "herm_eig.h"
extern "C" {

int cntmc_hermitian_eig(int n, const double* a_re_im, double* w, double* v_re_im) {
  return guarded([&] {
    require(n > 0 && a_re_im && w && v_re_im, "bad argument");
    std::vector<std::complex<double>> A((size_t)n * n), V((size_t)n * n);
    for (size_t k = 0; k < A.size(); k++) A[k] = std::complex<double>(a_re_im[2 * k], a_re_im[2 * k + 1]);
    const int sweeps = cntmc::hermitian_eig(n, A.data(), w, V.data());
    require(sweeps >= 0, "the Jacobi sweeps did not converge");
    for (size_t k = 0; k < V.size(); k++) {
      v_re_im[2 * k] = V[k].real();
      v_re_im[2 * k + 1] = V[k].imag();
    }
  });
}

int cntmc_fp64_peak(int device, double* tflops) {
  return guarded([&] {
    require(tflops != nullptr, "null argument");
    DV_CUDA(cudaSetDevice(device));
    int sms = 0;
    DV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    Dev<double> out;
    out.reserve(1);
    const int   iters = 1 << 15, blocks = sms * 8;
    cudaEvent_t e0, e1;
    DV_CUDA(cudaEventCreate(&e0));
    DV_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {  // the first pass warms up
      DV_CUDA(cudaEventRecord(e0, 0));
      fp64_peak_kernel<<<blocks, 256>>>(out.p, iters, 1.0 + rep);
      DV_CUDA(cudaEventRecord(e1, 0));
      DV_CUDA(cudaGetLastError());
      DV_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      DV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double t = 2.0 * 8 * (double)iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
      if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
  });
}

const char* cntmc_davoody_last_error(void) { return g_error.c_str(); }

cntmc_tube_t* cntmc_tube_create(int n, int m, int length_cells) {
  cntmc_tube_t* out = nullptr;
  guarded([&] {
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<cntmc_tube> h(new cntmc_tube);
    h->tube.reset(new Tube(n, m, length_cells));
    h->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out = h.release();
  });
  return out;
}

void cntmc_tube_destroy(cntmc_tube_t* t) { delete t; }

int cntmc_tube_info(const cntmc_tube_t* t, int32_t ints[8], double reals[4]) {
  return guarded([&] {
    require(t && ints && reals, "null argument");
    const Tube& T = *t->tube;
    const int32_t i[8] = {T.n, T.m, T.cells, T.Nu, T.M, T.Q, T.nk, T.n_sites()};
    const double  r[4] = {T.radius, T.length_in_meter(), T.cell_area(), t->build_seconds};
    std::memcpy(ints, i, sizeof(i));
    std::memcpy(reals, r, sizeof(r));
  });
}

int cntmc_tube_exciton_dims(const cntmc_tube_t* t, int which, int32_t dims[4]) {
  return guarded([&] {
    require(t && dims, "null argument");
    require(which >= 0 && which < 3, "which must be 0 (A1), 1 (A2 singlet) or 2 (A2 triplet)");
    const Exciton& ex = t->tube->excitons[which];
    dims[0] = ex.nk_cm;
    dims[1] = ex.n_principal;
    dims[2] = ex.nk_c;
    dims[3] = ex.ik_cm_begin;
  });
}

int cntmc_tube_exciton_energy(const cntmc_tube_t* t, int which, double* energy) {
  return guarded([&] {
    require(t && energy, "null argument");
    require(which >= 0 && which < 3, "which must be 0 (A1), 1 (A2 singlet) or 2 (A2 triplet)");
    const Exciton& ex = t->tube->excitons[which];
    std::copy(ex.energy.begin(), ex.energy.end(), energy);
  });
}

cntmc_transfer_t* cntmc_transfer_create(const cntmc_tube_t* donor, const cntmc_tube_t* acceptor, double temperature_kelvin,
                                        double broadening_joule, int device) {
  cntmc_transfer_t* out = nullptr;
  guarded([&] {
    require(donor && acceptor, "null tube");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw CudaFailure("no CUDA device: the davoody table has no CPU path");
    std::unique_ptr<cntmc_transfer> h(new cntmc_transfer);
    if (device >= 0) {
      DV_CUDA(cudaSetDevice(device));
      h->device = device;
    } else {
      DV_CUDA(cudaGetDevice(&h->device));
    }
    h->host.reset(new Transfer(*donor->tube, *acceptor->tube, temperature_kelvin, broadening_joule));
    const Transfer& X = *h->host;
    const int       Nd = X.donor.n_sites(), Na = X.acceptor.n_sites(), Kd = int(X.d_kcm.size()), Ka = int(X.a_kcm.size());
    h->chunk = Kd <= 4 ? 4 : (Kd <= 8 ? 8 : 16);
    h->Kd_pad = (Kd + h->chunk - 1) / h->chunk * h->chunk;
    // threads = acceptor sites per pass: the number of passes (of at most 320 threads) that leaves the fewest idle threads
    // (280 sites: one pass of 288, not 256 + 24 and not 2 x 160; every pass re-stages the donor tiles)
    int  best_threads = 0;
    long best_slots = -1;
    for (int passes = (Na + 319) / 320; passes <= (Na + 319) / 320 + 3; passes++) {
      const int  t = std::max(32, ((Na + passes - 1) / passes + 31) / 32 * 32);
      const long slots = (long)t * passes;
      if (t <= 320 && (best_slots < 0 || slots < best_slots)) {
        best_slots = slots;
        best_threads = t;
      }
    }
    h->threads = best_threads;
    h->smem = placement_smem_bytes(h->chunk, h->threads, h->Kd_pad, Ka);
    require(h->smem <= 200 * 1024, "too many distinct K_cm among the matched states for one thread block's shared memory");
    h->d_x.put(X.d_sites.x);
    h->d_y.put(X.d_sites.y_centred);
    h->d_z.put(X.d_sites.z);
    h->a_x.put(X.a_sites.x);
    h->a_y.put(X.a_sites.y_centred);
    h->a_z.put(X.a_sites.z);
    std::vector<double2> dp((size_t)h->Kd_pad * Nd, make_double2(0.0, 0.0)), ap((size_t)Na * Ka);
    for (size_t i = 0; i < X.d_phase.size(); i++) dp[i] = make_double2(X.d_phase[i].real(), X.d_phase[i].imag());
    for (size_t i = 0; i < X.a_phase.size(); i++) ap[i] = make_double2(X.a_phase[i].real(), X.a_phase[i].imag());
    h->d_phase.put(dp);
    h->a_phase.put(ap);
    std::vector<int2>    pk(X.pairs.size());
    std::vector<double2> q(X.pairs.size());
    std::vector<double>  bz(X.pairs.size()), lz(X.pairs.size());
    for (size_t p = 0; p < X.pairs.size(); p++) {
      pk[p] = make_int2(X.pairs[p].kd, X.pairs[p].ka);
      q[p] = make_double2(X.pairs[p].Q.real(), X.pairs[p].Q.imag());
      bz[p] = X.pairs[p].boltzmann_rate;
      lz[p] = X.pairs[p].lorentzian;
    }
    h->pair_k.put(pk);
    h->Q.put(q);
    h->boltzmann.put(bz);
    h->lorentz.put(lz);
    if (h->smem > 48 * 1024) {
      DV_CUDA(cudaFuncSetAttribute(placement_rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
      DV_CUDA(cudaFuncSetAttribute(placement_rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
      DV_CUDA(cudaFuncSetAttribute(placement_rate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    }
    out = h.release();
  });
  return out;
}

void cntmc_transfer_destroy(cntmc_transfer_t* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  delete x;
}

int cntmc_transfer_info(const cntmc_transfer_t* x, int32_t ints[8], double reals[4]) {
  return guarded([&] {
    require(x && ints && reals, "null argument");
    const Transfer& X = *x->host;
    const int32_t   i[8] = {(int32_t)X.d_states.size(), (int32_t)X.a_states.size(), (int32_t)X.pairs.size(), (int32_t)X.d_kcm.size(),
                          (int32_t)X.a_kcm.size(),    x->chunk,                   x->threads,              (int32_t)x->smem};
    const double    r[4] = {X.temperature, X.broadening, x->last_kernel_ms, double(x->launches)};
    std::memcpy(ints, i, sizeof(i));
    std::memcpy(reals, r, sizeof(r));
  });
}

int cntmc_transfer_pair_factors(const cntmc_transfer_t* x, double* q_re_im, double* boltzmann, double* lorentzian) {
  return guarded([&] {
    require(x, "null argument");
    const Transfer& X = *x->host;
    for (size_t p = 0; p < X.pairs.size(); p++) {
      if (q_re_im) {
        q_re_im[2 * p] = X.pairs[p].Q.real();
        q_re_im[2 * p + 1] = X.pairs[p].Q.imag();
      }
      if (boltzmann) boltzmann[p] = X.pairs[p].boltzmann_rate;
      if (lorentzian) lorentzian[p] = X.pairs[p].lorentzian;
    }
  });
}

int cntmc_transfer_first_order(cntmc_transfer_t* x, int64_t n, const double* z_shift, const double* axis_shift_1,
                               const double* axis_shift_2, const double* theta, double* rate) {
  return guarded([&] {
    require(x && z_shift && axis_shift_1 && axis_shift_2 && theta && rate, "null argument");
    require(n >= 0, "negative count");
    if (n == 0) return;
    DV_CUDA(cudaSetDevice(x->device));
    const Transfer& X = *x->host;
    if (X.pairs.empty()) {  // no energy-matched states: the reference's sum over pairs is empty
      std::fill(rate, rate + n, 0.0);
      return;
    }
    std::vector<double> c(n), s(n);
    for (int64_t g = 0; g < n; g++) {
      c[g] = std::cos(theta[g]);
      s[g] = std::sin(theta[g]);
    }
    Dev<double> dz, d1, d2, dc, ds, dr;
    dz.put(z_shift, n);
    d1.put(axis_shift_1, n);
    d2.put(axis_shift_2, n);
    dc.put(c);
    ds.put(s);
    dr.reserve(n);
    PlacementArgs A;
    A.Nd = X.donor.n_sites();
    A.Na = X.acceptor.n_sites();
    A.Kd_pad = x->Kd_pad;
    A.Ka = int(X.a_kcm.size());
    A.n_pairs = int(X.pairs.size());
    A.d_x = x->d_x.p;
    A.d_y = x->d_y.p;
    A.d_z = x->d_z.p;
    A.a_x = x->a_x.p;
    A.a_y = x->a_y.p;
    A.a_z = x->a_z.p;
    A.d_phase = x->d_phase.p;
    A.a_phase = x->a_phase.p;
    A.pair_k = x->pair_k.p;
    A.pair_Q = x->Q.p;
    A.pair_boltzmann = x->boltzmann.p;
    A.pair_lorentzian = x->lorentz.p;
    A.sqrt_lengths = X.sqrt_lengths;
    A.n_placements = n;
    A.z_shift = dz.p;
    A.shift_d = d1.p;
    A.shift_a = d2.p;
    A.cos_t = dc.p;
    A.sin_t = ds.p;
    A.rate = dr.p;
    const unsigned grid = (unsigned)std::min<int64_t>(n, 1 << 20);
    cudaEvent_t    e0, e1;
    DV_CUDA(cudaEventCreate(&e0));
    DV_CUDA(cudaEventCreate(&e1));
    DV_CUDA(cudaEventRecord(e0, 0));
    switch (x->chunk) {
      case 4: placement_rate_kernel<4><<<grid, x->threads, x->smem>>>(A); break;
      case 8: placement_rate_kernel<8><<<grid, x->threads, x->smem>>>(A); break;
      default: placement_rate_kernel<16><<<grid, x->threads, x->smem>>>(A); break;
    }
    DV_CUDA(cudaEventRecord(e1, 0));
    DV_CUDA(cudaGetLastError());
    DV_CUDA(cudaMemcpy(rate, dr.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    float ms = 0;
    DV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    x->last_kernel_ms = ms;
    x->launches++;
  });
}

int cntmc_transfer_table(cntmc_transfer_t* x, const int32_t dims[4], const double* theta, const double* z_shift,
                         const double* axis_shift_1, const double* axis_shift_2, double* rates) {
  int rc = guarded([&] {
    require(x && dims && theta && z_shift && axis_shift_1 && axis_shift_2 && rates, "null argument");
    require(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[3] > 0, "table dimensions must be positive");
  });
  if (rc != CNTMC_OK) return rc;
  // the loop nest of monte_carlo.cpp:114-137, flattened: one placement per table entry, theta slowest
  const int64_t       n = (int64_t)dims[0] * dims[1] * dims[2] * dims[3];
  std::vector<double> th(n), z(n), s1(n), s2(n);
  int64_t             g = 0;
  for (int i = 0; i < dims[0]; i++)
    for (int k = 0; k < dims[1]; k++)
      for (int p = 0; p < dims[2]; p++)
        for (int q = 0; q < dims[3]; q++, g++) {
          th[g] = theta[i];
          z[g] = z_shift[k];
          s1[g] = axis_shift_1[p];
          s2[g] = axis_shift_2[q];
        }
  return cntmc_transfer_first_order(x, n, z.data(), s1.data(), s2.data(), th.data(), rates);
}

int cntmc_create_davoody_table(cntmc_t* h, cntmc_transfer_t* x, const int32_t dims[4], const double* theta, const double* z_shift,
                               const double* axis_shift_1, const double* axis_shift_2) {
  if (!h || !x || !dims) {
    g_error = "null argument";
    return CNTMC_ERR_INVALID;
  }
  std::vector<double> rates;
  int                 rc = guarded([&] {
    require(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[3] > 0, "table dimensions must be positive");
    rates.resize((size_t)dims[0] * dims[1] * dims[2] * dims[3]);
  });
  if (rc != CNTMC_OK) return rc;
  rc = cntmc_transfer_table(x, dims, theta, z_shift, axis_shift_1, axis_shift_2, rates.data());
  if (rc != CNTMC_OK) return rc;
  rc = cntmc_set_rate_table(h, dims, theta, z_shift, axis_shift_1, axis_shift_2, rates.data());
  if (rc != CNTMC_OK) g_error = cntmc_last_error(h);
  return rc;
}

}  // extern "C"
