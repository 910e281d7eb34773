// davoody_kernels.cuh -- the placement-dependent part of exciton_transfer::first_order (exciton_transfer.cpp:395-441) on the GPU:
// one thread block per placement (z shift, two axis shifts, angle) of the acceptor tube relative to the donor.
//
// For a placement the reference evaluates, for every energy-matched state pair p,
//     J_p = sum_{i in donor sites} sum_{j in acceptor sites} conj-phase_i(K_d(p)) phase_j(K_a(p)) / |R_i - R_j|      (calculate_J, :307-378)
// i.e. the N_d x N_a distance matrix once per pair.  Distances do not depend on the pair, and pairs share their K_cm:
//     W[kd][j]  = sum_i  phase_d[kd][i] / |R_i - R_j|            stage 1: N_d x N_a distances, once, 2 FMA per K_d each
//     J[kd][ka] = sum_j  W[kd][j] phase_a[j][ka]                 stage 2: tiny
//     rate      = sum_p  boltzmann_p |Q_p J[kd(p)][ka(p)]|^2 / (L_d L_a) lorentzian_p        stage 3, in the reference's pair order
// Stage 1 is double-precision arithmetic on data that lives in shared memory and registers: the kernel is bound by the FP64
// pipe (a square root, a division and 2 K_C fused multiply-adds per site pair), not by HBM.
//
// Arithmetic contract: site positions, distances and the per-pair products are formed with the reference's operations in
// the reference's order (the file is compiled with -fmad=false, so nothing is contracted), hence each 1/|R_i - R_j| equals
// the reference's bit for bit; the long sums are accumulated with explicit fma() in a different (blocked) order than the
// reference's single sequential loop, which is where the last-digits difference to the reference comes from.
#pragma once
#include <cuda_runtime.h>

namespace cntmc {
namespace davoody {

struct PlacementArgs {
  int Nd, Na;      // sites of donor / acceptor
  int Kd_pad, Ka;  // distinct K_cm of the donor (padded to the chunk size with zero phases) / the acceptor
  int n_pairs;
  const double *d_x, *d_y, *d_z;  // donor sites, axis centred (before the axis shift)
  const double *a_x, *a_y, *a_z;  // acceptor sites, axis centred (before shift, lift and rotation)
  const double2* d_phase;         // [Kd_pad][Nd]   exp(-i K_d . r_i)
  const double2* a_phase;         // [Na][Ka]       exp(+i K_a . r_j)
  const int2*    pair_k;          // (kd, ka) of every state pair
  const double2* pair_Q;
  const double*  pair_boltzmann;  // (2 pi / hbar) exp(-E_i / kT) / Z
  const double*  pair_lorentzian;
  double         sqrt_lengths;    // sqrt(L_d L_a)
  long long      n_placements;
  const double * z_shift, *shift_d, *shift_a, *cos_t, *sin_t;
  double*        rate;
};

constexpr int kDonorTile = 128;  // donor sites staged in shared memory at a time
constexpr int kTermTile = 1024;  // state pairs whose rate terms are staged before they are summed in order

inline size_t placement_smem_bytes(int KC, int threads, int Kd_pad, int Ka) {
  return sizeof(double2) * ((size_t)Kd_pad * Ka + (size_t)KC * threads + (size_t)KC * kDonorTile) + sizeof(double) * (3 * kDonorTile + kTermTile);
}

// KC = donor K_cm values carried in registers per pass over the distance matrix
template <int KC>
__global__ void __launch_bounds__(320) placement_rate_kernel(const PlacementArgs A) {
  extern __shared__ double2 smem2[];
  const int T = blockDim.x, tid = threadIdx.x;
  double2*  Js = smem2;                            // [Kd_pad][Ka]
  double2*  Ws = Js + (size_t)A.Kd_pad * A.Ka;     // [KC][T]
  double2*  ph = Ws + (size_t)KC * T;              // [KC][kDonorTile]
  double*   xs = reinterpret_cast<double*>(ph + (size_t)KC * kDonorTile);
  double*   ys = xs + kDonorTile;
  double*   zs = ys + kDonorTile;
  double*   terms = zs + kDonorTile;

  for (long long g = blockIdx.x; g < A.n_placements; g += gridDim.x) {
    const double lift = A.z_shift[g], shift_d = A.shift_d[g], shift_a = A.shift_a[g], c = A.cos_t[g], s = A.sin_t[g];
    for (int e = tid; e < A.Kd_pad * A.Ka; e += T) Js[e] = make_double2(0.0, 0.0);

    for (int kd0 = 0; kd0 < A.Kd_pad; kd0 += KC) {
      for (int j0 = 0; j0 < A.Na; j0 += T) {
        const int  j = j0 + tid;
        const bool valid = j < A.Na;
        // the acceptor site: shift along its axis, lift, then turn about z (make_Ru_3d, :309-337)
        double ax = 0, ay = 0, az = 0;
        if (valid) {
          const double x0 = A.a_x[j], y0 = A.a_y[j] + shift_a;
          az = A.a_z[j] + lift;
          ax = x0 * c - y0 * s;
          ay = x0 * s + y0 * c;
        }
        double2 W[KC];
#pragma unroll
        for (int k = 0; k < KC; k++) W[k] = make_double2(0.0, 0.0);

        for (int i0 = 0; i0 < A.Nd; i0 += kDonorTile) {
          const int ni = min(kDonorTile, A.Nd - i0);
          __syncthreads();
          for (int e = tid; e < ni; e += T) {
            xs[e] = A.d_x[i0 + e];
            ys[e] = A.d_y[i0 + e] + shift_d;
            zs[e] = A.d_z[i0 + e];
          }
          for (int e = tid; e < KC * ni; e += T) {
            const int k = e / ni, i = e - k * ni;
            ph[k * kDonorTile + i] = A.d_phase[(size_t)(kd0 + k) * A.Nd + i0 + i];
          }
          __syncthreads();
          if (valid) {
#pragma unroll 2
            for (int i = 0; i < ni; i++) {
              const double dx = xs[i] - ax, dy = ys[i] - ay, dz = zs[i] - az;
              const double inv = 1.0 / sqrt((dx * dx + dz * dz) + dy * dy);  // Armadillo's norm: even and odd elements summed apart
#pragma unroll
              for (int k = 0; k < KC; k++) {
                const double2 p = ph[k * kDonorTile + i];
                W[k].x = fma(p.x, inv, W[k].x);
                W[k].y = fma(p.y, inv, W[k].y);
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < KC; k++) Ws[k * T + tid] = W[k];
        __syncthreads();
        // stage 2: every (kd, ka) entry of J has one owner thread for the whole kernel, so the order of its sum is fixed
        const int nj = min(T, A.Na - j0);
        for (int e = tid; e < KC * A.Ka; e += T) {
          const int k = e / A.Ka, ka = e - k * A.Ka;
          double2   acc = make_double2(0.0, 0.0);
          for (int jj = 0; jj < nj; jj++) {
            const double2 w = Ws[k * T + jj], p = A.a_phase[(size_t)(j0 + jj) * A.Ka + ka];
            acc.x = fma(w.x, p.x, acc.x);
            acc.x = fma(-w.y, p.y, acc.x);
            acc.y = fma(w.x, p.y, acc.y);
            acc.y = fma(w.y, p.x, acc.y);
          }
          double2& J = Js[(size_t)(kd0 + k) * A.Ka + ka];
          J.x += acc.x;
          J.y += acc.y;
        }
        // Ws is rewritten only after the next pass's tile barriers
      }
    }
    __syncthreads();
    // stage 3: M = |Q J| / sqrt(L_d L_a); rate += boltzmann M^2 lorentzian, pair after pair (first_order, :417-431)
    double total = 0;
    for (int p0 = 0; p0 < A.n_pairs; p0 += kTermTile) {
      const int np = min(kTermTile, A.n_pairs - p0);
      for (int e = tid; e < np; e += T) {
        const int2    k = A.pair_k[p0 + e];
        const double2 Q = A.pair_Q[p0 + e], J = Js[(size_t)k.x * A.Ka + k.y];
        const double  re = Q.x * J.x - Q.y * J.y, im = Q.x * J.y + Q.y * J.x;
        const double  M = hypot(re, im) / A.sqrt_lengths;
        terms[e] = A.pair_boltzmann[p0 + e] * (M * M) * A.pair_lorentzian[p0 + e];
      }
      __syncthreads();
      if (tid == 0)
        for (int e = 0; e < np; e++) total += terms[e];
      __syncthreads();
    }
    if (tid == 0) A.rate[g] = total;
    __syncthreads();
  }
}

}  // namespace davoody
}  // namespace cntmc
