// fast_log.h -- natural logarithm for arguments in (0, 1]: the free-flight draw's log(r / RAND_MAX) (scatterer.h:79).
//
// The CUDA library's log is ~100 instructions and was 14 % of everything the hop kernel executed.  This one is ~35:
// table-driven (128 intervals of [0.6875, 1.375), reciprocal of the interval centre and the logarithm of the centre it
// stands for as a double-double), a 7th-order Taylor tail, every product that matters through an explicit fused
// multiply-add.  Arguments within 1/16 of 1 take a series of their own (the table would lose relative accuracy there).
// It is the same IEEE arithmetic on the host and on the device, so the exhaustive comparison with glibc over all
// 2^31 - 1 possible draws runs on the CPU (tests/test_host_core.py; tools/log_exhaustive.c): it never differs from
// glibc's log by more than one unit in the last place, and is itself within 0.52 ulp of the true value on every
// sampled argument.  Replay of the reference's draws does not use it (the host's own logarithms are replayed).
#pragma once
#include <stdint.h>
#include <string.h>

// included by hop_core.h after CNTMC_HD, fma_rn() and load32() are defined

namespace cntmc {

struct alignas(32) LogRow {
  double invc, logc_hi, logc_lo, pad;
};
#define CNTMC_LOG_ROW(a, b, c) {a, b, c, 0.0},
#if defined(__CUDACC__)
static __device__ const LogRow d_log_tab[128] = {
#include "log_table.inc"
};
#endif
static const LogRow h_log_tab[128] = {
#include "log_table.inc"
};
#undef CNTMC_LOG_ROW

CNTMC_HD double fast_log_unit(double x) {
#if defined(CNTMC_LIBM_LOG)  // A/B builds only: the library's logarithm
  return log(x);
#endif
  uint64_t ix;
#if defined(__CUDA_ARCH__)
  ix = (uint64_t)__double_as_longlong(x);
#else
  memcpy(&ix, &x, 8);
#endif
  if (ix - 0x3fee000000000000ULL <= 0x3ff0000000000000ULL - 0x3fee000000000000ULL) {  // 0.9375 <= x <= 1
    const double r = x - 1.0;  // exact
    const double s = r * r, e = fma_rn(r, r, -s);  // r^2 = s + e exactly
    const double t = fma_rn(-0.5, s, r);
    const double err = fma_rn(-0.5, s, r - t);  // (r - s/2) - t, exact: |r| >= |s/2|
    double       q = -1.0 / 16.0;
    q = fma_rn(q, r, 1.0 / 15.0);
    q = fma_rn(q, r, -1.0 / 14.0);
    q = fma_rn(q, r, 1.0 / 13.0);
    q = fma_rn(q, r, -1.0 / 12.0);
    q = fma_rn(q, r, 1.0 / 11.0);
    q = fma_rn(q, r, -1.0 / 10.0);
    q = fma_rn(q, r, 1.0 / 9.0);
    q = fma_rn(q, r, -1.0 / 8.0);
    q = fma_rn(q, r, 1.0 / 7.0);
    q = fma_rn(q, r, -1.0 / 6.0);
    q = fma_rn(q, r, 1.0 / 5.0);
    q = fma_rn(q, r, -1.0 / 4.0);
    q = fma_rn(q, r, 1.0 / 3.0);
    const double lo = fma_rn(s * r, q, fma_rn(-0.5, e, err));
    return t + lo;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const int      i = (int)((tmp >> 45) & 127u);
  const int      k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  double         z;
#if defined(__CUDA_ARCH__)
  z = __longlong_as_double((long long)iz);
  const Quad   row = load32(d_log_tab + i);  // one 256-bit load
  const double invc = row.a, lchi = row.b, lclo = row.c;
#else
  memcpy(&z, &iz, 8);
  const double invc = h_log_tab[i].invc, lchi = h_log_tab[i].logc_hi, lclo = h_log_tab[i].logc_lo;
#endif
  const double r = fma_rn(z, invc, -1.0);
  const double kd = (double)k;
  const double w = fma_rn(kd, CNTMC_LN2_HI, lchi);  // exact: both on the 2^-43 grid
  const double hi = w + r;
  const double lo = fma_rn(kd, CNTMC_LN2_LO, ((w - hi) + r) + lclo);  // (w - hi) + r is exact: |w| >= |r| here
  double       p = 1.0 / 7.0;
  p = fma_rn(p, r, -1.0 / 6.0);
  p = fma_rn(p, r, 1.0 / 5.0);
  p = fma_rn(p, r, -1.0 / 4.0);
  p = fma_rn(p, r, 1.0 / 3.0);
  p = fma_rn(p, r, -0.5);
  return hi + fma_rn(r * r, p, lo);
}

}  // namespace cntmc
