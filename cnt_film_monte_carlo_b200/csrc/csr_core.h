// csr_core.h -- pair geometry and rate lookup for the neighbour-table builder (host/device shared, see hop_core.h).
//
// The reference has no stored neighbour table: scatterer::find_neighbors (scatterer.cpp:34-83) rebuilds one site's
// list on every hop.  The list is a pure function of the site, so the engine evaluates it once per site and keeps
// the result as a CSR row.  The functions below are that evaluation for one candidate pair.
#pragma once
#include "hop_core.h"

namespace cntmc {

struct RateTable {
  const double* theta;  // grids (monte_carlo.cpp:157-167)
  const double* z;
  const double* a1;
  const double* a2;
  const double* rates;  // [theta][z][a1][a2], C order (scattering_struct.h:82-90)
  int32_t       n_theta, n_z, n_a1, n_a2;
  // evenly spaced ascending grids (what linspace produces): where the nearest value must be, see argmin_abs_near.
  // inv_step = 0 marks an axis that is searched in full (a loaded table with an irregular grid).
  double        start[4], inv_step[4];
  double        guard_tol;  // a theta within this many grid pitches of a midpoint flags its row (see pair_rate)
};
// axis k of a table: start and 1/step if the grid is ascending and evenly spaced to 1e-9 of its pitch, else {*, 0}
inline void grid_hint(const double* g, int n, double* start, double* inv_step) {
  *start = n > 0 ? g[0] : 0.0;
  *inv_step = 0.0;
  if (n < 2) return;
  const double step = (g[n - 1] - g[0]) / double(n - 1);
  if (!(step > 0)) return;
  for (int i = 0; i < n; ++i)
    if (!(fabs(g[i] - (g[0] + step * i)) <= 1e-9 * step)) return;
  *inv_step = 1.0 / step;
}

// arma::abs(grid - x).index_min() (scattering_struct.h:42-49): first strict minimum from +inf, NaN never wins
CNTMC_HD int argmin_abs(const double* grid, int n, double x) {
  double best = INFINITY;
  int    idx = 0;
  for (int i = 0; i < n; ++i) {
    const double d = fabs(ro(grid + i) - x);
    if (d < best) {
      best = d;
      idx = i;
    }
  }
  return idx;
}

// The same index without scanning the whole grid.  |grid[i] - x| over an ascending grid falls, then rises, so the first
// strict minimum lies within one place of round((x - start) / step) -- two places are scanned on either side, in ascending
// order with the same strict comparison, starting from the same +inf / index 0, and a run of equal distances is followed
// down to its first place, so ties, values far beyond the ends, NaN and infinities resolve exactly as in the full scan
// (brute-force comparison: tests/test_host_core.py; whole tables against the oracle: test_oracle_*, test_gpu_*).
CNTMC_HD int argmin_abs_near(const double* grid, int n, double x, double start, double inv_step) {
  if (!(inv_step > 0)) return argmin_abs(grid, n, x);
  const double f = (x - start) * inv_step;
  int          j = (f > -4.0) ? ((f < (double)n + 4.0) ? (int)(f + 0.5) : n - 1) : 0;  // NaN, -inf -> 0
  j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
  const int lo = j - 2 < 0 ? 0 : j - 2, hi = j + 2 > n - 1 ? n - 1 : j + 2;
  double    best = INFINITY;
  int       idx = 0;
  for (int i = lo; i <= hi; ++i) {
    const double d = fabs(ro(grid + i) - x);
    if (d < best) {
      best = d;
      idx = i;
    }
  }
  // far beyond the upper end the rounded distances tie over many places (axis shifts of nearly parallel tubes reach 1e7
  // grid pitches): the scan keeps the FIRST of them
  while (idx > 0 && fabs(ro(grid + idx - 1) - x) == best) --idx;
  return idx;
}

// true if x sits within `tol` (relative to the grid pitch) of the midpoint between its two nearest grid values, i.e.
// an acos that differs from glibc's in the last place could have picked the other one
CNTMC_HD bool near_midpoint(const double* grid, int n, int idx, double x, double tol) {
  bool near = false;
  if (idx > 0) {
    const double lo = ro(grid + idx - 1), hi = ro(grid + idx);
    near |= fabs(x - 0.5 * (lo + hi)) <= tol * fabs(hi - lo);
  }
  if (idx + 1 < n) {
    const double lo = ro(grid + idx), hi = ro(grid + idx + 1);
    near |= fabs(x - 0.5 * (lo + hi)) <= tol * fabs(hi - lo);
  }
  return near;
}

struct SiteGeom {
  double px, py, pz, ox, oy, oz;
};

// scatterer.cpp:40-43
CNTMC_HD bool within_cutoff(const SiteGeom& s1, const SiteGeom& s2, double radius) {
  const double distance = norm3(s1.px - s2.px, s1.py - s2.py, s1.pz - s2.pz);
  return (distance < radius) && (distance > kMinDist);
}

// scatterer.cpp:44-63 for an accepted pair; s1 is the departure site.  *guard is set when theta is within R.guard_tol
// (1e-9) pitches of a grid midpoint -- the only place where the device's acos, which may differ from glibc's in the last
// place, could pick another index than the reference.  Rows flagged this way are recomputed on the host with glibc's acos
// and patched in (cntmc_api.cu, repair_flagged_rows).
CNTMC_HD double pair_rate(const SiteGeom& s1, const SiteGeom& s2, const RateTable& R, bool* guard) {
  const double dRx = s1.px - s2.px, dRy = s1.py - s2.py, dRz = s1.pz - s2.pz;
  const double cosTheta = dot3(s1.ox, s1.oy, s1.oz, s2.ox, s2.oy, s2.oz);
  double       theta, axis_shift_1, axis_shift_2, z_shift;
  if (cosTheta == 1) {
    axis_shift_1 = 0;
    axis_shift_2 = dot3(dRx, dRy, dRz, s1.ox, s1.oy, s1.oz);
    theta = 0;
    const double t = dot3(dRx, dRy, dRz, s1.ox, s1.oy, s1.oz);
    z_shift = norm3(dRx - s1.ox * t, dRy - s1.oy * t, dRz - s1.oz * t);
  } else {
    theta = acos(cosTheta);
    const double y1 = dot3(s1.ox, s1.oy, s1.oz, dRx, dRy, dRz);
    const double y2 = dot3(s2.ox, s2.oy, s2.oz, dRx, dRy, dRz);
    const double sin2Theta = 1 - cosTheta * cosTheta;
    axis_shift_1 = (y1 + y2 * cosTheta) / sin2Theta;
    axis_shift_2 = (y2 + y1 * cosTheta) / sin2Theta;
    z_shift = norm3((s1.ox * axis_shift_1 + s1.px) - (s2.ox * axis_shift_2 + s2.px),
                    (s1.oy * axis_shift_1 + s1.py) - (s2.oy * axis_shift_2 + s2.py),
                    (s1.oz * axis_shift_1 + s1.pz) - (s2.oz * axis_shift_2 + s2.pz));
  }
  const int i_th = argmin_abs_near(R.theta, R.n_theta, theta, R.start[0], R.inv_step[0]);
  const int i_z = argmin_abs_near(R.z, R.n_z, z_shift, R.start[1], R.inv_step[1]);
  const int i_1 = argmin_abs_near(R.a1, R.n_a1, axis_shift_1, R.start[2], R.inv_step[2]);
  const int i_2 = argmin_abs_near(R.a2, R.n_a2, axis_shift_2, R.start[3], R.inv_step[3]);
  if (guard != nullptr && near_midpoint(R.theta, R.n_theta, i_th, theta, R.guard_tol)) *guard = true;
  return ro(R.rates + (((size_t)i_th * R.n_z + i_z) * R.n_a1 + i_1) * R.n_a2 + i_2);
}

// cell of a position in the bucket grid (monte_carlo.h:384-387): truncation of (x - xmin)/R
CNTMC_HD int cell_coord(double x, double lo, double radius) { return (int)((x - lo) / radius); }

}  // namespace cntmc
