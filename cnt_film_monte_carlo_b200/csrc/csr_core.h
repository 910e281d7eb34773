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
};

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

// scatterer.cpp:44-63 for an accepted pair; s1 is the departure site.  *guard is set when theta is within 1e-9 pitch
// of a grid midpoint (the only place where the device's acos could change an index relative to glibc).
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
  const int i_th = argmin_abs(R.theta, R.n_theta, theta);
  const int i_z = argmin_abs(R.z, R.n_z, z_shift);
  const int i_1 = argmin_abs(R.a1, R.n_a1, axis_shift_1);
  const int i_2 = argmin_abs(R.a2, R.n_a2, axis_shift_2);
  if (guard != nullptr && near_midpoint(R.theta, R.n_theta, i_th, theta, 1e-9)) *guard = true;
  return ro(R.rates + (((size_t)i_th * R.n_z + i_z) * R.n_a1 + i_1) * R.n_a2 + i_2);
}

// cell of a position in the bucket grid (monte_carlo.h:384-387): truncation of (x - xmin)/R
CNTMC_HD int cell_coord(double x, double lo, double radius) { return (int)((x - lo) / radius); }

}  // namespace cntmc
