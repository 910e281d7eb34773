// herm_eig.h -- eigen-decomposition of a complex Hermitian matrix by cyclic Jacobi rotations.
//
// The reference obtains exciton states from arma::eig_sym (LAPACK zheev behind Armadillo; cnt.cpp:950-962).  Neither
// library exists in this image, and a tight-binding / Bethe-Salpeter kernel of a few dozen rows does not need one: Jacobi
// is unconditionally convergent for Hermitian matrices and as accurate as the input allows.  Eigenvalues come back in
// ascending order (like eig_sym), eigenvectors as the columns of V with V^H A V = diag(w); the phase of an eigenvector is
// arbitrary in any solver, and everything downstream (|Q J|^2, exciton_transfer.cpp:421-426) is invariant under it.
// Shared by the davoody table builder (csrc/davoody.cpp) and, so that both sides diagonalise identically, by the test-only
// Armadillo stand-in the reference's own sources are compiled against (oracle/arma_full/armadillo).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>
#include <vector>

namespace cntmc {

// A: n x n, row-major, Hermitian (only its values are read; a copy is rotated).  w: n eigenvalues ascending.  V: n x n row-major,
// column k = eigenvector k.  Returns the number of sweeps used (negative if the off-diagonal norm did not vanish in 100).
inline int hermitian_eig(int n, const std::complex<double>* A_in, double* w, std::complex<double>* V) {
  typedef std::complex<double> cd;
  std::vector<cd> A(A_in, A_in + (size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[(size_t)i * n + j] = (i == j) ? cd(1, 0) : cd(0, 0);
  auto at = [&](int i, int j) -> cd& { return A[(size_t)i * n + j]; };
  double scale = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) scale += std::norm(at(i, j));
  scale = std::sqrt(scale);
  int sweeps = 0;
  for (; sweeps < 100; ++sweeps) {
    double off = 0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += std::norm(at(p, q));
    if (std::sqrt(2 * off) <= 1e-17 * scale || off == 0) break;
    int rotations = 0;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const cd     apq = at(p, q);
        const double g = std::abs(apq);
        if (g == 0) continue;
        const double app = at(p, p).real(), aqq = at(q, q).real();
        if (g <= 1e-300) continue;
        // an element that no longer registers against either diagonal entry is done
        if (std::abs(app) + 100 * g == std::abs(app) && std::abs(aqq) + 100 * g == std::abs(aqq)) {
          at(p, q) = cd(0, 0);
          at(q, p) = cd(0, 0);
          continue;
        }
        ++rotations;
        // rotation that zeroes A[p][q]: phase e = apq/|apq|, angle from tan(2 theta) = 2|apq| / (aqq - app)
        const cd     e = apq / g;
        const double tau = (aqq - app) / (2 * g);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1 + tau * tau));
        const double c = 1 / std::sqrt(1 + t * t), s = t * c;
        // columns p, q of A and V:  [p q] <- [p q] * [[c, s e], [-s conj(e), c]]
        for (int k = 0; k < n; ++k) {
          const cd akp = at(k, p), akq = at(k, q);
          at(k, p) = c * akp - s * std::conj(e) * akq;
          at(k, q) = s * e * akp + c * akq;
          const cd vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * std::conj(e) * vkq;
          V[(size_t)k * n + q] = s * e * vkp + c * vkq;
        }
        // rows p, q of A:  [p; q] <- [[c, -s e], [s conj(e), c]] * [p; q]
        for (int k = 0; k < n; ++k) {
          const cd apk = at(p, k), aqk = at(q, k);
          at(p, k) = c * apk - s * e * aqk;
          at(q, k) = s * std::conj(e) * apk + c * aqk;
        }
        at(p, q) = cd(0, 0);
        at(q, p) = cd(0, 0);
        at(p, p) = cd(at(p, p).real(), 0);
        at(q, q) = cd(at(q, q).real(), 0);
      }
    }
    if (rotations == 0) break;
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return at(a, a).real() < at(b, b).real(); });
  std::vector<cd> Vs((size_t)n * n);
  for (int k = 0; k < n; ++k) {
    w[k] = at(order[k], order[k]).real();
    for (int i = 0; i < n; ++i) Vs[(size_t)i * n + k] = V[(size_t)i * n + order[k]];
  }
  std::copy(Vs.begin(), Vs.end(), V);
  return sweeps < 100 ? sweeps : -sweeps;
}

}  // namespace cntmc
