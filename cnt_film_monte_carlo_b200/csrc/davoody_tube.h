// davoody_tube.h -- host side of the "davoody" rate table: everything about ONE tube that does not depend on the relative
// placement of two tubes.  From a chirality (n,m) and a length in cnt unit cells it derives
//   geometry      cnt::get_parameters                cnt.cpp:16-117,   cnt::get_atom_coordinates  cnt.cpp:120-209
//   pi bands      cnt::electron_energy               cnt.cpp:302-360   (K2-extended zone, 2-atom cell, closed form)
//   valleys       cnt::find_valleys                  cnt.cpp:362-416,  cnt::find_relev_ik_range   cnt.cpp:419-517
//   v(q)          cnt::calculate_vq                  cnt.cpp:520-631   (Ohno potential, Fourier sum over the tube)
//   Pi(q), eps(q) cnt::calculate_polarization        cnt.cpp:634-732,  cnt::calculate_dielectric  cnt.cpp:735-780
//   excitons      cnt::calculate_A_excitons          cnt.cpp:783-1053  (Bethe-Salpeter kernels, 3 Hermitian problems per K_cm)
// in the order of cnt::calculate_exciton_dispersion (cnt.cpp:1056-1081).  This is set-up physics done once per chirality
// (milliseconds to seconds) and stays on the host; the geometry-dependent part (the N_site^2 Coulomb sums of
// exciton_transfer::calculate_J for every table entry) is the GPU's, in davoody_kernels.cuh.
//
// Layout differs from the reference (flat arrays indexed [mu][ik], no matrix library), the arithmetic does not: sums run in
// the reference's order and products associate the same way, so that the Hermitian matrices handed to the eigensolver agree
// with the reference's to the last bits and the comparison against oracle/_ref/libf1.so can be tight.  The eigensolver is
// herm_eig.h (cyclic Jacobi); the reference calls LAPACK through Armadillo, which this image does not have.
#ifndef CNTMC_DAVOODY_TUBE_H
#define CNTMC_DAVOODY_TUBE_H
#include <algorithm>
#include <array>
#include <cmath>
#include <atomic>
#include <complex>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "herm_eig.h"

namespace cntmc {
namespace davoody {

typedef std::complex<double> cplx;

// the reference's own rounded constants (exciton_transfer/constants.h:10-18); the table depends on them
struct Consts {
  static constexpr double pi = 3.141592;
  static double inv_pi() { return 1. / pi; }
  static double eV() { return 1.6 * std::pow(10, -19.0); }
  static double hb() { return 6.5 * std::pow(10, -16.0) * eV(); }
  static double kb() { return 1.3865 * std::pow(10, -23.0); }
  static double eps0() { return 8.85 * std::pow(10, -12); }
  static double q0() { return 1.6 * std::pow(10, -19); }
};

// independent iterations over host threads (each iteration owns its outputs, so the result does not depend on the split)
template <typename F>
inline void parallel_for(int n, int grain, F&& body) {
  const int workers = std::max(1, std::min<int>(n / std::max(1, grain), (int)std::thread::hardware_concurrency()));
  if (workers <= 1) {
    for (int i = 0; i < n; i++) body(i);
    return;
  }
  std::vector<std::thread> pool;
  for (int w = 0; w < workers; w++)
    pool.emplace_back([&, w] {
      for (int i = w; i < n; i += workers) body(i);
    });
  for (auto& th : pool) th.join();
}

struct Vec2 {
  double x, y;
};
// Armadillo's dot/norm keep two partial sums (even and odd elements); for two elements that is x*x' + y*y'
inline double dot2(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline double norm2(const Vec2& a) { return std::sqrt(a.x * a.x + a.y * a.y); }
inline Vec2   scaled(double k, const Vec2& a) { return Vec2{k * a.x, k * a.y}; }
inline Vec2   plus(const Vec2& a, const Vec2& b) { return Vec2{a.x + b.x, a.y + b.y}; }
inline Vec2   rotated(const double r[2][2], const Vec2& a) { return Vec2{r[0][0] * a.x + r[0][1] * a.y, r[1][0] * a.x + r[1][1] * a.y}; }

enum ExcitonKind { kA1 = 0, kA2Singlet = 1, kA2Triplet = 2 };

struct Exciton {
  int                 nk_cm = 0, n_principal = 0, nk_c = 0;  // K_cm points, states per K_cm, electron-hole pairs per state
  int                 ik_cm_begin = 0;                       // ik_cm of index 0
  std::vector<double> energy;                                // [ik_cm_idx][n]
  std::vector<cplx>   psi;                                   // [ik_cm_idx][n][pair]
};

class Tube {
 public:
  int n = 0, m = 0, cells = 0;
  // lattice
  Vec2   a1, a2, b1, b2, bond, ch, t, K1, K2, dk_l;
  double ch_len = 0, radius = 0, t3_y = 0;
  int    t1 = 0, t2 = 0, Nu = 0, M = 0, Q = 0;
  std::vector<Vec2> site_a, site_b;  // the two sublattices of one cnt unit cell, unrolled sheet
  std::vector<std::array<double, 3>> cell_3d;  // graphene unit cells (A sites) of one cnt unit cell on the cylinder
  // pi bands on the K2-extended zone: mu in [0,Q), ik in [0,nk)
  int                 nk = 0;
  std::vector<double> e_cond;  // [mu][ik]; the valence band is its negative
  std::vector<cplx>   wf;      // [mu][ik][band 0=v,1=c][atom 0=A,1=B]
  // lowest pair of valleys and the states within 1 eV of their bottoms
  std::vector<std::array<std::array<int, 2>, 2>> valleys;  // {ik, mu} x 2, sorted by energy
  std::vector<std::array<int, 2>>                relev[2];
  // screened interaction on iq in [-(nk-1), nk), mu in [-(Q-1), Q)
  int                 iq_begin = 0, nq = 0, mu_begin = 0, n_muq = 0;
  std::vector<cplx>   vq;   // [iq][mu][4]  (aa, ab, ba, bb)
  std::vector<double> pol;  // [iq][mu]
  std::vector<double> eps;  // [iq][mu]
  // excitons
  Exciton               excitons[3];
  std::vector<uint32_t> pair_index;  // [ik_cm_idx][pair][4] = ik_c, mu_c, ik_v, mu_v (indices into the band arrays)

  Tube(int n_, int m_, int cells_) : n(n_), m(m_), cells(cells_) {
    if (n <= 0 || m < 0 || m > n) throw std::invalid_argument("chirality (n,m) needs n >= m >= 0 and n > 0");
    if (cells <= 0) throw std::invalid_argument("tube length must be at least one cnt unit cell");
    lattice();
    sites();
    bands();
    find_valleys();
    relevant_states(1. * Consts::eV());
    coulomb();
    polarization();
    dielectric();
    solve_excitons();
  }

  double length_in_meter() const { return double(cells) * norm2(t); }
  double cell_area() const { return std::abs(a1.x * a2.y - a1.y * a2.x); }
  int    n_sites() const { return Nu * cells; }

  double        cond(int ik, int mu) const { return e_cond[(size_t)mu * nk + ik]; }
  const cplx&   w(int ik, int mu, int band, int atom) const { return wf[(((size_t)mu * nk + ik) * 2 + band) * 2 + atom]; }
  const cplx&   v_q(int iq, int mu, int c) const { return vq[(((size_t)(iq - iq_begin)) * n_muq + (mu - mu_begin)) * 4 + c]; }
  double        eps_q(int iq, int mu) const { return eps[(size_t)(iq - iq_begin) * n_muq + (mu - mu_begin)]; }

 private:
  static constexpr double a_cc = 1.42e-10;

  void lattice() {  // cnt.cpp:16-117
    const double pi = Consts::pi;
    const double a_l = std::sqrt(float(3.0)) * a_cc;  // single-precision root, as the reference has it (cnt.h:26)
    a1 = Vec2{a_l * std::sqrt(3.0) / 2.0, +a_l / 2.0};
    a2 = Vec2{a_l * std::sqrt(3.0) / 2.0, -a_l / 2.0};
    b1 = Vec2{1.0 / std::sqrt(3.0) * 2.0 * pi / a_l, +2.0 * pi / a_l};
    b2 = Vec2{1.0 / std::sqrt(3.0) * 2.0 * pi / a_l, -2.0 * pi / a_l};
    bond = scaled(1.0 / 3.0, plus(a1, a2));
    ch = plus(scaled(double(n), a1), scaled(double(m), a2));
    ch_len = norm2(ch);
    radius = ch_len / 2.0 / pi;
    const int dR = std::gcd(2 * n + m, n + 2 * m);
    t1 = +(2 * m + n) / dR;
    t2 = -(2 * n + m) / dR;
    t = plus(scaled(double(t1), a1), scaled(double(t2), a2));
    Nu = int(2 * (std::pow(n, 2) + std::pow(m, 2) + n * m) / dR);
    // turn the sheet so that the chiral vector is the x axis and the tube axis the y axis
    const double c = ch.x / norm2(ch), s = ch.y / norm2(ch);
    const double rot[2][2] = {{+c, +s}, {-s, +c}};
    ch = rotated(rot, ch);
    t = rotated(rot, t);
    a1 = rotated(rot, a1);
    a2 = rotated(rot, a2);
    b1 = rotated(rot, b1);
    b2 = rotated(rot, b2);
    bond = rotated(rot, bond);
    t3_y = t.y;
    // reciprocal vectors of the tube: K1 around, K2 along; dk_l is the k step the finite length allows
    const Vec2 k1n = plus(scaled(-double(t2), b1), scaled(double(t1), b2));
    K1 = Vec2{k1n.x / double(Nu), k1n.y / double(Nu)};
    const Vec2 k2n = Vec2{double(m) * b1.x - double(n) * b2.x, double(m) * b1.y - double(n) * b2.y};
    K2 = Vec2{k2n.x / double(Nu), k2n.y / double(Nu)};
    dk_l = Vec2{K2.x / double(cells), K2.y / double(cells)};
    // K2-extended representation: the M for which Q = gcd(Nu, M) cutting lines of length Nu/Q cover the zone
    const double slope = double(m) / double(n) - double(t2) / double(t1);
    const double p_min = (1. / double(t1) + 1. / double(n)) / slope;
    const double p_max = (1. / double(t1) + double(Nu) / double(n)) / slope;
    bool         found = false;
    for (int p = int(std::ceil(p_min)); p < std::ceil(p_max); p++)
      if (((1 + t2 * p) % t1) == 0) {
        const int q = (1 + t2 * p) / t1;
        M = m * p - n * q;
        Q = std::gcd(Nu, M);
        found = true;
        break;
      }
    if (!found) throw std::runtime_error("no K2-extended representation for this chirality");
  }

  void sites() {  // cnt.cpp:120-209
    site_a.assign(Nu, Vec2{0, 0});
    site_b.assign(Nu, Vec2{0, 0});
    auto fold = [&](Vec2& p) {
      if (p.x > ch.x) p.x -= ch.x;
      if (p.x < 0.0) p.x += ch.x;
      if (p.y > ch.y) p.y -= ch.y;
      if (p.y < 0.0) p.y += ch.y;
    };
    int k = 0;
    for (int i = 0; i <= t1 + n; i++)
      for (int j = t2; j <= m; j++) {
        const bool above_t = double(t2 * i) / double(t1) <= double(j);
        const bool below_ch = double(m * i) / double(n) >= double(j);
        const bool left_of_t = double(t2 * (i - n)) / double(t1) > double(j - m);
        const bool under_ch = double(m * (i - t1)) / double(n) < double(j - t2);
        if (!(above_t && below_ch && left_of_t && under_ch)) continue;
        if (k >= Nu) throw std::runtime_error("more sites in the cnt unit cell than Nu");
        Vec2 pa = plus(scaled(double(i), a1), scaled(double(j), a2));
        Vec2 pb = plus(pa, bond);
        fold(pa);
        fold(pb);
        site_a[k] = pa;
        site_b[k] = pb;
        k++;
      }
    if (k != Nu) throw std::runtime_error("fewer sites in the cnt unit cell than Nu");
    cell_3d.resize(Nu);
    for (int i = 0; i < Nu; i++)
      cell_3d[i] = {radius * std::cos(site_a[i].x / radius), site_a[i].y, radius * std::sin(site_a[i].x / radius)};
  }

  void bands() {  // cnt.cpp:302-360
    nk = Nu / Q * cells;
    e_cond.assign((size_t)Q * nk, 0.0);
    wf.assign((size_t)Q * nk * 4, cplx(0, 0));
    const double t0 = 2.7 * Consts::eV();
    const Vec2   d1 = Vec2{(a1.x + a2.x) / 3., (a1.y + a2.y) / 3.};
    const Vec2   d2 = Vec2{(a1.x - 2. * a2.x) / 3., (a1.y - 2. * a2.y) / 3.};
    const Vec2   d3 = Vec2{(a2.x - 2. * a1.x) / 3., (a2.y - 2. * a1.y) / 3.};
    for (int mu = 0; mu < Q; mu++)
      for (int ik = 0; ik < nk; ik++) {
        const Vec2 k = plus(scaled(double(mu), K1), scaled(double(ik), dk_l));
        const cplx fk = std::exp(cplx(0, dot2(k, d1))) + std::exp(cplx(0, dot2(k, d2))) + std::exp(cplx(0, dot2(k, d3)));
        e_cond[(size_t)mu * nk + ik] = +t0 * std::abs(fk);
        cplx* out = &wf[((size_t)mu * nk + ik) * 4];
        out[0 * 2 + 0] = +1. / std::sqrt(2.);                               // valence, A
        out[0 * 2 + 1] = +1. / std::sqrt(2.) * std::conj(fk) / std::abs(fk);  // valence, B
        out[1 * 2 + 0] = +1. / std::sqrt(2.);                               // conduction, A
        out[1 * 2 + 1] = -1. / std::sqrt(2.) * std::conj(fk) / std::abs(fk);  // conduction, B
      }
  }

  void find_valleys() {  // cnt.cpp:362-416
    std::vector<std::array<int, 2>> minima;
    for (int ik = 0; ik < nk; ik++)
      for (int mu = 0; mu < Q; mu++) {
        const int up = (ik + 1) % nk, down = (ik - 1 + nk) % nk;
        if (cond(ik, mu) < cond(down, mu) && cond(ik, mu) < cond(up, mu)) minima.push_back({ik, mu});
      }
    std::sort(minima.begin(), minima.end(), [&](const std::array<int, 2>& s1, const std::array<int, 2>& s2) { return cond(s1[0], s1[1]) < cond(s2[0], s2[1]); });
    for (size_t i = 0; i < minima.size() / 2; i++) valleys.push_back({minima[2 * i], minima[2 * i + 1]});
    if (valleys.empty()) throw std::runtime_error("no valley pair in the conduction band");
  }

  void relevant_states(double window) {  // cnt.cpp:419-517: grow outwards from each valley bottom while either side stays inside the window
    for (int v = 0; v < 2; v++) {
      const int    ik0 = valleys[0][v][0], mu0 = valleys[0][v][1];
      const double ceiling = cond(ik0, mu0) + window;
      auto&        list = relev[v];
      list.clear();
      list.push_back({ik0, mu0});
      bool inside = true;
      for (int count = 1; inside && count < nk; count++) {
        inside = false;
        const int hi = (ik0 + count) % nk;
        if (cond(hi, mu0) < ceiling) {
          list.push_back({hi, mu0});
          inside = true;
        }
        const int lo = ((ik0 - count) % nk + nk) % nk;
        if (cond(lo, mu0) < ceiling) {
          list.insert(list.begin(), {lo, mu0});
          inside = true;
        }
      }
    }
    if (relev[0].size() != relev[1].size()) throw std::runtime_error("the two valleys hold different numbers of relevant states");
  }

  void coulomb() {  // cnt.cpp:520-631
    iq_begin = -(nk - 1);
    nq = 2 * nk - 1;
    mu_begin = -(Q - 1);
    n_muq = 2 * Q - 1;
    int span = cells;
    if (span % 2 == 0) span++;
    const int half = span / 2;
    // separations between every site of the (odd number of) cells and the first A or B site, folded around the tube
    std::vector<Vec2> rel[4];
    for (int c = 0; c < 4; c++) rel[c].resize((size_t)Nu * span);
    for (int i = -half; i <= half; i++)
      for (int j = 0; j < Nu; j++) {
        Vec2 d[4] = {Vec2{site_a[j].x - site_a[0].x, site_a[j].y - site_a[0].y}, Vec2{site_a[j].x - site_b[0].x, site_a[j].y - site_b[0].y},
                     Vec2{site_b[j].x - site_a[0].x, site_b[j].y - site_a[0].y}, Vec2{site_b[j].x - site_b[0].x, site_b[j].y - site_b[0].y}};
        for (int c = 0; c < 4; c++) {
          if (d[c].x > ch.x / 2) d[c].x -= ch.x;
          rel[c][(size_t)(i + half) * Nu + j] = Vec2{d[c].x + double(i) * t.x, d[c].y + double(i) * t.y};
        }
      }
    const double Upp = 11.3 * Consts::eV();
    const double coeff = std::pow(4. * Consts::pi * Consts::eps0() * Upp / Consts::q0() / Consts::q0(), 2);
    const double count = double(2 * Nu * span);
    vq.assign((size_t)nq * n_muq * 4, cplx(0, 0));
    parallel_for(nq, 16, [&](int iq_idx) {
      const int iq = iq_begin + iq_idx;
      for (int mu_idx = 0; mu_idx < n_muq; mu_idx++) {
        const int  mu = mu_begin + mu_idx;
        const Vec2 q = plus(scaled(double(iq), dk_l), scaled(double(mu), K1));
        for (int c = 0; c < 4; c++) {
          cplx acc(0, 0);
          for (size_t k = 0; k < rel[c].size(); k++) {
            const Vec2& R = rel[c][k];
            acc += std::exp(cplx(0., 1.) * dot2(q, R)) * Upp / std::sqrt(coeff * (std::pow(R.x, 2) + std::pow(R.y, 2)) + 1);
          }
          vq[((size_t)iq_idx * n_muq + mu_idx) * 4 + c] = acc / cplx(count);
        }
      }
    });
  }

  // k+q folded back into the K2-extended zone: a step across the last cutting line shifts ik by cells*M
  void fold_kq(int ik, int mu_k, int iq, int mu_q, int* ikq, int* mu_kq) const {
    int mu = mu_k + mu_q, k = ik + iq;
    while (mu >= Q) {
      mu -= Q;
      k += cells * M;
    }
    while (mu < 0) {
      mu += Q;
      k -= cells * M;
    }
    while (k >= nk) k -= nk;
    while (k < 0) k += nk;
    *ikq = k;
    *mu_kq = mu;
  }

  void polarization() {  // cnt.cpp:634-732
    pol.assign((size_t)nq * n_muq, 0.0);
    auto overlap2 = [&](int ik, int mu, int band, int ik2, int mu2, int band2) {
      // |<k,band | k2,band2>|^2 over the two atoms of the graphene cell
      const cplx s = std::conj(w(ik, mu, band, 0)) * w(ik2, mu2, band2, 0) + std::conj(w(ik, mu, band, 1)) * w(ik2, mu2, band2, 1);
      return std::pow(std::abs(s), 2);
    };
    parallel_for(nq, 16, [&](int iq_idx) {
      const int iq = iq_begin + iq_idx;
      for (int mu_idx = 0; mu_idx < n_muq; mu_idx++) {
        const int mu_q = mu_begin + mu_idx;
        double    acc = 0;
        for (int ik = 0; ik < nk; ik++)
          for (int mu_k = 0; mu_k < Q; mu_k++) {
            int ikq, mu_kq;
            fold_kq(ik, mu_k, iq, mu_q, &ikq, &mu_kq);
            acc += overlap2(ik, mu_k, 0, ikq, mu_kq, 1) / (cond(ikq, mu_kq) - (-cond(ik, mu_k))) +
                   overlap2(ik, mu_k, 1, ikq, mu_kq, 0) / (cond(ik, mu_k) - (-cond(ikq, mu_kq)));
          }
        pol[(size_t)iq_idx * n_muq + mu_idx] = 2 * acc;
      }
    });
  }

  void dielectric() {  // cnt.cpp:735-780: eps = 1 + Re<v(q)> Pi(q), <> the mean over the four sublattice pairs
    eps.assign((size_t)nq * n_muq, 0.0);
    for (size_t i = 0; i < eps.size(); i++) {
      cplx acc(0, 0);
      for (int c = 0; c < 4; c++) acc += vq[i * 4 + c];
      const double mean_re = (acc / cplx(4.0)).real();
      eps[i] = mean_re * pol[i] + 1.;
    }
  }

  static int wrap(int k, int nk_) {
    while (k >= nk_) k -= nk_;
    while (k < 0) k += nk_;
    return k;
  }

  void solve_excitons() {  // cnt.cpp:783-1053
    const int nr = int(relev[0].size());
    const int nk_cm = 2 * nr, nk_c = 2 * nr;
    for (auto& ex : excitons) {
      ex.nk_cm = nk_cm;
      ex.n_principal = nr;
      ex.nk_c = nk_c;
      ex.ik_cm_begin = -nr;
      ex.energy.assign((size_t)nk_cm * nr, 0.0);
      ex.psi.assign((size_t)nk_cm * nr * nk_c, cplx(0, 0));
    }
    pair_index.assign((size_t)nk_cm * nk_c * 4, 0u);

    struct Pair {
      int ik_c, mu_c, ik_v, mu_v;
    };
    // screened direct term between the pairs (c,v) and (c',v'): momentum transfer k_c - k_c'
    auto direct = [&](const Pair& p, const Pair& pp) {
      const int dk = wrap_signed(p.ik_c - pp.ik_c), dmu = p.mu_c - pp.mu_c;
      cplx      acc = 0;
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          acc += std::conj(w(p.ik_c, p.mu_c, 1, i)) * w(p.ik_v, p.mu_v, 0, j) * w(pp.ik_c, pp.mu_c, 1, i) * std::conj(w(pp.ik_v, pp.mu_v, 0, j)) *
                 v_q(dk, dmu, 2 * i + j) / eps_q(dk, dmu);
      return acc;
    };
    // bare exchange term: momentum transfer K_cm
    auto exchange = [&](const Pair& p, const Pair& pp, int ik_cm) {
      cplx acc = 0;
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          acc += std::conj(w(p.ik_c, p.mu_c, 1, i)) * w(p.ik_v, p.mu_v, 0, i) * w(pp.ik_c, pp.mu_c, 1, j) * std::conj(w(pp.ik_v, pp.mu_v, 0, j)) *
                 v_q(ik_cm, 0, 2 * i + j);
      return acc;
    };

    const double inv_sqrt2 = 1 / std::sqrt(2.);
    std::atomic<bool> failed(false);

    // the 2 nr centre-of-mass momenta are independent problems (three eigen-decompositions each): one host thread each, every
    // iteration writes its own slices of the outputs, so the bits do not depend on how the iterations are split
    parallel_for(nk_cm, 1, [&](int idx) {
      const int ik_cm = idx - nr;
      std::vector<cplx>   k11((size_t)nr * nr, cplx(0, 0)), k12((size_t)nr * nr, cplx(0, 0)), kx((size_t)nr * nr, cplx(0, 0)), H((size_t)nr * nr),
          V((size_t)nr * nr);
      std::vector<double> E(nr);
      // valley 1 carries the electron index, valley 2 the hole index (its electron follows from K_cm)
      auto pair_v1 = [&](int i) {
        const int ik_c = relev[0][i][0], mu = relev[0][i][1];
        return Pair{ik_c, mu, wrap(ik_c - ik_cm, nk), mu};
      };
      auto pair_v2 = [&](int i) {
        const int ik_v = relev[1][i][0], mu = relev[1][i][1];
        return Pair{wrap(ik_v + ik_cm, nk), mu, ik_v, mu};
      };
      for (int a = 0; a < nr; a++) {
        const Pair p = pair_v1(a);
        k11[(size_t)a * nr + a] += cond(p.ik_c, p.mu_c) - (-cond(p.ik_v, p.mu_c));
        for (int b = 0; b <= a; b++) {
          const Pair pp = pair_v1(b);
          k11[(size_t)a * nr + b] -= direct(p, pp);
          kx[(size_t)a * nr + b] += cplx(2, 0) * exchange(p, pp, ik_cm);
        }
        for (int b = a; b < nr; b++) {
          const Pair pp = pair_v2(b);
          k12[(size_t)a * nr + (nr - 1 - b)] -= direct(p, pp);
        }
      }
      // the loops above fill one triangle; K + K^H doubles the diagonal, which is then halved
      auto symmetrise = [&](std::vector<cplx>& K) {
        std::vector<cplx> Kh((size_t)nr * nr);
        for (int a = 0; a < nr; a++)
          for (int b = 0; b < nr; b++) Kh[(size_t)a * nr + b] = std::conj(K[(size_t)b * nr + a]);
        for (size_t i = 0; i < K.size(); i++) K[i] += Kh[i];
        for (int a = 0; a < nr; a++) K[(size_t)a * nr + a] /= cplx(2, 0);
      };
      symmetrise(k11);
      symmetrise(k12);
      symmetrise(kx);

      auto solve = [&](ExcitonKind kind, double sign_tail) {
        if (hermitian_eig(nr, H.data(), E.data(), V.data()) < 0) failed = true;
        Exciton& ex = excitons[kind];
        for (int s = 0; s < nr; s++) {
          ex.energy[(size_t)idx * nr + s] = E[s];
          cplx* psi = &ex.psi[((size_t)idx * nr + s) * nk_c];
          for (int a = 0; a < nr; a++) {
            const cplx v = V[(size_t)a * nr + s];
            psi[a] = cplx(+inv_sqrt2 * v.real(), +inv_sqrt2 * v.imag());
            psi[nr + a] = cplx(sign_tail * inv_sqrt2 * v.real(), sign_tail * inv_sqrt2 * v.imag());
          }
        }
      };
      for (size_t i = 0; i < H.size(); i++) H[i] = k11[i] - k12[i];
      solve(kA1, -1.0);
      for (size_t i = 0; i < H.size(); i++) H[i] = k11[i] + k12[i];
      solve(kA2Triplet, +1.0);
      for (size_t i = 0; i < H.size(); i++) H[i] = k11[i] + k12[i] + cplx(2, 0) * kx[i];
      solve(kA2Singlet, +1.0);

      uint32_t* pi = &pair_index[(size_t)idx * nk_c * 4];
      for (int a = 0; a < nr; a++) {
        const Pair p = pair_v1(a), r = pair_v2(a);
        const uint32_t first[4] = {(uint32_t)p.ik_c, (uint32_t)p.mu_c, (uint32_t)p.ik_v, (uint32_t)p.mu_v};
        const uint32_t second[4] = {(uint32_t)r.ik_c, (uint32_t)r.mu_c, (uint32_t)r.ik_v, (uint32_t)r.mu_v};
        for (int c = 0; c < 4; c++) {
          pi[(size_t)a * 4 + c] = first[c];
          pi[(size_t)(nk_c - 1 - a) * 4 + c] = second[c];
        }
      }
    });
    if (failed) throw std::runtime_error("exciton eigenproblem did not converge");
  }

  // k_c - k_c' folded into [0, nk): the direct term looks v(q) up at that index (the band range starts at 0)
  int wrap_signed(int k) const { return wrap(k, nk); }
};

}  // namespace davoody
}  // namespace cntmc
#endif
