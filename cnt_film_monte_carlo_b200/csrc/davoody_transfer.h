// davoody_transfer.h -- host side of a donor -> acceptor transfer: everything of exciton_transfer::first_order
// (exciton_transfer.cpp:395-441) that does NOT depend on where the acceptor sits relative to the donor.
//
//   thermally relevant A2-singlet states of both tubes     get_relevant_states   exciton_transfer.cpp:226-271
//   partition function Z over the donor's states           first_order           exciton_transfer.cpp:409-412
//   energy-matched (donor, acceptor) state pairs           match_states          exciton_transfer.h:160-187
//   the k-space factor Q of every pair                     calculate_Q           exciton_transfer.cpp:274-304
//   plane-wave phases of every site for every distinct K_cm (the factors of calculate_J's inner loops, :363-374)
//
// The reference recomputes all of this for each of the table's 27 951 placements; it is the same every time.  What is left
// per placement is J = sum_ij conj-phase_i phase_j / |R_i - R_j| for the pair's two K_cm, over all N_donor x N_acceptor
// sites -- that is the GPU kernel's (davoody_kernels.cuh), which gets the flat arrays prepared here.
#ifndef CNTMC_DAVOODY_TRANSFER_H
#define CNTMC_DAVOODY_TRANSFER_H
#include "davoody_tube.h"

namespace cntmc {
namespace davoody {

struct ExState {
  int    ik_cm_idx, n;  // indices into Exciton::energy / psi
  int    ik_cm;
  double energy;
};

struct StatePair {
  int    donor, acceptor;  // indices into the sorted state lists
  int    kd, ka;           // indices into the lists of distinct K_cm of either tube
  cplx   Q;
  double boltzmann_rate;   // (2 pi / hbar) exp(-E_i / kT) / Z
  double lorentzian;       // of E_i - E_f
};

// the sites of a whole tube as calculate_J lays them out (make_Ru_3d / make_Ru_2d, exciton_transfer.cpp:309-352):
// cell after cell along the axis, centred, before the placement-dependent shift and rotation
struct TubeSites {
  std::vector<double> x, y_centred, z;  // on the cylinder; y is along the axis
  std::vector<Vec2>   sheet;            // on the unrolled sheet
  explicit TubeSites(const Tube& tube) {
    const int N = tube.n_sites();
    x.resize(N);
    y_centred.resize(N);
    z.resize(N);
    sheet.resize(N);
    for (int c = 0; c < tube.cells; c++)
      for (int j = 0; j < tube.Nu; j++) {
        const size_t i = (size_t)c * tube.Nu + j;
        x[i] = tube.cell_3d[j][0] + double(c) * 0.0;
        y_centred[i] = tube.cell_3d[j][1] + double(c) * tube.t3_y;
        z[i] = tube.cell_3d[j][2] + double(c) * 0.0;
        sheet[i] = Vec2{tube.site_a[j].x + double(c) * tube.t.x, tube.site_a[j].y + double(c) * tube.t.y};
      }
    const double y_max = *std::max_element(y_centred.begin(), y_centred.end());
    const double y_min = *std::min_element(y_centred.begin(), y_centred.end());
    const double mid = (y_max + y_min) / 2.;
    for (double& y : y_centred) y -= mid;
  }
};

class Transfer {
 public:
  const Tube& donor;
  const Tube& acceptor;
  double      temperature, broadening;
  std::vector<ExState>   d_states, a_states;
  std::vector<int>       d_kcm, a_kcm;  // distinct ik_cm among the states that appear in a pair
  std::vector<StatePair> pairs;
  TubeSites              d_sites, a_sites;
  std::vector<cplx>      d_phase;  // [kd][site]   exp(-i K_d . r_site)
  std::vector<cplx>      a_phase;  // [site][ka]   exp(+i K_a . r_site)
  double                 sqrt_lengths;  // sqrt(L_donor L_acceptor)

  Transfer(const Tube& d, const Tube& a, double temperature_kelvin, double broadening_joule)
      : donor(d), acceptor(a), temperature(temperature_kelvin), broadening(broadening_joule), d_sites(d), a_sites(a) {
    if (!(temperature > 0) || !(broadening > 0)) throw std::invalid_argument("temperature and broadening must be positive");
    const Exciton& dx = donor.excitons[kA2Singlet];
    const Exciton& ax = acceptor.excitons[kA2Singlet];
    const double   kT = Consts::kb() * temperature;
    const double   floor_energy = *std::min_element(dx.energy.begin(), dx.energy.end());  // the donor's, for both tubes (:400-408)
    d_states = relevant(dx, floor_energy, temperature);
    a_states = relevant(ax, floor_energy, temperature);
    double Z = 0;
    for (const ExState& s : d_states) Z += std::exp(-s.energy / kT);

    const double peak = lorentzian(0);
    const cplx   coeff((std::pow(Consts::q0(), 2) * donor.cell_area() * acceptor.cell_area()) /
                     (16 * std::pow(Consts::pi, 3) * Consts::eps0() * donor.radius * acceptor.radius *
                      std::sqrt(donor.length_in_meter() * acceptor.length_in_meter())));
    std::vector<cplx> dq(d_states.size()), aq(a_states.size());
    for (size_t i = 0; i < d_states.size(); i++) dq[i] = k_space_factor(donor, dx, d_states[i]);
    for (size_t i = 0; i < a_states.size(); i++) aq[i] = k_space_factor(acceptor, ax, a_states[i]);
    for (size_t i = 0; i < d_states.size(); i++)
      for (size_t f = 0; f < a_states.size(); f++) {
        const double gap = d_states[i].energy - a_states[f].energy;
        if (!(lorentzian(gap) > 1.e-2 * peak)) continue;
        StatePair p;
        p.donor = int(i);
        p.acceptor = int(f);
        p.kd = slot(d_kcm, d_states[i].ik_cm);
        p.ka = slot(a_kcm, a_states[f].ik_cm);
        p.Q = coeff * std::conj(dq[i]) * aq[f];
        p.boltzmann_rate = (2 * Consts::pi / Consts::hb()) * (std::exp(-d_states[i].energy / kT) / Z);
        p.lorentzian = lorentzian(gap);
        pairs.push_back(p);
      }
    sqrt_lengths = std::sqrt(donor.length_in_meter() * acceptor.length_in_meter());

    const cplx i1(0, 1);
    const int  Nd = donor.n_sites(), Na = acceptor.n_sites(), Kd = int(d_kcm.size()), Ka = int(a_kcm.size());
    d_phase.resize((size_t)Kd * Nd);
    a_phase.resize((size_t)Na * Ka);
    for (int k = 0; k < Kd; k++) {
      const Vec2 K = scaled(double(d_kcm[k]), donor.dk_l);
      for (int i = 0; i < Nd; i++) d_phase[(size_t)k * Nd + i] = std::exp(-i1 * dot2(K, d_sites.sheet[i]));
    }
    for (int k = 0; k < Ka; k++) {
      const Vec2 K = scaled(double(a_kcm[k]), acceptor.dk_l);
      for (int j = 0; j < Na; j++) a_phase[(size_t)j * Ka + k] = std::exp(+i1 * dot2(K, a_sites.sheet[j]));
    }
  }

  double lorentzian(double energy) const { return Consts::inv_pi() * broadening / (energy * energy + broadening * broadening); }

 private:
  static int slot(std::vector<int>& list, int ik_cm) {
    for (size_t i = 0; i < list.size(); i++)
      if (list[i] == ik_cm) return int(i);
    list.push_back(ik_cm);
    return int(list.size()) - 1;
  }

  // states whose thermal population relative to the floor exceeds 1e-3, by ascending energy
  static std::vector<ExState> relevant(const Exciton& ex, double floor_energy, double temperature) {
    const double         ceiling = floor_energy + std::abs(std::log(1.e-3) * Consts::kb() * temperature);
    std::vector<ExState> out;
    for (int n = 0; n < ex.n_principal; n++)
      for (int idx = 0; idx < ex.nk_cm; idx++) {
        const double e = ex.energy[(size_t)idx * ex.n_principal + n];
        if (e <= ceiling) out.push_back(ExState{idx, n, idx + ex.ik_cm_begin, e});
      }
    std::sort(out.begin(), out.end(), [](const ExState& s1, const ExState& s2) { return s1.energy < s2.energy; });
    return out;
  }

  // sum over the state's electron-hole pairs of psi * <conduction | e^{iK.r} | valence> on the two-atom cell (:277-296)
  static cplx k_space_factor(const Tube& tube, const Exciton& ex, const ExState& s) {
    const Vec2 K = scaled(double(s.ik_cm), tube.dk_l);
    const cplx i1(0., +1.);
    const cplx on_a = std::exp(i1 * dot2(K, Vec2{0, 0}));
    const cplx on_b = std::exp(i1 * dot2(K, tube.bond));
    const cplx*     psi = &ex.psi[((size_t)s.ik_cm_idx * ex.n_principal + s.n) * ex.nk_c];
    const uint32_t* idx = &tube.pair_index[(size_t)s.ik_cm_idx * ex.nk_c * 4];
    cplx            acc = 0;
    for (int p = 0; p < ex.nk_c; p++) {
      const int  ik_c = idx[p * 4 + 0], mu_c = idx[p * 4 + 1], ik_v = idx[p * 4 + 2], mu_v = idx[p * 4 + 3];
      const cplx ta = tube.w(ik_c, mu_c, 1, 0) * std::conj(tube.w(ik_v, mu_v, 0, 0)) * on_a;
      const cplx tb = tube.w(ik_c, mu_c, 1, 1) * std::conj(tube.w(ik_v, mu_v, 0, 1)) * on_b;
      acc += psi[p] * (ta + tb);
    }
    return acc;
  }
};

}  // namespace davoody
}  // namespace cntmc
#endif
