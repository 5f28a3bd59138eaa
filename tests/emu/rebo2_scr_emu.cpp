// Serial host run of the screened-REBO2 per-atom device functions
// (atomistica_b200/csrc/atx_rebo2_scr.cuh) on a neighbour list in the DEVICE format (pos4, CSR seed,
// int2 entries with packed shifts, order).  Mirrors what k_rbs_bonds / k_rbs_force / k_rbs_scr and
// rebo2_scr_compute of atx_rebo2.cu do, one atom after the other.  Test infrastructure only.
#include "device_shim.h"

#include <cstring>
#include <vector>

#include "atx_rebo2_scr.cuh"
#include "atx_rebo2_atom.cuh"

extern "C" int emu_rebo2_scr(const atx_rebo2_params *par, const atx_rebo2_screening *scr, const int *el2typ,
                             int nat, int nbs, int nss, const double *Abox, const double *pos4_,
                             const long long *seed, const int *list_, const int *order, double *sums,
                             double *f, double *epa, double *wpa, double *epb, double *fpb, double *wpb,
                             int *stats /* bonds, screening entries, max of each per atom */) {
  Rebo2Dev P;
  std::memset(&P, 0, sizeof(P));
  rb_fill_dev(P, par);
  P.Fcc = par->Fcc; P.Fch = par->Fch; P.Fhh = par->Fhh; P.Tcc = par->Tcc; P.Pcc = par->Pcc; P.Pch = par->Pch;
  for (int k = 0; k < 32; k++) P.el2typ[k] = el2typ[k];
  RbsCut S;
  rbs_fill_cut(S, P, scr);
  Mat3 A;
  for (int k = 0; k < 9; k++) A.m[k] = Abox[k];
  const double4 *pos4 = reinterpret_cast<const double4 *>(pos4_);
  const int2 *list = reinterpret_cast<const int2 *>(list_);
  if (nbs > RBS_NBL) nbs = RBS_NBL;
  const size_t nt = (size_t)nat * nbs + 1, ns = (size_t)nat * nss + 1;
  std::vector<int> b_cnt(nat + 1), b_nb(nt), b_typ(nt), b_shift(nt), b_slot(nt), b_sseed(nt), b_scnt(nt), s_ent(ns);
  std::vector<double4> b_vec(nt);
  std::vector<double2> b_car(nt), b_cbo(nt), b_cnc(nt), nn(nat + 1);
  // poison the screening tables: every entry that is read must have been written by loop 1
  std::vector<double> s_arik(ns, NAN), s_arjk(ns, NAN), s_boik(ns, NAN), s_bojk(ns, NAN), s_ncik(ns, NAN),
      s_ncjk(ns, NAN), s_facbo(ns, NAN), s_facnc(ns, NAN);
  int flag = 0;
  RbsTab T;
  T.nat = nat; T.nbs = nbs; T.nss = nss;
  T.b_cnt = b_cnt.data(); T.b_nb = b_nb.data(); T.b_typ = b_typ.data(); T.b_shift = b_shift.data();
  T.b_slot = b_slot.data(); T.b_sseed = b_sseed.data(); T.b_scnt = b_scnt.data(); T.b_vec = b_vec.data();
  T.b_car = b_car.data(); T.b_cbo = b_cbo.data(); T.b_cnc = b_cnc.data(); T.nn = nn.data();
  T.s_ent = s_ent.data(); T.s_arik = s_arik.data(); T.s_arjk = s_arjk.data(); T.s_boik = s_boik.data();
  T.s_bojk = s_bojk.data(); T.s_ncik = s_ncik.data(); T.s_ncjk = s_ncjk.data(); T.s_facbo = s_facbo.data();
  T.s_facnc = s_facnc.data(); T.flag = &flag;
  for (int s = 0; s < nat; s++) rbs_bonds_atom(T, A, P, S, pos4, seed, list, s);
  for (int k = 0; k < 4; k++) stats[k] = 0;
  for (int s = 0; s < nat && !flag; s++) {
    int nsa = 0;
    for (int b = 0; b < b_cnt[s]; b++) nsa += b_scnt[(size_t)s * nbs + b];
    stats[0] += b_cnt[s];
    stats[1] += nsa;
    if (b_cnt[s] > stats[2]) stats[2] = b_cnt[s];
    if (nsa > stats[3]) stats[3] = nsa;
  }
  if (flag) return flag;
  double acc[RBS_NSUM];
  for (int k = 0; k < RBS_NSUM; k++) acc[k] = 0.0;
  for (int i = 0; i < nat; i++) rbs_force_atom(T, A, P, S, pos4, seed, list, order, f, epa, wpa, epb, fpb, wpb, i, acc);
  for (int i = 0; i < nat; i++) rbs_scr_atom(T, A, P, pos4, seed, list, f, wpa, wpb, i, acc);
  for (int k = 0; k < RBS_NSUM; k++) sums[k] = acc[k];
  return 0;
}


// Unscreened REBO2: serial run of rb_bonds_atom / rb_force_atom (the bodies of k_rebo2_bonds /
// k_rebo2_force), same inputs and outputs as above.
extern "C" int emu_rebo2(const atx_rebo2_params *par, const int *el2typ, int nat, int nbs, const double *Abox,
                         const double *pos4_, const long long *seed, const int *list_, const int *order,
                         double *sums, double *f, double *epa, double *wpa, double *epb, double *fpb,
                         double *wpb, const unsigned char *role /* nullptr, or 2 owned / 1 ghost */,
                         int per_bond /* k_rebo2_own_count / scan / k_rebo2_own_fill / k_rebo2_force_bond */) {
  Rebo2Dev P;
  std::memset(&P, 0, sizeof(P));
  rb_fill_dev(P, par);
  P.Fcc = par->Fcc; P.Fch = par->Fch; P.Fhh = par->Fhh; P.Tcc = par->Tcc; P.Pcc = par->Pcc; P.Pch = par->Pch;
  for (int k = 0; k < 32; k++) P.el2typ[k] = el2typ[k];
  Mat3 A;
  for (int k = 0; k < 9; k++) A.m[k] = Abox[k];
  const double4 *pos4 = reinterpret_cast<const double4 *>(pos4_);
  const int2 *list = reinterpret_cast<const int2 *>(list_);
  if (nbs > RB_NBL) nbs = RB_NBL;
  const size_t nt = (size_t)nat * nbs + 1;
  std::vector<int> b_cnt(nat + 1);
  std::vector<RbBond> b_tab(nt);
  std::vector<double2> nn(nat + 1);
  int flag = 0;
  for (int s = 0; s < nat; s++)
    rb_bonds_atom(nbs, A, P, pos4, seed, list, b_cnt.data(), b_tab.data(), nn.data(), &flag, s);
  if (flag) return flag;
  double acc[RBS_NSUM];
  for (int k = 0; k < RBS_NSUM; k++) acc[k] = 0.0;
  if (role) {
    // k_rebo2_force_roles + k_rebo2_clear_ghosts
    for (int i = 0; i < nat; i++)
      rb_force_atom<true>(nat, nbs, P, seed, b_cnt.data(), b_tab.data(), nn.data(), pos4, order, f, epa, wpa, nullptr,
                          nullptr, nullptr, i, acc, role);
    for (int s = 0; s < nat; s++) {
      if (role[s] >= 2) continue;
      f[3 * s] = f[3 * s + 1] = f[3 * s + 2] = 0.0;
      if (epa) epa[s] = 0.0;
      if (wpa)
        for (int q = 0; q < 9; q++) wpa[9 * s + q] = 0.0;
    }
  } else if (per_bond) {
    std::vector<int> cnt(nat + 1, 0), off(nat + 2, 0);
    for (int s = 0; s < nat; s++)
      cnt[s] = rb_owned_bonds(nbs, P, b_cnt.data(), b_tab.data(), pos4,
                              order, s, nullptr);
    for (int s = 0; s <= nat; s++) off[s + 1] = off[s] + cnt[s];     // exclusive scan over nat + 1 inputs
    std::vector<int2> own(off[nat] + 1);
    if ((size_t)off[nat] > (size_t)nat * nbs / 2 + 1) return -7;      // the bound the host code sizes with
    for (int s = 0; s < nat; s++)
      rb_owned_bonds(nbs, P, b_cnt.data(), b_tab.data(), pos4, order, s,
                     own.data() + off[s]);
    for (int t = 0; t < off[nat]; t++)
      rb_force_atom<false, true>(nat, nbs, P, seed, b_cnt.data(), b_tab.data(), nn.data(), pos4, order, f, epa, wpa, epb,
                                 fpb, wpb, own[t].x, acc, nullptr, own[t].y);
  } else {
    for (int i = 0; i < nat; i++)
      rb_force_atom(nat, nbs, P, seed, b_cnt.data(), b_tab.data(), nn.data(), pos4, order, f, epa, wpa, epb, fpb, wpb, i, acc);
  }
  for (int k = 0; k < RBS_NSUM; k++) sums[k] = acc[k];
  return 0;
}
