// Screened REBO2 (Rebo2Scr) per-atom device functions.
//
// Replaces rebo2_kernel of src/potentials/bop/rebo2/bop_kernel_rebo2.f90 compiled with SCREENING
// NUM_NEIGHBORS and ALT_DIHEDRAL (rebo2_scr.f90:60-64; with_dihedral switches the dihedral term of this
// build, :2089-2371, on):
//   rbs_bonds_atom  loop 1 (:700-1181): bond table with the three cutoff families ar / bo / nc
//                   (attractive-repulsive, bond order, neighbour count), the Baskes screening
//                   function of the C-C bonds beyond the inner cutoff and the per-bond lists of
//                   screening neighbours; nn(C,H) (:1189-1200).
//   rbs_force_atom  loop 2 (:1209-2781): energies and forces, accumulation of the screening force
//                   factors sfacbo / sfacnc (:2611-2712) and the screening forces of the
//                   attractive/repulsive cutoff (:2717-2757).
//   rbs_scr_atom    loop 3 (:2783-2871): forces from the screening of the bond-order and
//                   neighbour-count cutoffs.
// As in the unscreened kernels (atx_rebo2.cu) the nebmax^2 second-neighbour caches of the
// reference are not materialised: the terms of k's and l's neighbours are recomputed from the bond
// table when their forces are applied.
//
// The functions only use plain pointers and the helpers named below, so that the test-suite can
// compile this header with a host compiler (tests/emu/) and run the same source serially against
// the CPU oracle.  Needs: atx_rebo2_func.cuh, Mat3, atx_unpack_shift, atx_image_vector,
// ATX_NONZERO_SHIFT, double2/double4/int2, RBS_ADD(ptr, val) (atomicAdd on the device).
#pragma once

#include "atx_rebo2_func.cuh"

#define RBS_NBL 24   // per-thread bond scratch of the screened kernels
#define RBS_NSUM 10  // epot + virial, same layout as ATX_NSUM

struct RbsCut {  // index = pair type (RB_CC, RB_CH, RB_HH), as Rebo2Dev::cut_*
  double ar_l[7], ar_h[7], bo_l[7], bo_h[7], nc_l[7], nc_h[7], max_cut_sq[7];
  double Cmin, Cmax, dC, C_dr_cut;
};

struct RbsTab {
  int nat, nbs, nss;  // atoms, bond stride per atom, screening-entry stride per atom
  // bond table, entry (size_t)i * nbs + b
  int *b_cnt, *b_nb, *b_typ, *b_shift, *b_slot, *b_sseed, *b_scnt;
  double4 *b_vec;                // unit vector, length
  double2 *b_car, *b_cbo, *b_cnc;  // (fc, dfc/dr) of the three cutoff families
  double2 *nn;                   // nn(C), nn(H)
  // screening neighbours, entry (size_t)i * nss + b_sseed + n
  int *s_ent;  // list entry of the screening neighbour k, relative to seed[i]
  double *s_arik, *s_arjk, *s_boik, *s_bojk, *s_ncik, *s_ncjk, *s_facbo, *s_facnc;
  int *flag;  // bit 0: bond table exhausted, bit 1: screening table exhausted
};

// host: cutoff families of Rebo2Scr (rebo2_db.f90:92-112, :170-253); only C-C is screened
inline void rbs_fill_cut(RbsCut &S, const Rebo2Dev &D, const atx_rebo2_screening *scr) {
  for (int t = 0; t < 7; t++) {
    S.ar_l[t] = S.bo_l[t] = S.nc_l[t] = D.cut_l[t];
    S.ar_h[t] = S.bo_h[t] = S.nc_h[t] = D.cut_h[t];
  }
  S.ar_l[RB_CC] = scr->cc_ar_r1; S.ar_h[RB_CC] = scr->cc_ar_r2;
  S.bo_l[RB_CC] = scr->cc_bo_r1; S.bo_h[RB_CC] = scr->cc_bo_r2;
  S.nc_l[RB_CC] = scr->cc_nc_r1; S.nc_h[RB_CC] = scr->cc_nc_r2;
  for (int t = 0; t < 7; t++) {
    double m = D.cut_h[t];
    if (S.ar_h[t] > m) m = S.ar_h[t];
    if (S.bo_h[t] > m) m = S.bo_h[t];
    if (S.nc_h[t] > m) m = S.nc_h[t];
    S.max_cut_sq[t] = m * m;
  }
  S.Cmin = scr->Cmin;
  S.Cmax = scr->Cmax;
  S.dC = scr->Cmax - scr->Cmin;
  S.C_dr_cut = scr->Cmax * scr->Cmax / (4 * (scr->Cmax - 1));
}

// trig_off cutoff between l and h (cutoff.f90:152-196); fCin/fCar/fCbo/fCnc of rebo2_func.f90:63-165
__device__ __forceinline__ void rbs_trig(double l, double h, double dr, double &val, double &dval) {
  if (dr > h) { val = 0.0; dval = 0.0; }
  else if (dr < l) { val = 1.0; dval = 0.0; }
  else if (dr <= l) { val = 1.0; dval = 0.0; }
  else if (dr >= h) { val = 0.0; dval = 0.0; }
  else {
    const double fac = RB_PI / (h - l);
    double sn, cs;
    sincos(fac * (dr - l), &sn, &cs);
    val = 0.5 * (1.0 + cs);
    dval = -0.5 * fac * sn;
  }
}

__device__ __forceinline__ void rbs_outer(double *w, double s, double ux, double uy, double uz, double vx,
                                          double vy, double vz) {
  w[0] += s * ux * vx; w[1] += s * uy * vx; w[2] += s * uz * vx;
  w[3] += s * ux * vy; w[4] += s * uy * vy; w[5] += s * uz * vy;
  w[6] += s * ux * vz; w[7] += s * uy * vz; w[8] += s * uz * vz;
}

__device__ __forceinline__ void rbs_add3(double *f, int at, double x, double y, double z) {
  RBS_ADD(&f[3 * (size_t)at], x);
  RBS_ADD(&f[3 * (size_t)at + 1], y);
  RBS_ADD(&f[3 * (size_t)at + 2], z);
}

// vector from sorted atom s (position pi) to the atom of list entry en
__device__ __forceinline__ void rbs_bond_vector(const Mat3 &A, const double4 &pi, const double4 *pos4, int2 en,
                                                double &dx, double &dy, double &dz) {
  const double4 pj = pos4[en.x];
  dx = pj.x - pi.x; dy = pj.y - pi.y; dz = pj.z - pi.z;
  if (ATX_NONZERO_SHIFT(en.y)) {
    int sx, sy, sz;
    atx_unpack_shift(en.y, sx, sy, sz);
    double ax, ay, az;
    atx_image_vector(A, sx, sy, sz, ax, ay, az);
    dx -= ax; dy -= ay; dz -= az;
  }
}

// ---- loop 1 ---------------------------------------------------------------------------------

__device__ __forceinline__ void rbs_bonds_atom(const RbsTab &T, const Mat3 &A, const Rebo2Dev &P, const RbsCut &S,
                                               const double4 *__restrict__ pos4,
                                               const long long *__restrict__ seed,
                                               const int2 *__restrict__ list, int s) {
  const double screening_threshold = -13.815510557964274;  // log(1d-6), rebo2_type.f90:68
  const double dot_threshold = 1e-10;                      // rebo2_type.f90:69
  const double4 pi = pos4[s];
  const int ti = P.el2typ[(int)pi.w];
  int nb = 0, ns = 0;
  double nC = 0.0, nH = 0.0;
  const size_t qs = (size_t)s * T.nss;
  if (ti > 0) {
    const long long b0 = seed[s], b1 = seed[s + 1];
    for (long long a = b0; a < b1; a++) {
      const int2 en = list[a];
      const int tj = P.el2typ[(int)pos4[en.x].w];
      if (tj <= 0) continue;
      double dx, dy, dz;
      rbs_bond_vector(A, pi, pos4, en, dx, dy, dz);
      const double r2 = dx * dx + dy * dy + dz * dz;
      const int ijpot = rb_Z2pair(ti, tj);
      const double l = P.cut_l[ijpot];
      double rl;
      double2 car, cbo, cnc;
      int nsb = 0;  // screening neighbours of this bond, parked at qs + ns + [0, nsb)
      if (r2 < l * l) {
        car = cbo = cnc = make_double2(1.0, 0.0);
        rl = sqrt(r2);
      } else {
        // :813-1101
        if (!(r2 < S.max_cut_sq[ijpot])) continue;
        bool screened = false, need_derivative = false;
        double sij = 0.0, dsijdrij = 0.0;
        if (ijpot == RB_CC) {
          for (long long kn = b0; kn < b1 && !(screened || sij < screening_threshold); kn++) {
            if (kn == a) continue;  // k == j with the same cell shift is this very list entry
            const int2 ek = list[kn];
            double kx, ky, kz;
            rbs_bond_vector(A, pi, pos4, ek, kx, ky, kz);
            const double rik2 = kx * kx + ky * ky + kz * kz;
            if (!(rik2 < S.C_dr_cut * r2)) continue;
            const double dot_ij_ik = dx * kx + dy * ky + dz * kz;
            const double jx = -dx + kx, jy = -dy + ky, jz = -dz + kz;
            const double dot_ij_jk = dx * jx + dy * jy + dz * jz;
            const double rljk = jx * jx + jy * jy + jz * jz;
            if (dot_ij_ik > dot_threshold && dot_ij_jk < -dot_threshold) {
              const double xik = rik2 / r2, xjk = rljk / r2;
              const double xm = xik - xjk, xp = xik + xjk;
              double fac = 1.0 / (1 - xm * xm);
              const double C = (2 * xp - xm * xm - 1) * fac;
              if (C <= S.Cmin) {
                screened = true;
              } else if (C < S.Cmax) {
                need_derivative = true;
                const double Cmax_C = S.Cmax - C, C_Cmin = C - S.Cmin;
                const double q = Cmax_C / C_Cmin;
                sij = sij - q * q;
                const double dCdrik = 4 * xik * fac * (1 + (C - 1) * xm);
                const double dCdrjk = 4 * xjk * fac * (1 - (C - 1) * xm);
                const double dCdrij = -(dCdrik + dCdrjk);
                fac = 2 * Cmax_C * S.dC / (C_Cmin * C_Cmin * C_Cmin);
                dsijdrij = dsijdrij + fac * dCdrij;
                if (ns + nsb >= T.nss) {
                  RBS_OR(T.flag, 2);
                  screened = true;  // leave the loop; the call fails with the flag set
                  break;
                }
                const size_t q2 = qs + ns + nsb;
                T.s_ent[q2] = (int)(kn - b0);
                T.s_arik[q2] = fac * dCdrik / rik2;
                T.s_arjk[q2] = fac * dCdrjk / rljk;
                nsb++;
              }
            }
          }
        }
        if ((screened || sij < screening_threshold) && r2 > P.cut_h2[ijpot]) continue;  // fully screened
        rl = sqrt(r2);
        double fcin, dfcin, fa, dfa, fb, dfb, fn, dfn;
        if (screened) {
          rbs_trig(P.cut_l[ijpot], P.cut_h[ijpot], rl, fcin, dfcin);
          car = cbo = cnc = make_double2(fcin, dfcin);
          nsb = 0;
        } else if (need_derivative) {
          sij = exp(sij);
          rbs_trig(P.cut_l[ijpot], P.cut_h[ijpot], rl, fcin, dfcin);
          rbs_trig(S.ar_l[ijpot], S.ar_h[ijpot], rl, fa, dfa);
          rbs_trig(S.bo_l[ijpot], S.bo_h[ijpot], rl, fb, dfb);
          rbs_trig(S.nc_l[ijpot], S.nc_h[ijpot], rl, fn, dfn);
          car = make_double2((1.0 - fcin) * sij * fa + fcin,
                             (1.0 - fcin) * sij * (dfa + fa * dsijdrij / rl) - dfcin * sij * fa + dfcin);
          cbo = make_double2((1.0 - fcin) * sij * fb + fcin,
                             (1.0 - fcin) * sij * (dfb + fb * dsijdrij / rl) - dfcin * sij * fb + dfcin);
          cnc = make_double2((1.0 - fcin) * sij * fn + fcin,
                             (1.0 - fcin) * sij * (dfn + fn * dsijdrij / rl) - dfcin * sij * fn + dfcin);
          for (int n = 0; n < nsb; n++) {
            const size_t q2 = qs + ns + n;
            const double aik = T.s_arik[q2], ajk = T.s_arjk[q2];
            T.s_boik[q2] = aik * sij * fb * (1.0 - fcin);
            T.s_bojk[q2] = ajk * sij * fb * (1.0 - fcin);
            T.s_ncik[q2] = aik * sij * fn * (1.0 - fcin);
            T.s_ncjk[q2] = ajk * sij * fn * (1.0 - fcin);
            T.s_arik[q2] = aik * sij * fa * (1.0 - fcin);
            T.s_arjk[q2] = ajk * sij * fa * (1.0 - fcin);
            T.s_facbo[q2] = 0.0;
            T.s_facnc[q2] = 0.0;
          }
        } else {
          rbs_trig(S.ar_l[ijpot], S.ar_h[ijpot], rl, fa, dfa);
          rbs_trig(S.bo_l[ijpot], S.bo_h[ijpot], rl, fb, dfb);
          rbs_trig(S.nc_l[ijpot], S.nc_h[ijpot], rl, fn, dfn);
          if (rl < P.cut_h[ijpot]) {
            rbs_trig(P.cut_l[ijpot], P.cut_h[ijpot], rl, fcin, dfcin);
            car = make_double2((1.0 - fcin) * fa + fcin, (1.0 - fcin) * dfa - dfcin * fa + dfcin);
            cbo = make_double2((1.0 - fcin) * fb + fcin, (1.0 - fcin) * dfb - dfcin * fb + dfcin);
            cnc = make_double2((1.0 - fcin) * fn + fcin, (1.0 - fcin) * dfn - dfcin * fn + dfcin);
          } else {
            car = make_double2(fa, dfa);
            cbo = make_double2(fb, dfb);
            cnc = make_double2(fn, dfn);
          }
          nsb = 0;
        }
      }
      if (nb >= T.nbs || nb >= RBS_NBL) { RBS_OR(T.flag, 1); break; }
      const size_t q = (size_t)s * T.nbs + nb;
      T.b_nb[q] = en.x;
      T.b_typ[q] = ijpot;
      T.b_shift[q] = en.y;
      T.b_slot[q] = (int)(a - b0);
      T.b_vec[q] = make_double4(dx / rl, dy / rl, dz / rl, rl);
      T.b_car[q] = car;
      T.b_cbo[q] = cbo;
      T.b_cnc[q] = cnc;
      T.b_sseed[q] = ns;
      T.b_scnt[q] = nsb;
      ns += nsb;
      if (tj == RB_C) nC += cnc.x; else nH += cnc.x;
      nb++;
    }
  }
  T.b_cnt[s] = nb;
  T.nn[s] = make_double2(nC, nH);
}

// ---- loop 2 ---------------------------------------------------------------------------------

// add v to the screening force factor sfacnc of every screening neighbour of bond q (table index)
__device__ __forceinline__ void rbs_add_facnc(const RbsTab &T, int at, size_t q, double v) {
  const int n = T.b_scnt[q];
  if (n <= 0) return;
  const size_t q2 = (size_t)at * T.nss + T.b_sseed[q];
  for (int m = 0; m < n; m++) RBS_ADD(&T.s_facnc[q2 + m], v);
}

__device__ __forceinline__ void rbs_force_atom(const RbsTab &T, const Mat3 &A, const Rebo2Dev &P, const RbsCut &S,
                                               const double4 *__restrict__ pos4,
                                               const long long *__restrict__ seed,
                                               const int2 *__restrict__ list, const int *__restrict__ order,
                                               double *__restrict__ f, double *__restrict__ epa,
                                               double *__restrict__ wpa, double *__restrict__ epb,
                                               double *__restrict__ fpb, double *__restrict__ wpb, int i,
                                               double *acc /* RBS_NSUM */) {
  const int nbs = T.nbs;
  const double4 pi = pos4[i];
  const int ktypi = P.el2typ[(int)pi.w];
  const int nbi = ktypi > 0 ? T.b_cnt[i] : 0;
  if (nbi <= 0) return;
  const size_t qi = (size_t)i * nbs;
  double fix = 0.0, fiy = 0.0, fiz = 0.0;
  // ---- ik_loop1 (:1231-1317): conjugation inputs of the neighbours of i (neighbour-count cutoff)
  double fxik[RBS_NBL], dncx[RBS_NBL];  // fconj(x_ik), fcik * dfconj/dx
  double nconjit = 0.0;
  for (int ik = 0; ik < nbi; ik++) {
    const int k = T.b_nb[qi + ik];
    const int tk = P.el2typ[(int)pos4[k].w];
    fxik[ik] = 0.0;
    dncx[ik] = 0.0;
    if (tk == RB_C) {
      const double2 ck = T.b_cnc[qi + ik];
      const double2 nk = T.nn[k];
      double xik = nk.x + nk.y - ck.x, dfx;
      rb_fconj(xik, fxik[ik], dfx);
      dncx[ik] = ck.x * dfx;
      nconjit += ck.x * fxik[ik];
    }
  }
  const double2 nni = T.nn[i];

  for (int ij = 0; ij < nbi; ij++) {
    const int j = T.b_nb[qi + ij];
    int jsx, jsy, jsz;
    atx_unpack_shift(T.b_shift[qi + ij], jsx, jsy, jsz);
    // j_gt_i (:1332); the index comparison is made in ORIGINAL atom numbering so that per-bond
    // outputs land in the same list slot as in the reference
    const bool zero = (jsx == 0 && jsy == 0 && jsz == 0);
    const bool pos = jsx != 0 ? jsx > 0 : (jsy != 0 ? jsy > 0 : jsz > 0);
    if (!((zero && order[j] > order[i]) || pos)) continue;
    const int ijpot = T.b_typ[qi + ij];
    const double4 vij = T.b_vec[qi + ij];
    const double rlij = vij.w;
    if (!(rlij < S.ar_h[ijpot])) continue;
    const int ktypj = P.el2typ[(int)pos4[j].w];
    const double rlijr = 1.0 / rlij;
    const double nx = vij.x, ny = vij.y, nz = vij.z;
    const double rijx = rlij * nx, rijy = rlij * ny, rijz = rlij * nz;
    const double2 carij = T.b_car[qi + ij];
    const double fcarij = carij.x, dfcarijr = carij.y;
    const double fcncij = T.b_cnc[qi + ij].x;
    const double2 nnj = T.nn[j];
    double niC = nni.x, niH = nni.y, njC = nnj.x, njH = nnj.y;
    if (ktypj == RB_C) niC -= fcncij; else niH -= fcncij;
    if (ktypi == RB_C) njC -= fcncij; else njH -= fcncij;
    if (niC > 4.0) niC = 4.0;
    if (niH > 4.0) niH = 4.0;
    double nti = niC + niH;
    if (njC > 4.0) njC = 4.0;
    if (njH > 4.0) njH = 4.0;
    double ntj = njC + njH;
    double faij, dfaijr, frij, dfrijr;
    rb_VA(P, ijpot, rlij, faij, dfaijr);
    rb_VR(P, ijpot, rlij, frij, dfrijr);
    // virial of the bond: w = -sum_a (r_a - r_i) (x) F_a; the bond-order parts are added where the final force
    // on each neighbour is known (as in atx_rebo2_atom.cuh)
    double wij[9];
    for (int q = 0; q < 9; q++) wij[q] = 0.0;
    double fjx = 0.0, fjy = 0.0, fjz = 0.0;
    double zij = 0.0, dix = 0, diy = 0, diz = 0, djx = 0, djy = 0, djz = 0, dzdni = 0.0;
    double nconji = 0.0;
    double dbk[RBS_NBL][3], zfaci[RBS_NBL];

    // ---- ik_loop2 (:1407-1587): angular terms with the bond-order cutoff
    for (int ik = 0; ik < nbi; ik++) {
      zfaci[ik] = 0.0;
      dbk[ik][0] = dbk[ik][1] = dbk[ik][2] = 0.0;
      if (ik == ij) {
        nconji = nconjit - T.b_cnc[qi + ik].x * fxik[ik];
        continue;
      }
      const int ikpot = T.b_typ[qi + ik];
      const double4 vik = T.b_vec[qi + ik];
      const double rlik = vik.w;
      if (!(rlik < S.bo_h[ikpot])) continue;
      const double2 cik = T.b_cbo[qi + ik];
      const double kx = vik.x, ky = vik.y, kz = vik.z;
      const double fcik = cik.x, dfcikr = cik.y;
      double qfacan, qfadan, gfacan, gddan, dgdn;
      rb_h(P, ijpot, ikpot, rlij - rlik, qfacan, qfadan);
      const double costh = kx * nx + ky * ny + kz * nz;
      rb_g(P, ktypi, costh, nti, gfacan, gddan, dgdn);
      double ex = kx * rlik - nx * rlij, ey = ky * rlik - ny * rlij, ez = kz * rlik - nz * rlij;
      const double disjk = sqrt(ex * ex + ey * ey + ez * ez);
      ex /= disjk; ey /= disjk; ez /= disjk;
      const double dcsdij = 1.0 / rlik - costh * rlijr;
      const double dcsdik = rlijr - costh / rlik;
      const double dcsdjk = -disjk * rlijr / rlik;
      dzdni += fcik * dgdn * qfacan;
      const double dzfac = fcik * gddan * qfacan;
      zfaci[ik] = gfacan * qfacan;
      zij += fcik * gfacan * qfacan;
      const double dzdrij = gfacan * fcik * qfadan;
      const double dzdrik = gfacan * (dfcikr * qfacan - fcik * qfadan);
      const double dfx = dzdrij * nx + dzfac * (dcsdij * nx - dcsdjk * ex);
      const double dfy = dzdrij * ny + dzfac * (dcsdij * ny - dcsdjk * ey);
      const double dfz = dzdrij * nz + dzfac * (dcsdij * nz - dcsdjk * ez);
      dix += -dzdrij * nx - dzdrik * kx + dzfac * (-dcsdij * nx - dcsdik * kx);
      diy += -dzdrij * ny - dzdrik * ky + dzfac * (-dcsdij * ny - dcsdik * ky);
      diz += -dzdrij * nz - dzdrik * kz + dzfac * (-dcsdij * nz - dcsdik * kz);
      djx += dfx; djy += dfy; djz += dfz;
      const double kx_ = dzdrik * kx + dzfac * (dcsdik * kx + dcsdjk * ex);
      const double ky_ = dzdrik * ky + dzfac * (dcsdik * ky + dcsdjk * ey);
      const double kz_ = dzdrik * kz + dzfac * (dcsdik * kz + dcsdjk * ez);
      dbk[ik][0] = kx_; dbk[ik][1] = ky_; dbk[ik][2] = kz_;
    }

    double pij = 0.0, dpdnci = 0.0, dpdnhi = 0.0;
    if (ktypi == RB_C) {
      rb_table2d(ijpot == RB_CC ? P.Pcc : P.Pch, 5, 5, niH, niC, pij, dpdnhi, dpdnci);
      zij += pij;
      dpdnci += dzdni;
      dpdnhi += dzdni;
    }
    double bij, dfbij;
    rb_bo(P, ktypi, zij, fcarij, faij, bij, dfbij);

    // ---- jl_loop (:1644-1887)
    const size_t qj = (size_t)j * nbs;
    const int nbj = T.b_cnt[j];
    double zji = 0.0, bix = 0, biy = 0, biz = 0, bjx = 0, bjy = 0, bjz = 0, dzdnj = 0.0;
    double nconjj = 0.0;
    double dbl[RBS_NBL][3], zfacj[RBS_NBL];
    double fxjl[RBS_NBL], dnlx[RBS_NBL];
    for (int jl = 0; jl < nbj; jl++) {
      fxjl[jl] = 0.0; dnlx[jl] = 0.0; zfacj[jl] = 0.0;
      dbl[jl][0] = dbl[jl][1] = dbl[jl][2] = 0.0;
      const int l = T.b_nb[qj + jl];
      int lsx, lsy, lsz;
      atx_unpack_shift(T.b_shift[qj + jl], lsx, lsy, lsz);
      lsx += jsx; lsy += jsy; lsz += jsz;
      if (l == i && lsx == 0 && lsy == 0 && lsz == 0) continue;  // l_neq_i
      const int ktypl = P.el2typ[(int)pos4[l].w];
      const int jlpot = T.b_typ[qj + jl];
      const double4 vjl = T.b_vec[qj + jl];
      const double rljl = vjl.w;
      const double lx = vjl.x, ly = vjl.y, lz = vjl.z;
      if (ktypl == RB_C) {
        const double fcjl = T.b_cnc[qj + jl].x;
        const double2 nl_ = T.nn[l];
        double xjl = nl_.x + nl_.y - fcjl, dfx;
        rb_fconj(xjl, fxjl[jl], dfx);
        dnlx[jl] = fcjl * dfx;
        nconjj += fcjl * fxjl[jl];
      }
      if (rljl < S.bo_h[jlpot]) {
        const double2 cjl = T.b_cbo[qj + jl];  // the angular part uses the bond-order cutoff (:1755-1756)
        const double fcjl = cjl.x, dfcjlr = cjl.y;
        double qfacan, qfadan, gfacan, gddan, dgdn;
        rb_h(P, ijpot, jlpot, rlij - rljl, qfacan, qfadan);
        const double costh = -(lx * nx + ly * ny + lz * nz);
        rb_g(P, ktypj, costh, ntj, gfacan, gddan, dgdn);
        double ex = lx * rljl + nx * rlij, ey = ly * rljl + ny * rlij, ez = lz * rljl + nz * rlij;
        const double disil = sqrt(ex * ex + ey * ey + ez * ez);
        ex /= disil; ey /= disil; ez /= disil;
        const double dcsdji = 1.0 / rljl - costh * rlijr;
        const double dcsdjl = rlijr - costh / rljl;
        const double dcsdil = -disil * rlijr / rljl;
        dzdnj += fcjl * dgdn * qfacan;
        const double dzfac = fcjl * gddan * qfacan;
        zfacj[jl] = gfacan * qfacan;
        zji += fcjl * gfacan * qfacan;
        const double dzdrji = gfacan * fcjl * qfadan;
        const double dzdrjl = gfacan * (dfcjlr * qfacan - fcjl * qfadan);
        bjx += dzdrji * nx - dzdrjl * lx + dzfac * (dcsdji * nx - dcsdjl * lx);
        bjy += dzdrji * ny - dzdrjl * ly + dzfac * (dcsdji * ny - dcsdjl * ly);
        bjz += dzdrji * nz - dzdrjl * lz + dzfac * (dcsdji * nz - dcsdjl * lz);
        const double dfx = -dzdrji * nx + dzfac * (-dcsdji * nx - dcsdil * ex);
        const double dfy = -dzdrji * ny + dzfac * (-dcsdji * ny - dcsdil * ey);
        const double dfz = -dzdrji * nz + dzfac * (-dcsdji * nz - dcsdil * ez);
        bix += dfx; biy += dfy; biz += dfz;
        const double lx_ = dzdrjl * lx + dzfac * (dcsdjl * lx + dcsdil * ex);
        const double ly_ = dzdrjl * ly + dzfac * (dcsdjl * ly + dcsdil * ey);
        const double lz_ = dzdrjl * lz + dzfac * (dcsdjl * lz + dcsdil * ez);
        dbl[jl][0] = lx_; dbl[jl][1] = ly_; dbl[jl][2] = lz_;
      }
    }

    double pji = 0.0, dpdncj = 0.0, dpdnhj = 0.0;
    if (ktypj == RB_C) {
      rb_table2d(ijpot == RB_CC ? P.Pcc : P.Pch, 5, 5, njH, njC, pji, dpdnhj, dpdncj);
      zji += pji;
      dpdncj += dzdnj;
      dpdnhj += dzdnj;
    }
    double bji, dfbji;
    rb_bo(P, ktypj, zji, fcarij, faij, bji, dfbji);

    double nconj = nconji * nconji + nconjj * nconjj;
    if (nconj > 8.0) nconj = 8.0;
    if (nti > 3.0) nti = 3.0;
    if (ntj > 3.0) ntj = 3.0;

    // ---- dihedral term of the screened build (ALT_DIHEDRAL, :2089-2371): the angle between the planes
    //      (r_ij, r_k1k2) and (r_ij, r_l1l2) over the pairs k1 < k2 of bond partners of i and l1 < l2 of j
    double bdh = 0.0, tij = 0.0, dtdni = 0.0, dtdnj = 0.0, dtdncn = 0.0;
    if (P.with_dihedral && ijpot == RB_CC) {
      rb_table3d(P.Tcc, 4, 4, 9, nti, ntj, nconj, tij, dtdni, dtdnj, dtdncn);
      tij = 2 * tij; dtdni = 2 * dtdni; dtdnj = 2 * dtdnj; dtdncn = 2 * dtdncn;   // :2109-2112
      const double tije = tij * faij * fcarij;
      if (tij != 0) {
        const double rlijsq = rlij * rlij;
        for (int ik1 = 0; ik1 < nbi - 1; ik1++) {
          if (ik1 == ij) continue;
          const double4 v1 = T.b_vec[qi + ik1];
          if (!(v1.w < S.bo_h[T.b_typ[qi + ik1]])) continue;
          const int k1 = T.b_nb[qi + ik1], k1s = T.b_shift[qi + ik1];
          const double2 c1 = T.b_cbo[qi + ik1];
          for (int ik2 = ik1 + 1; ik2 < nbi; ik2++) {
            if (ik2 == ij) continue;
            const double4 v2 = T.b_vec[qi + ik2];
            if (!(v2.w < S.bo_h[T.b_typ[qi + ik2]])) continue;
            const int k2 = T.b_nb[qi + ik2], k2s = T.b_shift[qi + ik2];
            const double2 c2 = T.b_cbo[qi + ik2];
            const double kx = v2.w * v2.x - v1.w * v1.x, ky = v2.w * v2.y - v1.w * v1.y, kz = v2.w * v2.z - v1.w * v1.z;
            const double dot_ij_k = rijx * kx + rijy * ky + rijz * kz;
            const double ksq = kx * kx + ky * ky + kz * kz;
            const double dck = rlijsq * ksq - dot_ij_k * dot_ij_k;
            for (int jl1 = 0; jl1 < nbj - 1; jl1++) {
              const int l1 = T.b_nb[qj + jl1];
              int s1x, s1y, s1z;
              atx_unpack_shift(T.b_shift[qj + jl1], s1x, s1y, s1z);
              s1x += jsx; s1y += jsy; s1z += jsz;
              const int l1s = atx_pack_shift(s1x, s1y, s1z);
              // l1 is neither i nor k1 nor k2 (as atoms incl. their periodic image, :2163-2167)
              if ((l1 == i && s1x == 0 && s1y == 0 && s1z == 0) ||
                  (l1 == k1 && (l1s & ATX_SHIFT_MASK) == (k1s & ATX_SHIFT_MASK)) ||
                  (l1 == k2 && (l1s & ATX_SHIFT_MASK) == (k2s & ATX_SHIFT_MASK))) continue;
              const double4 w1 = T.b_vec[qj + jl1];
              if (!(w1.w < S.bo_h[T.b_typ[qj + jl1]])) continue;
              const double2 d1 = T.b_cbo[qj + jl1];
              for (int jl2 = jl1 + 1; jl2 < nbj; jl2++) {
                const int l2 = T.b_nb[qj + jl2];
                int s2x, s2y, s2z;
                atx_unpack_shift(T.b_shift[qj + jl2], s2x, s2y, s2z);
                s2x += jsx; s2y += jsy; s2z += jsz;
                const int l2s = atx_pack_shift(s2x, s2y, s2z);
                if ((l2 == i && s2x == 0 && s2y == 0 && s2z == 0) ||
                    (l2 == k1 && (l2s & ATX_SHIFT_MASK) == (k1s & ATX_SHIFT_MASK)) ||
                    (l2 == k2 && (l2s & ATX_SHIFT_MASK) == (k2s & ATX_SHIFT_MASK))) continue;
                const double4 w2 = T.b_vec[qj + jl2];
                if (!(w2.w < S.bo_h[T.b_typ[qj + jl2]])) continue;
                const double2 d2 = T.b_cbo[qj + jl2];
                const double lx = w2.w * w2.x - w1.w * w1.x, ly = w2.w * w2.y - w1.w * w1.y,
                             lz = w2.w * w2.z - w1.w * w1.z;
                const double dot_ij_l = rijx * lx + rijy * ly + rijz * lz;
                const double dot_k_l = kx * lx + ky * ly + kz * lz;
                const double lsq = lx * lx + ly * ly + lz * lz;
                const double dcl = rlijsq * lsq - dot_ij_l * dot_ij_l;
                const double abs_dc = sqrt(dck * dcl);
                const double cost = (dot_ij_k * dot_ij_l - rlijsq * dot_k_l) / abs_dc;
                double bdhij = 1 - cost * cost;
                const double fc4 = c1.x * c2.x * d1.x * d2.x;
                bdh += bdhij * fc4;
                bdhij = bdhij * tij * faij * fcarij / 2;
                const double dbd = -2 * cost * tije * fc4 / 2;
                const double ak = dot_ij_l / abs_dc + cost * dot_ij_k / dck;
                const double al = dot_ij_k / abs_dc + cost * dot_ij_l / dcl;
                const double ar = 2 * dot_k_l / abs_dc + cost * (ksq / dck + lsq / dcl);
                double dx_ = dbd * (ak * kx + al * lx - ar * rijx), dy_ = dbd * (ak * ky + al * ly - ar * rijy),
                       dz_ = dbd * (ak * kz + al * lz - ar * rijz);
                fix += dx_; fiy += dy_; fiz += dz_;
                fjx -= dx_; fjy -= dy_; fjz -= dz_;
                rbs_outer(wij, 1.0, rijx, rijy, rijz, dx_, dy_, dz_);
                dx_ = dbd * (-(cost / dck * kx + lx / abs_dc) * rlijsq + ak * rijx);
                dy_ = dbd * (-(cost / dck * ky + ly / abs_dc) * rlijsq + ak * rijy);
                dz_ = dbd * (-(cost / dck * kz + lz / abs_dc) * rlijsq + ak * rijz);
                rbs_add3(f, k1, dx_, dy_, dz_);
                rbs_add3(f, k2, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, kx, ky, kz, dx_, dy_, dz_);
                dx_ = dbd * (-(cost / dcl * lx + kx / abs_dc) * rlijsq + al * rijx);
                dy_ = dbd * (-(cost / dcl * ly + ky / abs_dc) * rlijsq + al * rijy);
                dz_ = dbd * (-(cost / dcl * lz + kz / abs_dc) * rlijsq + al * rijz);
                rbs_add3(f, l1, dx_, dy_, dz_);
                rbs_add3(f, l2, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, lx, ly, lz, dx_, dy_, dz_);
                // derivatives of the four bond-order cutoffs
                double c = bdhij * c1.y * c2.x * d1.x * d2.x;
                dx_ = c * v1.x; dy_ = c * v1.y; dz_ = c * v1.z;
                fix += dx_; fiy += dy_; fiz += dz_;
                rbs_add3(f, k1, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, v1.w * v1.x, v1.w * v1.y, v1.w * v1.z, dx_, dy_, dz_);
                c = bdhij * c2.y * c1.x * d1.x * d2.x;
                dx_ = c * v2.x; dy_ = c * v2.y; dz_ = c * v2.z;
                fix += dx_; fiy += dy_; fiz += dz_;
                rbs_add3(f, k2, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, v2.w * v2.x, v2.w * v2.y, v2.w * v2.z, dx_, dy_, dz_);
                c = bdhij * d1.y * d2.x * c1.x * c2.x;
                dx_ = c * w1.x; dy_ = c * w1.y; dz_ = c * w1.z;
                fjx += dx_; fjy += dy_; fjz += dz_;
                rbs_add3(f, l1, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, w1.w * w1.x, w1.w * w1.y, w1.w * w1.z, dx_, dy_, dz_);
                c = bdhij * d2.y * d1.x * c1.x * c2.x;
                dx_ = c * w2.x; dy_ = c * w2.y; dz_ = c * w2.z;
                fjx += dx_; fjy += dy_; fjz += dz_;
                rbs_add3(f, l2, -dx_, -dy_, -dz_);
                rbs_outer(wij, 1.0, w2.w * w2.x, w2.w * w2.y, w2.w * w2.z, dx_, dy_, dz_);
                // screening neighbours of the four bonds (:2322-2354)
                {
                  const size_t q1 = (size_t)i * T.nss + T.b_sseed[qi + ik1], q2 = (size_t)i * T.nss + T.b_sseed[qi + ik2];
                  const size_t q3 = (size_t)j * T.nss + T.b_sseed[qj + jl1], q4 = (size_t)j * T.nss + T.b_sseed[qj + jl2];
                  const int n1 = T.b_scnt[qi + ik1], n2 = T.b_scnt[qi + ik2], n3 = T.b_scnt[qj + jl1], n4 = T.b_scnt[qj + jl2];
                  for (int m = 0; m < n1; m++) RBS_ADD(&T.s_facbo[q1 + m], bdhij * c2.x * d1.x * d2.x);
                  for (int m = 0; m < n2; m++) RBS_ADD(&T.s_facbo[q2 + m], bdhij * c1.x * d1.x * d2.x);
                  for (int m = 0; m < n3; m++) RBS_ADD(&T.s_facbo[q3 + m], bdhij * d2.x * c1.x * c2.x);
                  for (int m = 0; m < n4; m++) RBS_ADD(&T.s_facbo[q4 + m], bdhij * d1.x * c1.x * c2.x);
                }
              }
            }
          }
        }
      }
    }

    double fij = 0.0, dfdni = 0.0, dfdnj = 0.0, dfdncn = 0.0;
    if (ijpot == RB_CC) rb_table3d(P.Fcc, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
    else if (ijpot == RB_HH) rb_table3d(P.Fhh, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
    else if (ktypi == RB_C) rb_table3d(P.Fch, 4, 4, 9, ntj, nti, nconj, fij, dfdnj, dfdni, dfdncn);
    else if (ktypj == RB_C) rb_table3d(P.Fch, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
    dfdni += dtdni * bdh;     // :2405-2411
    dfdnj += dtdnj * bdh;
    dfdncn += dtdncn * bdh;
    dfdni = 0.5 * fcarij * faij * dfdni;
    dfdnj = 0.5 * fcarij * faij * dfdnj;
    dfdncn = 0.5 * fcarij * faij * dfdncn;
    const double dfdncni = 2 * dfdncn * nconji;
    const double dfdncnj = 2 * dfdncn * nconjj;

    // ---- forces through N_i, N^conj_i on the neighbours k of i and their neighbours m
    //      (:2433-2470, :2590-2662), screening force factors of the bonds i-k and k-m (:2611-2648)
    for (int ik = 0; ik < nbi; ik++) {
      if (ik == ij) continue;
      const int k = T.b_nb[qi + ik];
      const int tk = P.el2typ[(int)pos4[k].w];
      const double4 vik = T.b_vec[qi + ik];
      const double dfcnc = T.b_cnc[qi + ik].y;
      // dnidk(:, ikc, type) = rnik*dfcikr for the type of k, 0 for the other type
      const double sC = (tk == RB_C) ? dfcnc : 0.0, sH = (tk == RB_H) ? dfcnc : 0.0;
      const double dncdk = fxik[ik] * dfcnc;  // dncnidk = nconjdr * rnik (0 unless k is C)
      const double pref = -(dfdni * (sC + sH) + dfdncni * dncdk) - dfbij * (dpdnci * sC + dpdnhi * sH);
      const double dx_ = pref * vik.x, dy_ = pref * vik.y, dz_ = pref * vik.z;
      double fkx = dx_, fky = dy_, fkz = dz_;
      fix -= dx_; fiy -= dy_; fiz -= dz_;
      rbs_outer(wij, -1.0, vik.w * vik.x, vik.w * vik.y, vik.w * vik.z, dx_, dy_, dz_);
      const int nsik = T.b_scnt[qi + ik];
      if (nsik > 0) {
        const double dffac2 = (tk == RB_C ? dpdnci : dpdnhi) * dfbij + dfdni + dfdncni * fxik[ik];
        const size_t q2 = (size_t)i * T.nss + T.b_sseed[qi + ik];
        for (int m = 0; m < nsik; m++) {
          RBS_ADD(&T.s_facbo[q2 + m], zfaci[ik] * dfbij);
          RBS_ADD(&T.s_facnc[q2 + m], dffac2);
        }
      }
      if (tk == RB_C) {
        int ksx, ksy, ksz;
        atx_unpack_shift(T.b_shift[qi + ik], ksx, ksy, ksz);
        const size_t qk = (size_t)k * nbs;
        const int nbk = T.b_cnt[k];
        const double dffac3 = dfdncni * dncx[ik];
        for (int km = 0; km < nbk; km++) {
          const int m = T.b_nb[qk + km];
          int msx, msy, msz;
          atx_unpack_shift(T.b_shift[qk + km], msx, msy, msz);
          if (m == i && msx + ksx == 0 && msy + ksy == 0 && msz + ksz == 0) continue;
          const double4 vkm = T.b_vec[qk + km];
          const double c = -dffac3 * T.b_cnc[qk + km].y;
          const double mx = c * vkm.x, my = c * vkm.y, mz = c * vkm.z;
          rbs_add3(f, m, mx, my, mz);
          fkx -= mx; fky -= my; fkz -= mz;
          rbs_outer(wij, -1.0, vkm.w * vkm.x, vkm.w * vkm.y, vkm.w * vkm.z, mx, my, mz);
          rbs_add_facnc(T, k, qk + km, dffac3);
        }
      }
      fkx += -dfbij * dbk[ik][0]; fky += -dfbij * dbk[ik][1]; fkz += -dfbij * dbk[ik][2];
      rbs_outer(wij, dfbij, vik.w * vik.x, vik.w * vik.y, vik.w * vik.z, dbk[ik][0], dbk[ik][1], dbk[ik][2]);
      rbs_add3(f, k, fkx, fky, fkz);
    }
    // ---- same on the j side (:2472-2517, :2664-2712)
    for (int jl = 0; jl < nbj; jl++) {
      const int l = T.b_nb[qj + jl];
      int lsx, lsy, lsz;
      atx_unpack_shift(T.b_shift[qj + jl], lsx, lsy, lsz);
      lsx += jsx; lsy += jsy; lsz += jsz;
      if (l == i && lsx == 0 && lsy == 0 && lsz == 0) continue;
      const int tl = P.el2typ[(int)pos4[l].w];
      const double4 vjl = T.b_vec[qj + jl];
      const double dfcnc = T.b_cnc[qj + jl].y;
      const double sC = (tl == RB_C) ? dfcnc : 0.0, sH = (tl == RB_H) ? dfcnc : 0.0;
      const double dncdl = fxjl[jl] * dfcnc;
      const double pref = -(dfdnj * (sC + sH) + dfdncnj * dncdl) - dfbji * (dpdncj * sC + dpdnhj * sH);
      const double dx_ = pref * vjl.x, dy_ = pref * vjl.y, dz_ = pref * vjl.z;
      double flx = dx_, fly = dy_, flz = dz_;
      fjx -= dx_; fjy -= dy_; fjz -= dz_;
      rbs_outer(wij, -1.0, vjl.w * vjl.x, vjl.w * vjl.y, vjl.w * vjl.z, dx_, dy_, dz_);
      const int nsjl = T.b_scnt[qj + jl];
      if (nsjl > 0) {
        const double dffac2 = (tl == RB_C ? dpdncj : dpdnhj) * dfbji + dfdnj + dfdncnj * fxjl[jl];
        const size_t q2 = (size_t)j * T.nss + T.b_sseed[qj + jl];
        for (int m = 0; m < nsjl; m++) {
          RBS_ADD(&T.s_facbo[q2 + m], zfacj[jl] * dfbji);
          RBS_ADD(&T.s_facnc[q2 + m], dffac2);
        }
      }
      if (tl == RB_C) {
        const size_t ql = (size_t)l * nbs;
        const int nbl = T.b_cnt[l];
        const double dffac3 = dfdncnj * dnlx[jl];
        for (int ln = 0; ln < nbl; ln++) {
          const int n = T.b_nb[ql + ln];
          int nsx, nsy, nsz;
          atx_unpack_shift(T.b_shift[ql + ln], nsx, nsy, nsz);
          // n /= j .or. ndc /= jdc with ndc = ldc + dcell(ln)
          if (n == j && nsx + lsx == jsx && nsy + lsy == jsy && nsz + lsz == jsz) continue;
          const double4 vln = T.b_vec[ql + ln];
          const double c = -dffac3 * T.b_cnc[ql + ln].y;
          const double mx = c * vln.x, my = c * vln.y, mz = c * vln.z;
          rbs_add3(f, n, mx, my, mz);
          flx -= mx; fly -= my; flz -= mz;
          rbs_outer(wij, -1.0, vln.w * vln.x, vln.w * vln.y, vln.w * vln.z, mx, my, mz);
          rbs_add_facnc(T, l, ql + ln, dffac3);
        }
      }
      flx += -dfbji * dbl[jl][0]; fly += -dfbji * dbl[jl][1]; flz += -dfbji * dbl[jl][2];
      rbs_outer(wij, dfbji, vjl.w * vjl.x, vjl.w * vjl.y, vjl.w * vjl.z, dbl[jl][0], dbl[jl][1], dbl[jl][2]);
      rbs_add3(f, l, flx, fly, flz);
    }

    // ---- pair terms (:2525-2716)
    const double baveij = 0.5 * (bij + bji + fij + tij * bdh);   // :2524-2526
    const double hlfvij = fcarij * (frij + baveij * faij) / 2;
    acc[0] += 2 * hlfvij;
    if (epa) {
      RBS_ADD(&epa[i], hlfvij);
      RBS_ADD(&epa[j], hlfvij);
    }
    const double dffac = dfrijr * fcarij + baveij * dfaijr * fcarij + frij * dfcarijr + baveij * faij * dfcarijr;
    const double dfx = dffac * nx, dfy = dffac * ny, dfz = dffac * nz;
    fix += dfx; fiy += dfy; fiz += dfz;
    fjx -= dfx; fjy -= dfy; fjz -= dfz;
    rbs_outer(wij, 1.0, rijx, rijy, rijz, dfx, dfy, dfz);
    rbs_outer(wij, dfbij, rijx, rijy, rijz, djx, djy, djz);
    rbs_outer(wij, -dfbji, rijx, rijy, rijz, bix, biy, biz);
    fix += -(dfbij * dix + dfbji * bix); fiy += -(dfbij * diy + dfbji * biy); fiz += -(dfbij * diz + dfbji * biz);
    fjx += -(dfbij * djx + dfbji * bjx); fjy += -(dfbij * djy + dfbji * bjy); fjz += -(dfbij * djz + dfbji * bjz);

    // ---- forces on the screening neighbours of bond i-j, attractive/repulsive cutoff (:2717-2757)
    const int nsij = T.b_scnt[qi + ij];
    if (nsij > 0) {
      const double dffs = frij + baveij * faij;
      const size_t q2 = (size_t)i * T.nss + T.b_sseed[qi + ij];
      const long long b0 = seed[i];
      for (int m = 0; m < nsij; m++) {
        const int2 ek = list[b0 + T.s_ent[q2 + m]];
        double kx, ky, kz;
        rbs_bond_vector(A, pi, pos4, ek, kx, ky, kz);
        const double gx = -rijx + kx, gy = -rijy + ky, gz = -rijz + kz;
        const double cik = dffs * T.s_arik[q2 + m], cjk = dffs * T.s_arjk[q2 + m];
        const double ax = cik * kx, ay = cik * ky, az = cik * kz;
        const double bx = cjk * gx, by = cjk * gy, bz = cjk * gz;
        fix += ax; fiy += ay; fiz += az;
        fjx += bx; fjy += by; fjz += bz;
        rbs_add3(f, ek.x, -ax - bx, -ay - by, -az - bz);
        rbs_outer(wij, 1.0, kx, ky, kz, ax, ay, az);
        rbs_outer(wij, 1.0, gx, gy, gz, bx, by, bz);
      }
    }

    rbs_add3(f, j, fjx, fjy, fjz);
    for (int q = 0; q < 9; q++) acc[1 + q] += wij[q];
    const long long a = seed[i] + T.b_slot[qi + ij];
    if (epb) epb[a] = 2 * hlfvij;
    if (fpb) { fpb[3 * a] = dfx; fpb[3 * a + 1] = dfy; fpb[3 * a + 2] = dfz; }
    if (wpb) {
      for (int q = 0; q < 9; q++) wpb[9 * a + q] = wij[q];
    }
    if (wpa) {
      for (int q = 0; q < 9; q++) {
        RBS_ADD(&wpa[9 * (size_t)i + q], 0.5 * wij[q]);
        RBS_ADD(&wpa[9 * (size_t)j + q], 0.5 * wij[q]);
      }
    }
  }
  rbs_add3(f, i, fix, fiy, fiz);
}

// ---- loop 3 ---------------------------------------------------------------------------------

__device__ __forceinline__ void rbs_scr_atom(const RbsTab &T, const Mat3 &A, const Rebo2Dev &P,
                                             const double4 *__restrict__ pos4,
                                             const long long *__restrict__ seed,
                                             const int2 *__restrict__ list, double *__restrict__ f,
                                             double *__restrict__ wpa, double *__restrict__ wpb, int i,
                                             double *acc /* RBS_NSUM */) {
  const double4 pi = pos4[i];
  const int ktypi = P.el2typ[(int)pi.w];
  const int nbi = ktypi > 0 ? T.b_cnt[i] : 0;
  if (nbi <= 0) return;
  const size_t qi = (size_t)i * T.nbs;
  const long long b0 = seed[i];
  double fix = 0.0, fiy = 0.0, fiz = 0.0;
  for (int ij = 0; ij < nbi; ij++) {
    const int nsij = T.b_scnt[qi + ij];
    if (nsij <= 0) continue;  // wij = 0: nothing is added for this bond
    const int j = T.b_nb[qi + ij];
    const double4 vij = T.b_vec[qi + ij];
    const double rijx = vij.w * vij.x, rijy = vij.w * vij.y, rijz = vij.w * vij.z;
    double fjx = 0.0, fjy = 0.0, fjz = 0.0, wij[9];
    for (int q = 0; q < 9; q++) wij[q] = 0.0;
    const size_t q2 = (size_t)i * T.nss + T.b_sseed[qi + ij];
    for (int m = 0; m < nsij; m++) {
      const double sbo = T.s_facbo[q2 + m], snc = T.s_facnc[q2 + m];
      const double cik = sbo * T.s_boik[q2 + m] + snc * T.s_ncik[q2 + m];
      const double cjk = sbo * T.s_bojk[q2 + m] + snc * T.s_ncjk[q2 + m];
      const int2 ek = list[b0 + T.s_ent[q2 + m]];
      double kx, ky, kz;
      rbs_bond_vector(A, pi, pos4, ek, kx, ky, kz);
      const double gx = -rijx + kx, gy = -rijy + ky, gz = -rijz + kz;
      const double ax = cik * kx, ay = cik * ky, az = cik * kz;
      const double bx = cjk * gx, by = cjk * gy, bz = cjk * gz;
      fix += ax; fiy += ay; fiz += az;
      fjx += bx; fjy += by; fjz += bz;
      rbs_add3(f, ek.x, -ax - bx, -ay - by, -az - bz);
      rbs_outer(wij, 1.0, kx, ky, kz, ax, ay, az);
      rbs_outer(wij, 1.0, gx, gy, gz, bx, by, bz);
    }
    rbs_add3(f, j, fjx, fjy, fjz);
    for (int q = 0; q < 9; q++) acc[1 + q] += wij[q];
    if (wpb) {
      const long long a = b0 + T.b_slot[qi + ij];
      for (int q = 0; q < 9; q++) RBS_ADD(&wpb[9 * a + q], wij[q]);
    }
    if (wpa) {
      for (int q = 0; q < 9; q++) {
        RBS_ADD(&wpa[9 * (size_t)i + q], 0.5 * wij[q]);
        RBS_ADD(&wpa[9 * (size_t)j + q], 0.5 * wij[q]);
      }
    }
  }
  rbs_add3(f, i, fix, fiy, fiz);
}
