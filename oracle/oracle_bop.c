/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see atomistica_oracle.h).
 *
 * Generic Tersoff-Brenner bond-order kernel (SCREENING undefined, PYTHON host),
 * restated from
 *   src/potentials/bop/bop_kernel.f90:563-1068 (loop 1), 1075-1529 (loop 2), 1613-1628
 *   src/potentials/bop/tersoff/tersoff_func.f90:32-225
 *   src/potentials/bop/kumagai/kumagai_func.f90:32-225
 *   src/potentials/bop/brenner/brenner_func.f90:27-215, brenner_module.f90:269-288
 *   src/support/cutoff.f90:152-196 (trig_off)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_oracle.h"

#define M3(M, i, j) (M)[(j) * 3 + (i)]
static const double PI = 3.14159265358979323846264338327950288;

static int imin(int a, int b) { return a < b ? a : b; }

/* macros.inc:123 PAIR_INDEX (1-based in, 1-based out) */
static int pair_index(int i, int j, int maxval) {
  return 1 + imin((i - 1) + (j - 1) * maxval, (j - 1) + (i - 1) * maxval) -
         imin((i - 1) * i / 2, (j - 1) * j / 2);
}

/* cutoff.f90:152-196 */
static void trig_off(double r1, double r2, double r, double *val, double *dval) {
  double fac = PI / (r2 - r1);
  if (r <= r1) { *val = 1.0; *dval = 0.0; }
  else if (r >= r2) { *val = 0.0; *dval = 0.0; }
  else {
    double x = fac * (r - r1);
    *val = 0.5 * (1.0 + cos(x));
    *dval = -0.5 * fac * sin(x);
  }
}

/* pair index of the potential, 0-based: PAIR_INDEX (macros.inc:123) for Tersoff/Kumagai/Brenner,
 * PAIR_INDEX_NS (macros.inc:139, juslin_func.f90:300-313) for Juslin */
static int pidx(const orc_bop_params_t *p, int i, int j) {
  if (p->kind == ORC_JUSLIN) return (j + (i - 1) * p->nel) - 1;
  return pair_index(i, j, p->nel) - 1;
}

/* Brenner / Juslin derived constants (brenner_module.f90:269-288, juslin_module.f90:322-343) */
typedef struct {
  double bo_exp[9], bo_fac[9], bo_exp1[9], expR[9], expA[9], c_sq[9], d_sq[9], c_d[9], VR_f[9],
      VA_f[9];
} brenner_derived_t;

static void brenner_derive(const orc_bop_params_t *p, brenner_derived_t *d) {
  int npairs = p->kind == ORC_JUSLIN ? p->nel * p->nel : p->nel * (p->nel + 1) / 2;
  for (int i = 0; i < npairs; i++) {
    d->bo_exp[i] = -0.5 / p->pp[OB_N][i];
    d->bo_fac[i] = 0.5 * d->bo_exp[i] * p->pp[OB_N][i];
    d->bo_exp1[i] = d->bo_exp[i] - 1.0;
    d->expR[i] = p->pp[OB_BETA][i] * sqrt(2 * p->pp[OB_S][i]);
    d->expA[i] = p->pp[OB_BETA][i] * sqrt(2 / p->pp[OB_S][i]);
    d->c_sq[i] = p->pp[OB_C][i] * p->pp[OB_C][i];
    d->d_sq[i] = p->pp[OB_D][i] * p->pp[OB_D][i];
    d->c_d[i] = d->c_sq[i] / d->d_sq[i];
    d->VR_f[i] = p->pp[OB_D0][i] / (p->pp[OB_S][i] - 1);
    d->VA_f[i] = p->pp[OB_S][i] * p->pp[OB_D0][i] / (p->pp[OB_S][i] - 1);
  }
}

static void f_VA(const orc_bop_params_t *p, const brenner_derived_t *bd, int ij, double dr,
                 double *val, double *dval) {
  double e;
  switch (p->kind) {
    case ORC_TERSOFF:
      e = exp(-p->pp[OT_MU][ij] * dr);
      *val = -p->pp[OT_B][ij] * e;
      *dval = p->pp[OT_B][ij] * p->pp[OT_MU][ij] * e;
      break;
    case ORC_KUMAGAI:
      e = exp(-p->pp[OK_LAMBDA2][ij] * dr);
      *val = -p->pp[OK_B][ij] * e;
      *dval = p->pp[OK_B][ij] * p->pp[OK_LAMBDA2][ij] * e;
      break;
    default:
      e = exp(-bd->expA[ij] * (dr - p->pp[OB_R0][ij]));
      *val = -bd->VA_f[ij] * e;
      *dval = bd->VA_f[ij] * bd->expA[ij] * e;
  }
}

static void f_VR(const orc_bop_params_t *p, const brenner_derived_t *bd, int ij, double dr,
                 double *val, double *dval) {
  double e;
  switch (p->kind) {
    case ORC_TERSOFF:
      e = exp(-p->pp[OT_LAMBDA][ij] * dr);
      *val = p->pp[OT_A][ij] * e;
      *dval = -p->pp[OT_A][ij] * p->pp[OT_LAMBDA][ij] * e;
      break;
    case ORC_KUMAGAI:
      e = exp(-p->pp[OK_LAMBDA1][ij] * dr);
      *val = p->pp[OK_A][ij] * e;
      *dval = -p->pp[OK_A][ij] * p->pp[OK_LAMBDA1][ij] * e;
      break;
    default:
      e = exp(-bd->expR[ij] * (dr - p->pp[OB_R0][ij]));
      *val = bd->VR_f[ij] * e;
      *dval = -bd->VR_f[ij] * bd->expR[ij] * e;
  }
}

/* g(cos theta); ktypi/ijpot/ikpot are 0-based here */
static void f_g(const orc_bop_params_t *p, const brenner_derived_t *bd, int ktypi, int ikpot,
                double costh, double *val, double *dval) {
  switch (p->kind) {
    case ORC_TERSOFF: {
      double omega = p->pp[OT_OMEGA][ikpot];
      double h_c = p->ep[OTE_H][ktypi] - costh;
      double c_sq = p->ep[OTE_C][ktypi] * p->ep[OTE_C][ktypi];
      double d_sq = p->ep[OTE_D][ktypi] * p->ep[OTE_D][ktypi];
      double h = d_sq + h_c * h_c;
      *val = omega * (1.0 + c_sq / d_sq - c_sq / h);
      *dval = -2 * omega * c_sq * h_c / (h * h);
      break;
    }
    case ORC_KUMAGAI: {
      double c1 = p->ep[OKE_C1][ktypi], c2 = p->ep[OKE_C2][ktypi], c3 = p->ep[OKE_C3][ktypi];
      double c4 = p->ep[OKE_C4][ktypi], c5 = p->ep[OKE_C5][ktypi], h = p->ep[OKE_H][ktypi];
      double h_cos = h - costh;
      double h_cos_sq = h_cos * h_cos;
      double tmp = h_cos / (c3 + h_cos_sq);
      double go = c2 * tmp;
      double ga1 = c4 * exp(-c5 * h_cos_sq);
      double v = go * (1.0 + ga1);
      *dval = -2 * (1.0 - h_cos * tmp) * v + 2 * c5 * h_cos_sq * go * ga1;
      *val = c1 + h_cos * v;
      break;
    }
    default: {
      double hc = p->pp[OB_H][ikpot] + costh;
      double h = bd->d_sq[ikpot] + hc * hc;
      *val = p->pp[OB_GAMMA][ikpot] * (1 + bd->c_d[ikpot] - bd->c_sq[ikpot] / h);
      *dval = 2 * p->pp[OB_GAMMA][ikpot] * bd->c_sq[ikpot] * hc / (h * h);
    }
  }
}

static void f_bo(const orc_bop_params_t *p, const brenner_derived_t *bd, int ktypi, int ijpot,
                 double zij, double fcij, double faij, double *bij, double *dfbij) {
  switch (p->kind) {
    case ORC_TERSOFF:
      if (zij > 0.0) {
        double n = p->ep[OTE_N][ktypi];
        double e = -0.5 / n;
        double b = pow(p->ep[OTE_BETA][ktypi], n);
        double arg = 1.0 + b * pow(zij, n);
        *bij = p->pp[OT_XI][ijpot] * pow(arg, e);
        *dfbij = -0.25 * fcij * faij * p->pp[OT_XI][ijpot] * b * pow(zij, n - 1.0) *
                 pow(arg, e - 1.0);
      } else { *bij = 1.0; *dfbij = 0.0; }
      break;
    case ORC_KUMAGAI:
      if (zij > 0.0) {
        double eta = p->ep[OKE_ETA][ktypi];
        double delta = -p->ep[OKE_DELTA][ktypi];
        double arg = 1.0 + pow(zij, eta);
        *bij = pow(arg, delta);
        *dfbij = 0.5 * fcij * faij * eta * pow(zij, eta - 1.0) * delta * pow(arg, delta - 1.0);
      } else { *bij = 1.0; *dfbij = 0.0; }
      break;
    default:
      if (p->pp[OB_N][ijpot] == 1.0) {
        double arg = 1.0 + zij;
        *bij = pow(arg, bd->bo_exp[ijpot]);
        *dfbij = bd->bo_fac[ijpot] * fcij * faij * pow(arg, bd->bo_exp1[ijpot]);
      } else if (zij > 0.0) {
        double arg = 1.0 + pow(zij, p->pp[OB_N][ijpot]);
        *bij = pow(arg, bd->bo_exp[ijpot]);
        *dfbij = bd->bo_fac[ijpot] * fcij * faij * pow(zij, p->pp[OB_N][ijpot] - 1.0) *
                 pow(arg, bd->bo_exp1[ijpot]);
      } else { *bij = 1.0; *dfbij = 0.0; }
  }
}

static void f_h(const orc_bop_params_t *p, int ktypj, int ktypi, int ktypk, int ikpot, double dr,
                double *val, double *dval) {
  switch (p->kind) {
    case ORC_JUSLIN: {
      /* juslin_func.f90:255-294; ktyp* are 1-based database element indices */
      int t = (ktypk + p->nel * (ktypj - 1 + p->nel * (ktypi - 1))) - 1;
      double alpha = p->t_alpha[t], omega = p->t_omega[t];
      int m = p->t_m[t];
      if (m == 1) { *val = omega * exp(alpha * dr); *dval = alpha * (*val); }
      else if (m == 3) {
        double arg = alpha * dr;
        *val = omega * exp(arg * arg * arg);
        *dval = 3 * alpha * arg * arg * (*val);
      } else {
        double arg = alpha * dr;
        *val = omega * exp(pow(arg, m));
        *dval = m * pow(arg, m - 1) * alpha * (*val);
      }
      break;
    }
    case ORC_KUMAGAI: {
      double alpha = p->pp[OK_ALPHA][ikpot];
      if (alpha == 0.0) { *val = 1.0; *dval = 0.0; }
      else {
        int beta = p->ip[ikpot];
        if (beta == 1) { *val = exp(alpha * dr); *dval = alpha * (*val); }
        else if (beta == 3) { *val = exp(dr * dr * dr); *dval = 3 * alpha * dr * dr * (*val); } /* sic, kumagai_func.f90:211-213 */
        else { *val = exp(alpha * pow(dr, beta)); *dval = beta * alpha * pow(dr, beta - 1) * (*val); }
      }
      break;
    }
    default: {
      double mu = (p->kind == ORC_TERSOFF) ? p->pp[OT_MUBO][ikpot] : p->pp[OB_MU][ikpot];
      if (mu == 0.0) { *val = 1.0; *dval = 0.0; }
      else {
        int m = p->ip[ikpot];
        if (m == 1) { *val = exp(2 * mu * dr); *dval = 2 * mu * (*val); }
        else if (m == 3) {
          double arg = 2 * mu * dr;
          *val = exp(arg * arg * arg);
          *dval = 2 * mu * m * arg * arg * (*val);
        } else {
          *val = exp(pow(2 * mu * dr, m));
          *dval = 2 * mu * m * pow(2 * mu * dr, m - 1) * (*val);
        }
      }
    }
  }
}

int orc_bop_energy_and_forces(const orc_bop_params_t *par, int nat, int natloc, const double *r,
                              const double *Abox, const int *el, const intptr_t *seed,
                              const intptr_t *last, const int *neighbors, const int *dc,
                              const int *mask, double *epot, double *f_inout, double *wpot_inout,
                              double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                              double *wpot_per_at, double *wpot_per_bond) {
  brenner_derived_t bd;
  memset(&bd, 0, sizeof(bd));
  if (par->kind == ORC_BRENNER || par->kind == ORC_JUSLIN) brenner_derive(par, &bd);

  /* default_compute_func.f90:61-73 */
  long ntot = 0;
  int nebmax = 0;
  for (int i = 0; i < nat; i++) {
    int d = (int)(last[i] - seed[i] + 1);
    if (d > nebmax) nebmax = d;
    ntot += d;
  }
  long nebsize = ntot + nat + 1;

  int *neb = (int *)malloc(sizeof(int) * nebsize);
  long *nbb = (long *)malloc(sizeof(long) * nebsize);
  int *bndtyp = (int *)malloc(sizeof(int) * nebsize);
  double *bndlen = (double *)malloc(sizeof(double) * nebsize);
  double *bndnm = (double *)malloc(sizeof(double) * 3 * nebsize);
  double *cutfcn = (double *)malloc(sizeof(double) * nebsize);
  double *cutdrv = (double *)malloc(sizeof(double) * nebsize);
  long *neb_seed = (long *)malloc(sizeof(long) * (nat + 1));
  long *neb_last = (long *)malloc(sizeof(long) * (nat + 1));
  double *pe = (double *)calloc(nat > 0 ? nat : 1, sizeof(double));
  double *f = (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double));
  double *dbidk = (double *)malloc(sizeof(double) * 3 * (nebmax + 1));
  int *nebofi = (int *)malloc(sizeof(int) * (nebmax + 1));
  long *slotofi = (long *)malloc(sizeof(long) * (nebmax + 1));
  double wpot[9] = {0};

  /* loop 1: bop_kernel.f90:563-1068 */
  long nebtot = 0;
  for (int i = 0; i < natloc; i++) {
    int eli = el[i];
    neb_seed[i] = nebtot;
    neb_last[i] = nebtot - 1;
    if (eli <= 0) continue;
    for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
      int j = neighbors[jn - 1] - 1;
      int elj = el[j];
      if (elj <= 0) continue;
      double rij[3];
      for (int k = 0; k < 3; k++) {
        double s = 0.0;
        for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (jn - 1) + c];
        rij[k] = r[3 * j + k] - r[3 * i + k] - s;
      }
      double rlij = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
      int el2ij = pidx(par, eli, elj);
      double r1 = par->r1[el2ij], r2 = par->r2[el2ij];
      if (rlij < r1 * r1) {
        cutfcn[nebtot] = 1.0;
        cutdrv[nebtot] = 0.0;
      } else if (rlij < r2 * r2) {
        trig_off(r1, r2, sqrt(rlij), &cutfcn[nebtot], &cutdrv[nebtot]);
      } else
        continue;
      rlij = sqrt(rlij);
      neb[nebtot] = j;
      nbb[nebtot] = jn - 1;
      bndlen[nebtot] = rlij;
      for (int k = 0; k < 3; k++) bndnm[3 * nebtot + k] = rij[k] / rlij;
      bndtyp[nebtot] = el2ij;
      neb_last[i] = nebtot;
      nebtot++;
    }
  }

  /* loop 2: bop_kernel.f90:1075-1529 */
  for (int i = 0; i < natloc; i++) {
    int eli = el[i];
    if (eli <= 0) continue;
    double fi[3] = {0, 0, 0};
    long istart = neb_seed[i], ifinsh = neb_last[i];
    for (long ij = istart; ij <= ifinsh; ij++) {
      int j = neb[ij];
      int maskfac = 2;
      if (mask) {
        if (mask[i] == 0 && mask[j] == 0) maskfac = 0;
        else if (mask[i] == 0 || mask[j] == 0) maskfac = 1;
      }
      int el2ij = bndtyp[ij];
      double rlij = bndlen[ij];
      if (!(maskfac > 0 && rlij < par->r2[el2ij])) continue;

      double fj[3] = {0, 0, 0};
      double rlijr = 1.0 / rlij;
      const double *rnij = &bndnm[3 * ij];
      double rij[3] = {rlij * rnij[0], rlij * rnij[1], rlij * rnij[2]};
      double fcarij = cutfcn[ij], dfcarijr = cutdrv[ij];
      double VAij, dVAij, VRij, dVRij;
      f_VA(par, &bd, el2ij, rlij, &VAij, &dVAij);
      f_VR(par, &bd, el2ij, rlij, &VRij, &dVRij);
      VAij = 0.5 * maskfac * VAij;
      dVAij = 0.5 * maskfac * dVAij;
      VRij = 0.5 * maskfac * VRij;
      dVRij = 0.5 * maskfac * dVRij;

      double wij[9] = {0}, wijb[9] = {0};
      double zij = 0.0, dbidi[3] = {0, 0, 0}, dbidj[3] = {0, 0, 0};
      int ikc = 0;
      for (long ik = istart; ik <= ifinsh; ik++) {
        int k = neb[ik];
        nebofi[ikc] = k;
        slotofi[ikc] = ik;
        double fcik = cutfcn[ik];
        if (ik != ij) {
          int ikpot = bndtyp[ik];
          double rlik = bndlen[ik];
          if (rlik < par->r2[ikpot]) {
            const double *rnik = &bndnm[3 * ik];
            double rik[3] = {rlik * rnik[0], rlik * rnik[1], rlik * rnik[2]};
            double dfcikr = cutdrv[ik];
            double h_Dr, dh_dDr, g_costh, dg_dcosth;
            f_h(par, el[j], eli, el[neb[ik]], ikpot, rlij - rlik, &h_Dr, &dh_dDr);
            double costh = rnik[0] * rnij[0] + rnik[1] * rnij[1] + rnik[2] * rnij[2];
            f_g(par, &bd, eli - 1, ikpot, costh, &g_costh, &dg_dcosth);
            double dkc[3];
            for (int c = 0; c < 3; c++) dkc[c] = rnik[c] * rlik - rnij[c] * rlij;
            double disjk = sqrt(dkc[0] * dkc[0] + dkc[1] * dkc[1] + dkc[2] * dkc[2]);
            for (int c = 0; c < 3; c++) dkc[c] = dkc[c] / disjk;
            double dcsdij = 1.0 / rlik - costh * rlijr;
            double dcsdik = rlijr - costh / rlik;
            double dcsdjk = -disjk * rlijr / rlik;
            double dzfac = fcik * dg_dcosth * h_Dr;
            zij = zij + fcik * g_costh * h_Dr;
            double dzdrij = g_costh * fcik * dh_dDr;
            double dzdrik = g_costh * (dfcikr * h_Dr - fcik * dh_dDr);
            double df[3];
            for (int c = 0; c < 3; c++) {
              double dcsdi = -dcsdij * rnij[c] - dcsdik * rnik[c];
              double dcsdj = dcsdij * rnij[c] - dcsdjk * dkc[c];
              double dcsdk = dcsdik * rnik[c] + dcsdjk * dkc[c];
              double dgdi = dzfac * dcsdi, dgdj = dzfac * dcsdj, dgdk = dzfac * dcsdk;
              dbidi[c] = dbidi[c] - dzdrij * rnij[c] - dzdrik * rnik[c] + dgdi;
              df[c] = dzdrij * rnij[c] + dgdj;
              dbidj[c] = dbidj[c] + df[c];
              dbidk[3 * ikc + c] = dzdrik * rnik[c] + dgdk;
            }
            for (int b = 0; b < 3; b++)
              for (int a = 0; a < 3; a++)
                M3(wijb, a, b) = M3(wijb, a, b) - rij[a] * df[b] - rik[a] * dbidk[3 * ikc + b];
          } else {
            dbidk[3 * ikc + 0] = dbidk[3 * ikc + 1] = dbidk[3 * ikc + 2] = 0.0;
          }
        }
        ikc++;
      }
      int numnbi = ikc;

      double bij, dbij_dzij;
      f_bo(par, &bd, eli - 1, el2ij, zij, fcarij, VAij, &bij, &dbij_dzij);

      double dffac = 0.5 * fcarij * (VRij + bij * VAij);
      pe[i] += dffac;
      pe[j] += dffac;
      if (epot_per_bond) epot_per_bond[nbb[ij]] += dffac;

      dffac = 0.5 * (dVRij * fcarij + bij * dVAij * fcarij + VRij * dfcarijr + bij * VAij * dfcarijr);
      double df[3];
      for (int c = 0; c < 3; c++) {
        df[c] = dffac * rnij[c];
        fi[c] += df[c];
        fj[c] -= df[c];
      }
      for (int b = 0; b < 3; b++)
        for (int a = 0; a < 3; a++)
          M3(wij, a, b) = M3(wij, a, b) + rij[a] * df[b] - dbij_dzij * M3(wijb, a, b);
      if (f_per_bond)
        for (int c = 0; c < 3; c++) f_per_bond[3 * nbb[ij] + c] += df[c];
      for (int c = 0; c < 3; c++) {
        fi[c] += -dbij_dzij * dbidi[c];
        fj[c] += -dbij_dzij * dbidj[c];
      }
      for (ikc = 0; ikc < numnbi; ikc++) {
        /* reference: k /= j .or. kdc /= jdc, i.e. a different list slot */
        if (slotofi[ikc] != ij) {
          int k = nebofi[ikc];
          for (int c = 0; c < 3; c++) f[3 * k + c] += -dbij_dzij * dbidk[3 * ikc + c];
        }
      }
      for (int c = 0; c < 9; c++) wpot[c] += wij[c];
      if (wpot_per_bond)
        for (int c = 0; c < 9; c++) wpot_per_bond[9 * nbb[ij] + c] += wij[c];
      if (wpot_per_at)
        for (int c = 0; c < 9; c++) {
          wpot_per_at[9 * i + c] += wij[c] / 2;
          wpot_per_at[9 * j + c] += wij[c] / 2;
        }
      for (int c = 0; c < 3; c++) f[3 * j + c] += fj[c];
    }
    for (int c = 0; c < 3; c++) f[3 * i + c] += fi[c];
  }

  double e = 0.0;
  for (int i = 0; i < nat; i++) e += pe[i];
  *epot += 0.5 * e;
  for (int i = 0; i < nat; i++) {
    if (epot_per_at) epot_per_at[i] += 0.5 * pe[i];
    for (int c = 0; c < 3; c++) f_inout[3 * i + c] += f[3 * i + c];
  }
  for (int c = 0; c < 9; c++) wpot_inout[c] += wpot[c];

  free(neb); free(nbb); free(bndtyp); free(bndlen); free(bndnm); free(cutfcn); free(cutdrv);
  free(neb_seed); free(neb_last); free(pe); free(f); free(dbidk); free(nebofi); free(slotofi);
  return 0;
}

/* ======================================================================================
 * Screened variants (TersoffScr, KumagaiScr, BrennerScr): the same kernel compiled with
 * SCREENING defined and CUTOFF_T = exp_cutoff_t (tersoff_scr.f90:46-47 & co), restated from
 *   src/potentials/bop/bop_kernel.f90:563-1068 (loop 1 incl. 682-995 screening function),
 *                                     1075-1529 (loop 2 incl. 1448-1520 screening forces, ar part),
 *                                     1531-1611 (loop 3: screening forces, bond-order part)
 *   src/potentials/bop/default_bind_to_func.f90:44-104 (Cmin/Cmax/C_dr_cut, cutoffs)
 *   src/support/cutoff.f90:232-293 (exp_cutoff)
 * Build options of the Python host: PARTIAL_SCREENING, SIN_S, BO_WITH_D, SEPARATE_H_ARGUMENTS
 * undefined; screening_threshold = log(1e-6), dot_threshold = 1e-10 (tersoff_type.f90:86-87).
 * ====================================================================================== */

typedef struct { double r1, r2, fac1, fac2, c, d, off; } exp_cutoff_t;

static void exp_cutoff_init(exp_cutoff_t *t, double r1, double r2) {
  t->r1 = r1;
  t->r2 = r2;
  t->fac1 = 1.0 / (r2 - r1);
  double val1 = exp(-8.0);
  double dval1 = -24 * val1;
  double ddval1 = -48 * val1 - 24 * dval1;
  t->c = (-3 * dval1 + ddval1) / 3;
  t->d = (2 * dval1 - ddval1) / 4;
  t->fac2 = 1.0 / (1 - val1 - t->c - t->d);
  t->off = val1 + t->c + t->d;
}

/* JuslinScr keeps the trigonometric form for all three cutoffs (juslin_func.f90:27-122: fCin, fCar,
 * fCbo; constants juslin_module.f90:345-372) */
static void juslin_trig_f(const exp_cutoff_t *t, double r, double *val, double *dval) {
  if (r > t->r2) { *val = 0.0; *dval = 0.0; }
  else if (r < t->r1) { *val = 1.0; *dval = 0.0; }
  else {
    double fca = PI / (t->r2 - t->r1), fc = -0.5 * fca;
    double arg = fca * (r - t->r1);
    *val = 0.5 * (1.0 + cos(arg));
    *dval = fc * sin(arg);
  }
}

static void exp_cutoff_f(const exp_cutoff_t *t, double r, double *val, double *dval) {
  if (r <= t->r1) { *val = 1.0; *dval = 0.0; }
  else if (r >= t->r2) { *val = 0.0; *dval = 0.0; }
  else {
    double x = t->fac1 * (r - t->r1);
    double x2 = x * x;
    double v = exp(-8 * x * x2);
    double dv = -24 * x2 * v;
    *dval = t->fac1 * t->fac2 * (dv + 3 * t->c * x2 + 4 * t->d * x * x2);
    *val = t->fac2 * (v + t->c * x * x2 + t->d * x2 * x2 - t->off);
  }
}

#define GROW(ptr, type, cap, need)                                      \
  do {                                                                  \
    if ((need) > (cap)) {                                               \
      long ncap_ = (cap) * 2 > (need) ? (cap) * 2 : (need);             \
      ptr = (type *)realloc(ptr, sizeof(type) * ncap_);                 \
    }                                                                   \
  } while (0)

int orc_bop_scr_energy_and_forces(const orc_bop_params_t *par, const orc_bop_scr_t *scr, int nat,
                                  int natloc, const double *r, const double *Abox, const int *el,
                                  const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                  const int *dc, const int *mask, double *epot, double *f_inout,
                                  double *wpot_inout, double *epot_per_at, double *epot_per_bond,
                                  double *f_per_bond, double *wpot_per_at, double *wpot_per_bond) {
  brenner_derived_t bd;
  memset(&bd, 0, sizeof(bd));
  if (par->kind == ORC_BRENNER || par->kind == ORC_JUSLIN) brenner_derive(par, &bd);
  const double screening_threshold = log(1e-6), dot_threshold = 1e-10;

  const int juslin = par->kind == ORC_JUSLIN;
  int npairs = juslin ? par->nel * par->nel : par->nel * (par->nel + 1) / 2;
  void (*cutoff_f)(const exp_cutoff_t *, double, double *, double *) = juslin ? juslin_trig_f : exp_cutoff_f;
  exp_cutoff_t cut_in[9], cut_out[9], cut_bo[9];
  double cut_in_l[9], cut_in_h[9], cut_in_h2[9], cut_out_l[9], cut_out_h[9], cut_bo_h[9],
      max_cut_sq[9], Cmin[9], Cmax[9], dC[9], C_dr_cut[9];
  for (int i = 0; i < npairs; i++) {
    Cmin[i] = scr->Cmin[i];
    Cmax[i] = scr->Cmax[i];
    dC[i] = Cmax[i] - Cmin[i];
    /* brenner_module.f90:198-205 & co: 1 unless Cmax > 2; juslin_module.f90:307: unconditional */
    C_dr_cut[i] = (juslin || Cmax[i] > 2.0) ? Cmax[i] * Cmax[i] / (4 * (Cmax[i] - 1)) : 1.0;
    exp_cutoff_init(&cut_in[i], par->r1[i], par->r2[i]);
    cut_in_l[i] = par->r1[i];
    cut_in_h[i] = par->r2[i];
    cut_in_h2[i] = par->r2[i] * par->r2[i];
    exp_cutoff_init(&cut_out[i], scr->or1[i], scr->or2[i]);
    cut_out_l[i] = scr->or1[i];
    cut_out_h[i] = scr->or2[i];
    exp_cutoff_init(&cut_bo[i], scr->bor1[i], scr->bor2[i]);
    cut_bo_h[i] = scr->bor2[i];
    double m = cut_in_h[i];
    if (cut_out_h[i] > m) m = cut_out_h[i];
    if (cut_bo_h[i] > m) m = cut_bo_h[i];
    max_cut_sq[i] = m * m;
  }
#define cut_ar_h cut_out_h /* tersoff_type.f90:71 */

  long ntot = 0;
  int nebmax = 0;
  for (int i = 0; i < nat; i++) {
    int d = (int)(last[i] - seed[i] + 1);
    if (d > nebmax) nebmax = d;
    ntot += d;
  }
  long nebsize = ntot + nat + 1;

  int *neb = (int *)malloc(sizeof(int) * nebsize);
  long *nbb = (long *)malloc(sizeof(long) * nebsize);
  int *bndtyp = (int *)malloc(sizeof(int) * nebsize);
  double *bndlen = (double *)malloc(sizeof(double) * nebsize);
  double *bndnm = (double *)malloc(sizeof(double) * 3 * nebsize);
  double *cutfcnar = (double *)malloc(sizeof(double) * nebsize);
  double *cutdrvar = (double *)malloc(sizeof(double) * nebsize);
  double *cutfcnbo = (double *)malloc(sizeof(double) * nebsize);
  double *cutdrvbo = (double *)malloc(sizeof(double) * nebsize);
  long *sneb_seed = (long *)malloc(sizeof(long) * nebsize);
  long *sneb_last = (long *)malloc(sizeof(long) * nebsize);
  long *neb_seed = (long *)malloc(sizeof(long) * (nat + 1));
  long *neb_last = (long *)malloc(sizeof(long) * (nat + 1));
  double *pe = (double *)calloc(nat > 0 ? nat : 1, sizeof(double));
  double *f = (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double));
  double *dbidk = (double *)malloc(sizeof(double) * 3 * (nebmax + 1));
  double *zfaci = (double *)calloc(nebmax + 1, sizeof(double));
  int *nebofi = (int *)malloc(sizeof(int) * (nebmax + 1));
  long *slotofi = (long *)malloc(sizeof(long) * (nebmax + 1));
  long *seedi = (long *)malloc(sizeof(long) * (nebmax + 1));
  long *lasti = (long *)malloc(sizeof(long) * (nebmax + 1));
  double wpot[9] = {0};

  /* screening-neighbour arrays grow on demand (the reference sizes them nebsize*nebmax-ish and
   * raises "snebsize too small"; only the contents matter for parity) */
  long scap = nebsize + 16;
  int *sneb = (int *)malloc(sizeof(int) * scap);
  long *sbnd = (long *)malloc(sizeof(long) * scap);
  double *cutdrarik = (double *)malloc(sizeof(double) * scap);
  double *cutdrarjk = (double *)malloc(sizeof(double) * scap);
  double *cutdrboik = (double *)malloc(sizeof(double) * scap);
  double *cutdrbojk = (double *)malloc(sizeof(double) * scap);

  /* ---- loop 1 ---- */
  long nebtot = 0, snebtot = 0;
  for (int i = 0; i < natloc; i++) {
    int eli = el[i];
    neb_seed[i] = nebtot;
    neb_last[i] = nebtot - 1;
    if (eli <= 0) continue;
    intptr_t jbeg = seed[i], jend = last[i];
    for (intptr_t jn = jbeg; jn <= jend; jn++) {
      int j = neighbors[jn - 1] - 1;
      int elj = el[j];
      if (elj <= 0) continue;
      double rij[3];
      for (int k = 0; k < 3; k++) {
        double s = 0.0;
        for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (jn - 1) + c];
        rij[k] = r[3 * j + k] - r[3 * i + k] - s;
      }
      double rlij = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
      int el2ij = pidx(par, eli, elj);

      if (rlij < cut_in_l[el2ij] * cut_in_l[el2ij]) {
        /* region (a): bop_kernel.f90:634-680 */
        cutfcnar[nebtot] = 1.0; cutdrvar[nebtot] = 0.0;
        cutfcnbo[nebtot] = 1.0; cutdrvbo[nebtot] = 0.0;
        neb[nebtot] = j; nbb[nebtot] = jn - 1;
        sneb_seed[nebtot] = snebtot; sneb_last[nebtot] = snebtot - 1;
        rlij = sqrt(rlij);
        bndlen[nebtot] = rlij;
        for (int k = 0; k < 3; k++) bndnm[3 * nebtot + k] = rij[k] / rlij;
        bndtyp[nebtot] = el2ij;
        neb_last[i] = nebtot;
        nebtot++;
      } else if (rlij < max_cut_sq[el2ij] && cut_out_l[el2ij] < cut_out_h[el2ij]) {
        /* bop_kernel.f90:689-995 */
        int screened = 0, need_derivative = 0;
        double sij = 0.0;
        sneb_seed[nebtot] = snebtot;
        sneb_last[nebtot] = snebtot - 1;
        long ineb = snebtot;
        double dsijdrij = 0.0;
        intptr_t kn = jbeg;
        while (!(screened || sij < screening_threshold) && kn <= jend) {
          int k = neighbors[kn - 1] - 1;
          double rik[3];
          for (int a = 0; a < 3; a++) {
            double s = 0.0;
            for (int c = 0; c < 3; c++) s += M3(Abox, a, c) * (double)dc[3 * (kn - 1) + c];
            rik[a] = r[3 * k + a] - r[3 * i + a] - s;
          }
          double rik2 = rik[0] * rik[0] + rik[1] * rik[1] + rik[2] * rik[2];
          if (rik2 < C_dr_cut[el2ij] * rlij) {
            int same = (k == j) && dc[3 * (kn - 1)] == dc[3 * (jn - 1)] &&
                       dc[3 * (kn - 1) + 1] == dc[3 * (jn - 1) + 1] &&
                       dc[3 * (kn - 1) + 2] == dc[3 * (jn - 1) + 2];
            if (!same) {
              double dot_ij_ik = rij[0] * rik[0] + rij[1] * rik[1] + rij[2] * rik[2];
              double rlik = rik2;
              double rjk[3] = {-rij[0] + rik[0], -rij[1] + rik[1], -rij[2] + rik[2]};
              double dot_ij_jk = rij[0] * rjk[0] + rij[1] * rjk[1] + rij[2] * rjk[2];
              double rljk = rjk[0] * rjk[0] + rjk[1] * rjk[1] + rjk[2] * rjk[2];
              if (dot_ij_ik > dot_threshold && dot_ij_jk < -dot_threshold) {
                double xik = rlik / rlij, xjk = rljk / rlij;
                double xik_m_xjk = xik - xjk, xik_p_xjk = xik + xjk;
                double fac = 1.0 / (1 - xik_m_xjk * xik_m_xjk);
                double C = (2 * xik_p_xjk - xik_m_xjk * xik_m_xjk - 1) * fac;
                if (C <= Cmin[el2ij]) {
                  screened = 1;
                } else if (C < Cmax[el2ij]) {
                  need_derivative = 1;
                  double Cmax_C = Cmax[el2ij] - C, C_Cmin = C - Cmin[el2ij];
                  double q = Cmax_C / C_Cmin;
                  sij = sij - q * q;
                  double dCdrik = 4 * xik * fac * (1 + (C - 1) * xik_m_xjk);
                  double dCdrjk = 4 * xjk * fac * (1 - (C - 1) * xik_m_xjk);
                  double dCdrij = -(dCdrik + dCdrjk);
                  fac = 2 * Cmax_C * dC[el2ij] / (C_Cmin * C_Cmin * C_Cmin);
                  dsijdrij = dsijdrij + fac * dCdrij;
                  double dsijdrik = fac * dCdrik, dsijdrjk = fac * dCdrjk;
                  if (snebtot + 1 > scap) {
                    long need = snebtot + 1;
                    GROW(sneb, int, scap, need); GROW(sbnd, long, scap, need);
                    GROW(cutdrarik, double, scap, need); GROW(cutdrarjk, double, scap, need);
                    GROW(cutdrboik, double, scap, need); GROW(cutdrbojk, double, scap, need);
                    scap = scap * 2 > need ? scap * 2 : need;
                  }
                  sneb[snebtot] = k;
                  sbnd[snebtot] = kn - 1;
                  cutdrarik[snebtot] = dsijdrik / rlik;
                  cutdrarjk[snebtot] = dsijdrjk / rljk;
                  sneb_last[nebtot] = snebtot;
                  snebtot++;
                }
              }
            }
          }
          kn++;
        }

        if ((screened || sij < screening_threshold) && rlij > cut_in_h2[el2ij]) {
          /* fully screened: drop the bond and its screening neighbours */
          snebtot = ineb;
          sneb_last[nebtot] = ineb - 1;
        } else {
          neb[nebtot] = j; nbb[nebtot] = jn - 1;
          rlij = sqrt(rlij);
          bndlen[nebtot] = rlij;
          for (int k = 0; k < 3; k++) bndnm[3 * nebtot + k] = rij[k] / rlij;
          bndtyp[nebtot] = el2ij;
          double fcinij, dfcinijr, fcarij, dfcarijr, fcboij, dfcboijr;
          if (screened) {
            cutoff_f(&cut_in[el2ij], rlij, &fcinij, &dfcinijr);
            cutfcnar[nebtot] = fcinij; cutdrvar[nebtot] = dfcinijr;
            cutfcnbo[nebtot] = fcinij; cutdrvbo[nebtot] = dfcinijr;
            snebtot = ineb;
            sneb_last[nebtot] = ineb - 1;
          } else if (need_derivative) {
            sij = exp(sij);
            cutoff_f(&cut_in[el2ij], rlij, &fcinij, &dfcinijr);
            cutoff_f(&cut_out[el2ij], rlij, &fcarij, &dfcarijr);
            cutoff_f(&cut_bo[el2ij], rlij, &fcboij, &dfcboijr);
            cutfcnar[nebtot] = (1.0 - fcinij) * sij * fcarij + fcinij;
            cutdrvar[nebtot] = (1.0 - fcinij) * sij * (dfcarijr + fcarij * dsijdrij / rlij) -
                               dfcinijr * sij * fcarij + dfcinijr;
            cutfcnbo[nebtot] = (1.0 - fcinij) * sij * fcboij + fcinij;
            cutdrvbo[nebtot] = (1.0 - fcinij) * sij * (dfcboijr + fcboij * dsijdrij / rlij) -
                               dfcinijr * sij * fcboij + dfcinijr;
            for (long q = ineb; q < snebtot; q++) {
              cutdrboik[q] = cutdrarik[q] * sij * fcboij * (1.0 - fcinij);
              cutdrbojk[q] = cutdrarjk[q] * sij * fcboij * (1.0 - fcinij);
              cutdrarik[q] = cutdrarik[q] * sij * fcarij * (1.0 - fcinij);
              cutdrarjk[q] = cutdrarjk[q] * sij * fcarij * (1.0 - fcinij);
            }
          } else {
            cutoff_f(&cut_out[el2ij], rlij, &fcarij, &dfcarijr);
            cutoff_f(&cut_bo[el2ij], rlij, &fcboij, &dfcboijr);
            if (rlij < cut_in_h[el2ij]) {
              cutoff_f(&cut_in[el2ij], rlij, &fcinij, &dfcinijr);
              cutfcnar[nebtot] = (1.0 - fcinij) * fcarij + fcinij;
              cutdrvar[nebtot] = (1.0 - fcinij) * dfcarijr - dfcinijr * fcarij + dfcinijr;
              cutfcnbo[nebtot] = (1.0 - fcinij) * fcboij + fcinij;
              cutdrvbo[nebtot] = (1.0 - fcinij) * dfcboijr - dfcinijr * fcboij + dfcinijr;
            } else {
              cutfcnar[nebtot] = fcarij; cutdrvar[nebtot] = dfcarijr;
              cutfcnbo[nebtot] = fcboij; cutdrvbo[nebtot] = dfcboijr;
            }
          }
          neb_last[i] = nebtot;
          nebtot++;
        }
      } else if (rlij < cut_in_h2[el2ij]) {
        /* pair without an outer cutoff (or1 >= or2): unscreened bond, bop_kernel.f90:997-1055 */
        rlij = sqrt(rlij);
        bndlen[nebtot] = rlij;
        for (int k = 0; k < 3; k++) bndnm[3 * nebtot + k] = rij[k] / rlij;
        bndtyp[nebtot] = el2ij;
        double fcinij, dfcinijr;
        cutoff_f(&cut_in[el2ij], rlij, &fcinij, &dfcinijr);
        cutfcnar[nebtot] = fcinij; cutdrvar[nebtot] = dfcinijr;
        cutfcnbo[nebtot] = fcinij; cutdrvbo[nebtot] = dfcinijr;
        neb[nebtot] = j; nbb[nebtot] = jn - 1;
        sneb_seed[nebtot] = snebtot; sneb_last[nebtot] = snebtot - 1;
        neb_last[i] = nebtot;
        nebtot++;
      }
    }
  }

  double *sfacbo = (double *)calloc(snebtot > 0 ? snebtot : 1, sizeof(double)); /* bop_kernel.f90:434 */

  /* ---- loop 2 ---- */
  for (int i = 0; i < natloc; i++) {
    int eli = el[i];
    if (eli <= 0) continue;
    double fi[3] = {0, 0, 0};
    long istart = neb_seed[i], ifinsh = neb_last[i];
    for (long ij = istart; ij <= ifinsh; ij++) {
      int j = neb[ij];
      int maskfac = 2;
      if (mask) {
        if (mask[i] == 0 && mask[j] == 0) maskfac = 0;
        else if (mask[i] == 0 || mask[j] == 0) maskfac = 1;
      }
      int el2ij = bndtyp[ij];
      double rlij = bndlen[ij];
      if (!(maskfac > 0 && rlij < cut_ar_h[el2ij])) continue;

      double fj[3] = {0, 0, 0};
      double rlijr = 1.0 / rlij;
      const double *rnij = &bndnm[3 * ij];
      double rij[3] = {rlij * rnij[0], rlij * rnij[1], rlij * rnij[2]};
      double fcarij = cutfcnar[ij], dfcarijr = cutdrvar[ij];
      double VAij, dVAij, VRij, dVRij;
      f_VA(par, &bd, el2ij, rlij, &VAij, &dVAij);
      f_VR(par, &bd, el2ij, rlij, &VRij, &dVRij);
      VAij = 0.5 * maskfac * VAij;
      dVAij = 0.5 * maskfac * dVAij;
      VRij = 0.5 * maskfac * VRij;
      dVRij = 0.5 * maskfac * dVRij;

      double wij[9] = {0}, wijb[9] = {0};
      double zij = 0.0, dbidi[3] = {0, 0, 0}, dbidj[3] = {0, 0, 0};
      int ikc = 0;
      for (long ik = istart; ik <= ifinsh; ik++) {
        int k = neb[ik];
        nebofi[ikc] = k;
        slotofi[ikc] = ik;
        seedi[ikc] = sneb_seed[ik];
        lasti[ikc] = sneb_last[ik];
        double fcik = cutfcnbo[ik];
        if (ik != ij) {
          int ikpot = bndtyp[ik];
          double rlik = bndlen[ik];
          if (rlik < cut_bo_h[ikpot]) {
            const double *rnik = &bndnm[3 * ik];
            double rik[3] = {rlik * rnik[0], rlik * rnik[1], rlik * rnik[2]};
            double dfcikr = cutdrvbo[ik];
            double h_Dr, dh_dDr, g_costh, dg_dcosth;
            f_h(par, el[j], eli, el[neb[ik]], ikpot, rlij - rlik, &h_Dr, &dh_dDr);
            double costh = rnik[0] * rnij[0] + rnik[1] * rnij[1] + rnik[2] * rnij[2];
            f_g(par, &bd, eli - 1, ikpot, costh, &g_costh, &dg_dcosth);
            double dkc[3];
            for (int c = 0; c < 3; c++) dkc[c] = rnik[c] * rlik - rnij[c] * rlij;
            double disjk = sqrt(dkc[0] * dkc[0] + dkc[1] * dkc[1] + dkc[2] * dkc[2]);
            for (int c = 0; c < 3; c++) dkc[c] = dkc[c] / disjk;
            double dcsdij = 1.0 / rlik - costh * rlijr;
            double dcsdik = rlijr - costh / rlik;
            double dcsdjk = -disjk * rlijr / rlik;
            double dzfac = fcik * dg_dcosth * h_Dr;
            zfaci[ikc] = g_costh * h_Dr;
            zij = zij + fcik * g_costh * h_Dr;
            double dzdrij = g_costh * fcik * dh_dDr;
            double dzdrik = g_costh * (dfcikr * h_Dr - fcik * dh_dDr);
            double df[3];
            for (int c = 0; c < 3; c++) {
              double dcsdi = -dcsdij * rnij[c] - dcsdik * rnik[c];
              double dcsdj = dcsdij * rnij[c] - dcsdjk * dkc[c];
              double dcsdk = dcsdik * rnik[c] + dcsdjk * dkc[c];
              double dgdi = dzfac * dcsdi, dgdj = dzfac * dcsdj, dgdk = dzfac * dcsdk;
              dbidi[c] = dbidi[c] - dzdrij * rnij[c] - dzdrik * rnik[c] + dgdi;
              df[c] = dzdrij * rnij[c] + dgdj;
              dbidj[c] = dbidj[c] + df[c];
              dbidk[3 * ikc + c] = dzdrik * rnik[c] + dgdk;
            }
            for (int b = 0; b < 3; b++)
              for (int a = 0; a < 3; a++)
                M3(wijb, a, b) = M3(wijb, a, b) - rij[a] * df[b] - rik[a] * dbidk[3 * ikc + b];
          } else {
            zfaci[ikc] = 0.0;
            dbidk[3 * ikc + 0] = dbidk[3 * ikc + 1] = dbidk[3 * ikc + 2] = 0.0;
          }
        }
        ikc++;
      }
      int numnbi = ikc;

      double bij, dbij_dzij;
      f_bo(par, &bd, eli - 1, el2ij, zij, fcarij, VAij, &bij, &dbij_dzij);

      double dffac = 0.5 * fcarij * (VRij + bij * VAij);
      pe[i] += dffac;
      pe[j] += dffac;
      if (epot_per_bond) epot_per_bond[nbb[ij]] += dffac;

      dffac = 0.5 * (dVRij * fcarij + bij * dVAij * fcarij + VRij * dfcarijr + bij * VAij * dfcarijr);
      double df[3];
      for (int c = 0; c < 3; c++) {
        df[c] = dffac * rnij[c];
        fi[c] += df[c];
        fj[c] -= df[c];
      }
      for (int b = 0; b < 3; b++)
        for (int a = 0; a < 3; a++)
          M3(wij, a, b) = M3(wij, a, b) + rij[a] * df[b] - dbij_dzij * M3(wijb, a, b);
      if (f_per_bond)
        for (int c = 0; c < 3; c++) f_per_bond[3 * nbb[ij] + c] += df[c];
      for (int c = 0; c < 3; c++) {
        fi[c] += -dbij_dzij * dbidi[c];
        fj[c] += -dbij_dzij * dbidj[c];
      }
      for (ikc = 0; ikc < numnbi; ikc++) {
        if (slotofi[ikc] != ij) {
          int k = nebofi[ikc];
          for (int c = 0; c < 3; c++) f[3 * k + c] += -dbij_dzij * dbidk[3 * ikc + c];
          /* forces due to screening of the bonds i-k that enter z_ij */
          for (long q = seedi[ikc]; q <= lasti[ikc]; q++) sfacbo[q] = sfacbo[q] + zfaci[ikc] * dbij_dzij;
        }
      }

      /* forces on the screening neighbours of bond i-j, attractive/repulsive part */
      dffac = 0.5 * (VRij + bij * VAij);
      for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
        int k = sneb[q];
        double rik[3], rjk[3];
        for (int a = 0; a < 3; a++) {
          double s = 0.0;
          for (int c = 0; c < 3; c++) s += M3(Abox, a, c) * (double)dc[3 * sbnd[q] + c];
          rik[a] = r[3 * k + a] - r[3 * i + a] - s;
          rjk[a] = -rij[a] + rik[a];
        }
        for (int c = 0; c < 3; c++) {
          df[c] = dffac * cutdrarik[q] * rik[c];
          fi[c] += df[c];
          f[3 * k + c] += -df[c];
        }
        for (int b = 0; b < 3; b++)
          for (int a = 0; a < 3; a++) M3(wij, a, b) = M3(wij, a, b) + rik[a] * df[b];
        for (int c = 0; c < 3; c++) {
          df[c] = dffac * cutdrarjk[q] * rjk[c];
          fj[c] += df[c];
          f[3 * k + c] += -df[c];
        }
        for (int b = 0; b < 3; b++)
          for (int a = 0; a < 3; a++) M3(wij, a, b) = M3(wij, a, b) + rjk[a] * df[b];
      }

      for (int c = 0; c < 9; c++) wpot[c] += wij[c];
      if (wpot_per_bond)
        for (int c = 0; c < 9; c++) wpot_per_bond[9 * nbb[ij] + c] += wij[c];
      if (wpot_per_at)
        for (int c = 0; c < 9; c++) {
          wpot_per_at[9 * i + c] += wij[c] / 2;
          wpot_per_at[9 * j + c] += wij[c] / 2;
        }
      for (int c = 0; c < 3; c++) f[3 * j + c] += fj[c];
    }
    for (int c = 0; c < 3; c++) f[3 * i + c] += fi[c];
  }

  /* ---- loop 3: forces due to screening, bond-order part (bop_kernel.f90:1531-1611) ---- */
  for (int i = 0; i < natloc; i++) {
    if (el[i] <= 0) continue;
    double fi[3] = {0, 0, 0};
    for (long ij = neb_seed[i]; ij <= neb_last[i]; ij++) {
      int j = neb[ij];
      double fj[3] = {0, 0, 0}, wij[9] = {0};
      double rij[3] = {bndlen[ij] * bndnm[3 * ij], bndlen[ij] * bndnm[3 * ij + 1], bndlen[ij] * bndnm[3 * ij + 2]};
      for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
        cutdrboik[q] = sfacbo[q] * cutdrboik[q];
        cutdrbojk[q] = sfacbo[q] * cutdrbojk[q];
      }
      for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
        int k = sneb[q];
        double rik[3], rjk[3], df[3];
        for (int a = 0; a < 3; a++) {
          double s = 0.0;
          for (int c = 0; c < 3; c++) s += M3(Abox, a, c) * (double)dc[3 * sbnd[q] + c];
          rik[a] = r[3 * k + a] - r[3 * i + a] - s;
          rjk[a] = -rij[a] + rik[a];
        }
        for (int c = 0; c < 3; c++) {
          df[c] = cutdrboik[q] * rik[c];
          fi[c] += df[c];
          f[3 * k + c] += -df[c];
        }
        for (int b = 0; b < 3; b++)
          for (int a = 0; a < 3; a++) M3(wij, a, b) = M3(wij, a, b) + rik[a] * df[b];
        for (int c = 0; c < 3; c++) {
          df[c] = cutdrbojk[q] * rjk[c];
          fj[c] += df[c];
          f[3 * k + c] += -df[c];
        }
        for (int b = 0; b < 3; b++)
          for (int a = 0; a < 3; a++) M3(wij, a, b) = M3(wij, a, b) + rjk[a] * df[b];
      }
      for (int c = 0; c < 9; c++) wpot[c] += wij[c];
      if (wpot_per_bond)
        for (int c = 0; c < 9; c++) wpot_per_bond[9 * nbb[ij] + c] += wij[c];
      if (wpot_per_at)
        for (int c = 0; c < 9; c++) {
          wpot_per_at[9 * i + c] += wij[c] / 2;
          wpot_per_at[9 * j + c] += wij[c] / 2;
        }
      for (int c = 0; c < 3; c++) f[3 * j + c] += fj[c];
    }
    for (int c = 0; c < 3; c++) f[3 * i + c] += fi[c];
  }

  double e = 0.0;
  for (int i = 0; i < nat; i++) e += pe[i];
  *epot += 0.5 * e;
  for (int i = 0; i < nat; i++) {
    if (epot_per_at) epot_per_at[i] += 0.5 * pe[i];
    for (int c = 0; c < 3; c++) f_inout[3 * i + c] += f[3 * i + c];
  }
  for (int c = 0; c < 9; c++) wpot_inout[c] += wpot[c];

  free(neb); free(nbb); free(bndtyp); free(bndlen); free(bndnm); free(cutfcnar); free(cutdrvar);
  free(cutfcnbo); free(cutdrvbo); free(sneb_seed); free(sneb_last); free(neb_seed); free(neb_last);
  free(pe); free(f); free(dbidk); free(zfaci); free(nebofi); free(slotofi); free(seedi); free(lasti);
  free(sneb); free(sbnd); free(cutdrarik); free(cutdrarjk); free(cutdrboik); free(cutdrbojk);
  free(sfacbo);
#undef cut_ar_h
  return 0;
}


/* test hook for the cutoff functions (FRUIT tests src/unittests/test_cutoff.f90): kind 0 = trig_off,
 * 1 = exp_cutoff */
void orc_cutoff_eval(int kind, double r1, double r2, double r, double *val, double *dval) {
  if (kind == 0) {
    trig_off(r1, r2, r, val, dval);
  } else {
    exp_cutoff_t t;
    exp_cutoff_init(&t, r1, r2);
    exp_cutoff_f(&t, r, val, dval);
  }
}

/* test hook for the per-pair / per-triplet functions of tersoff_func.f90, kumagai_func.f90,
 * brenner_func.f90 and juslin_func.f90 (tests/test_func_vs_reference.py evaluates the reference's
 * own source next to it).  which: 0 VA(dr), 1 VR(dr), 2 g(costh), 3 bo(zij; fcij, faij), 4 h(dr).
 * All indices are 1-based as in the Fortran. */
void orc_bop_func(const orc_bop_params_t *p, int which, int ktypj, int ktypi, int ktypk, int ijpot,
                  int ikpot, double x, double fcij, double faij, double *val, double *dval) {
  brenner_derived_t bd;
  if (p->kind == ORC_BRENNER || p->kind == ORC_JUSLIN) brenner_derive(p, &bd);
  switch (which) {
    case 0: f_VA(p, &bd, ijpot - 1, x, val, dval); break;
    case 1: f_VR(p, &bd, ijpot - 1, x, val, dval); break;
    case 2: f_g(p, &bd, ktypi - 1, ikpot - 1, x, val, dval); break;
    case 3: f_bo(p, &bd, ktypi - 1, ijpot - 1, x, fcij, faij, val, dval); break;
    default: f_h(p, ktypj, ktypi, ktypk, ikpot - 1, x, val, dval);
  }
}
