/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see atomistica_oracle.h).
 *
 * REBO2 (Brenner 2002) kernel, non-screened build (DIHEDRAL + NUM_NEIGHBORS,
 * rebo2.f90:58-60), PYTHON host, restated from
 *   src/potentials/bop/rebo2/bop_kernel_rebo2.f90:700-1200 (loop 1 + nn)
 *   src/potentials/bop/rebo2/bop_kernel_rebo2.f90:1209-2883 (loop 2)
 *   src/potentials/bop/rebo2/rebo2_func.f90 (fconj, fCin, VA, VR, g, bo, h, Z2pair)
 *   src/special/table2d.f90:255-318, table3d.f90:313-389 (eval)
 * Periodic-image identity along neighbour paths is tracked with 3-vector shift
 * sums instead of the reference's linearised dcell (identical whenever the
 * reference's encoding is unambiguous).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_oracle.h"

#define M3(M, i, j) (M)[(j) * 3 + (i)]
static const double PI = 3.14159265358979323846264338327950288;

enum { REBO2_C = 1, REBO2_H = 3 };
enum { C_C = 1, C_H = 3, H_H = 6 };

void orc_table2d_eval(const orc_table2d_t *t, double nhi, double nci, double *hch, double *dhchdh,
                      double *dhchdc) {
  int nboxs = t->nx * t->ny;
  int nhbox = (int)nhi;
  if (nhbox < 0) nhbox = 0;
  if (nhbox >= t->nx) nhbox = t->nx - 1;
  int ncbox = (int)nci;
  if (ncbox < 0) ncbox = 0;
  if (ncbox >= t->ny) ncbox = t->ny - 1;
  int ibox = t->ny * nhbox + ncbox; /* 0-based */
  double x1 = nhi - nhbox, x2 = nci - ncbox;
  *hch = 0.0; *dhchdh = 0.0; *dhchdc = 0.0;
  for (int i = 4; i >= 1; i--) {
    double shch = 0.0, shchdc = 0.0;
    for (int j = 4; j >= 1; j--) {
      double coefij = t->coeff[ibox + nboxs * ((i - 1) + 4 * (j - 1))];
      shch = shch * x2 + coefij;
      if (j > 1) shchdc = shchdc * x2 + (j - 1) * coefij;
    }
    *hch = *hch * x1 + shch;
    if (i > 1) *dhchdh = *dhchdh * x1 + (i - 1) * shch;
    *dhchdc = *dhchdc * x1 + shchdc;
  }
}

void orc_table3d_eval(const orc_table3d_t *t, double nti, double ntj, double nconji, double *fcc,
                      double *dfccdi, double *dfccdj, double *dfccdc) {
  int nboxs = t->nx * t->ny * t->nz;
  int nibox = (int)nti;
  if (nibox < 0) nibox = 0;
  if (nibox >= t->nx) nibox = t->nx - 1;
  int njbox = (int)ntj;
  if (njbox < 0) njbox = 0;
  if (njbox >= t->ny) njbox = t->ny - 1;
  int ncbox = (int)nconji;
  if (ncbox < 0) ncbox = 0;
  if (ncbox >= t->nz) ncbox = t->nz - 1;
  int ibox = t->nx * (t->ny * ncbox + njbox) + nibox; /* 0-based */
  double x1 = nti - nibox, x2 = ntj - njbox, x3 = nconji - ncbox;
  *fcc = 0.0; *dfccdi = 0.0; *dfccdj = 0.0; *dfccdc = 0.0;
  for (int i = 4; i >= 1; i--) {
    double sfcc = 0.0, sfccdj = 0.0, sfccdc = 0.0;
    for (int j = 4; j >= 1; j--) {
      double tfcc = 0.0, tfccdc = 0.0;
      for (int k = 4; k >= 1; k--) {
        double coefij = t->coeff[ibox + nboxs * ((i - 1) + 4 * ((j - 1) + 4 * (k - 1)))];
        tfcc = tfcc * x3 + coefij;
        if (k > 1) tfccdc = tfccdc * x3 + (k - 1) * coefij;
      }
      sfcc = sfcc * x2 + tfcc;
      if (j > 1) sfccdj = sfccdj * x2 + (j - 1) * tfcc;
      sfccdc = sfccdc * x2 + tfccdc;
    }
    *fcc = *fcc * x1 + sfcc;
    if (i > 1) *dfccdi = *dfccdi * x1 + (i - 1) * sfcc;
    *dfccdj = *dfccdj * x1 + sfccdj;
    *dfccdc = *dfccdc * x1 + sfccdc;
  }
}

/* rebo2_func.f90:31-57 */
static void fconj(double x, double *fx, double *dfx) {
  if (x <= 2.0) { *fx = 1.0; *dfx = 0.0; }
  else if (x >= 3.0) { *fx = 0.0; *dfx = 0.0; }
  else {
    double arg = PI * (x - 2.0);
    *fx = 0.5 * (1.0 + cos(arg));
    *dfx = -0.5 * PI * sin(arg);
  }
}

/* rebo2_func.f90:63-85 with CUTOFF_T = trig_off_t (cutoff.f90:152-196) */
static void fCin(const orc_rebo2_params_t *p, int ijpot, double dr, double *val, double *dval) {
  double l = p->cut_in_l[ijpot - 1], h = p->cut_in_h[ijpot - 1];
  if (dr > h) { *val = 0.0; *dval = 0.0; }
  else if (dr < l) { *val = 1.0; *dval = 0.0; }
  else {
    double fac = PI / (h - l);
    if (dr <= l) { *val = 1.0; *dval = 0.0; }
    else if (dr >= h) { *val = 0.0; *dval = 0.0; }
    else {
      double x = fac * (dr - l);
      *val = 0.5 * (1.0 + cos(x));
      *dval = -0.5 * fac * sin(x);
    }
  }
}

/* fCar / fCbo / fCnc of the screened variant (rebo2_func.f90:87-165): the same trig_off between
 * the pair's (l, h) of that cutoff family */
static void trig_cut(double l, double h, double dr, double *val, double *dval) {
  if (dr > h) { *val = 0.0; *dval = 0.0; }
  else if (dr < l) { *val = 1.0; *dval = 0.0; }
  else {
    double fac = PI / (h - l);
    if (dr <= l) { *val = 1.0; *dval = 0.0; }
    else if (dr >= h) { *val = 0.0; *dval = 0.0; }
    else {
      double x = fac * (dr - l);
      *val = 0.5 * (1.0 + cos(x));
      *dval = -0.5 * fac * sin(x);
    }
  }
}

/* rebo2_func.f90:173-221 */
static void VA(const orc_rebo2_params_t *p, int ijpot, double dr, double *val, double *dval) {
  if (ijpot == C_C) {
    double e1 = p->cc_B1 * exp(-p->cc_beta1 * dr);
    double e2 = p->cc_B2 * exp(-p->cc_beta2 * dr);
    double e3 = p->cc_B3 * exp(-p->cc_beta3 * dr);
    *val = -(e1 + e2 + e3);
    *dval = -(-p->cc_beta1 * e1 - p->cc_beta2 * e2 - p->cc_beta3 * e3);
  } else if (ijpot == C_H) {
    double e1 = p->ch_B1 * exp(-p->ch_beta1 * dr);
    *val = -e1;
    *dval = p->ch_beta1 * e1;
  } else {
    double e1 = p->hh_B1 * exp(-p->hh_beta1 * dr);
    *val = -e1;
    *dval = p->hh_beta1 * e1;
  }
}

/* rebo2_func.f90:230-277 */
static void VR(const orc_rebo2_params_t *p, int ijpot, double dr, double *val, double *dval) {
  double A, Q, alpha;
  if (ijpot == C_C) { A = p->cc_A; Q = p->cc_Q; alpha = p->cc_alpha; }
  else if (ijpot == C_H) { A = p->ch_A; Q = p->ch_Q; alpha = p->ch_alpha; }
  else { A = p->hh_A; Q = p->hh_Q; alpha = p->hh_alpha; }
  double e1 = A * exp(-alpha * dr);
  double hlp1 = 1 + Q / dr;
  *val = hlp1 * e1;
  *dval = (-Q / (dr * dr) - hlp1 * alpha) * e1;
}

static double ipow(double x, int n) {
  /* gfortran expands x**n for small integer n into multiplications */
  double v = 1.0;
  for (int i = 0; i < n; i++) v *= x;
  return v;
}

/* rebo2_func.f90:351-397 */
static void cc_g_from_spline(const orc_rebo2_params_t *p, const double *c, double costh,
                             double *val, double *dval) {
  int j;
  if (costh < p->cc_g_theta[1]) j = 0;
  else if (costh < p->cc_g_theta[2]) j = 1;
  else j = 2;
  const double *cj = &c[6 * j];
  double h = cj[0] + cj[1] * costh;
  double dh = cj[1];
  for (int i = 3; i <= 6; i++) {
    h = h + cj[i - 1] * ipow(costh, i - 1);
    dh = dh + (i - 1) * cj[i - 1] * ipow(costh, i - 2);
  }
  *val = h;
  *dval = dh;
}

/* rebo2_func.f90:289-347 */
static void gfun(const orc_rebo2_params_t *p, int ktyp, double costh, double n, double *val,
                 double *dval_dcosth, double *dval_dN) {
  if (ktyp == REBO2_C) {
    if (n < 3.2) {
      cc_g_from_spline(p, p->cc_g2_coeff, costh, val, dval_dcosth);
      *dval_dN = 0.0;
    } else if (n > (double)3.7f) { /* single-precision literal in the reference */
      cc_g_from_spline(p, p->cc_g1_coeff, costh, val, dval_dcosth);
      *dval_dN = 0.0;
    } else {
      double v1, v2, dv1, dv2;
      cc_g_from_spline(p, p->cc_g1_coeff, costh, &v1, &dv1);
      cc_g_from_spline(p, p->cc_g2_coeff, costh, &v2, &dv2);
      double arg = 2 * PI * (n - 3.2);
      double s = (1 + cos(arg)) / 2;
      double ds = -PI * sin(arg);
      *val = v1 * (1 - s) + v2 * s;
      *dval_dcosth = dv1 * (1 - s) + dv2 * s;
      *dval_dN = (v2 - v1) * ds;
    }
  } else {
    int ig = p->igh[(int)(-costh * 12.0) + 13 - 1];
    const double *s = &p->spgh[6 * (ig - 1)];
    *val = s[0] + s[1] * costh;
    *dval_dcosth = s[1];
    for (int i = 3; i <= 6; i++) {
      *val = *val + s[i - 1] * ipow(costh, i - 1);
      *dval_dcosth = *dval_dcosth + (i - 1) * s[i - 1] * ipow(costh, i - 2);
    }
    *dval_dN = 0.0; /* intent(out) left undefined by the reference; only ever multiplied into C terms */
  }
}

/* rebo2_func.f90:403-425 */
static void bo(const orc_rebo2_params_t *p, int ktypi, double zij, double fcij, double faij,
               double *bij, double *dfbij) {
  double arg = 1.0 + zij;
  *bij = pow(arg, p->conpe[ktypi - 1]);
  *dfbij = p->conan[ktypi - 1] * fcij * faij * pow(arg, p->conpf[ktypi - 1]);
}

/* rebo2_func.f90:431-461 */
static void hfun(const orc_rebo2_params_t *p, int ijpot, int ikpot, double dr, double *val,
                 double *dval) {
  if (ijpot + ikpot <= 4) { *val = 1.0; *dval = 0.0; }
  else {
    *val = p->conear[(ijpot - 1) + 6 * (ikpot - 1)] * exp(p->conalp * dr);
    *dval = p->conalp * (*val);
  }
}

/* rebo2_func.f90:467-485 */
static int Z2pair(int ktypi, int ktypj) {
  if (ktypi == REBO2_C) return ktypj;
  if (ktypj == REBO2_C) return ktypi;
  return ktypi + ktypj;
}

static void outer_add(double *w, double s, const double *a, const double *b) {
  for (int q = 0; q < 3; q++)
    for (int pq = 0; pq < 3; pq++) M3(w, pq, q) += s * (a[pq] * b[q]);
}

typedef struct { int s[3]; } shift_t;
static shift_t sadd(shift_t a, const int *b) {
  shift_t c = {{a.s[0] + b[0], a.s[1] + b[1], a.s[2] + b[2]}};
  return c;
}
static int szero(shift_t a) { return a.s[0] == 0 && a.s[1] == 0 && a.s[2] == 0; }
static int seq(shift_t a, shift_t b) { return a.s[0] == b.s[0] && a.s[1] == b.s[1] && a.s[2] == b.s[2]; }
/* reference j_gt_i test uses the sign of the linearised dcell: lexicographic sign of (x,y,z) */
static int spositive(const int *s) {
  if (s[0] != 0) return s[0] > 0;
  if (s[1] != 0) return s[1] > 0;
  return s[2] > 0;
}

/* One kernel for Rebo2 (scr == NULL) and Rebo2Scr (bop_kernel_rebo2.f90 compiled with SCREENING,
 * NUM_NEIGHBORS, ALT_DIHEDRAL; rebo2_scr.f90:60-64).  In the unscreened build the bond-order and
 * neighbour-count cutoffs are aliases of the single cutoff (bop_kernel_rebo2.f90:22-30), here the
 * three arrays then point to the same storage. */
static int rebo2_kernel(const orc_rebo2_params_t *par, const orc_rebo2_scr_t *scr, int nat, int natloc,
                        const double *r, const double *Abox, const int *ktyp,
                        const intptr_t *seed, const intptr_t *last, const int *neighbors,
                        const int *dc, double *epot, double *f_inout, double *wpot_inout,
                        double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                        double *wpot_per_at, double *wpot_per_bond) {
  const int typemax = 3;
  /* cutoff families, index ijpot-1 (rebo2_db.f90:170-253) */
  double cut_ar_l[10], cut_ar_h[10], cut_bo_l[10], cut_bo_h[10], cut_nc_l[10], cut_nc_h[10], max_cut_sq[10];
  for (int q = 0; q < 10; q++) {
    cut_ar_l[q] = cut_bo_l[q] = cut_nc_l[q] = par->cut_in_l[q];
    cut_ar_h[q] = cut_bo_h[q] = cut_nc_h[q] = par->cut_in_h[q];
    max_cut_sq[q] = par->cut_in_h2[q];
  }
  double Cmin = 0, Cmax = 0, dC = 0, C_dr_cut = 0;
  const double screening_threshold = log(1e-6), dot_threshold = 1e-10; /* rebo2_type.f90:68-69 */
  if (scr) {
    cut_ar_l[C_C - 1] = scr->cc_ar_r1; cut_ar_h[C_C - 1] = scr->cc_ar_r2;
    cut_bo_l[C_C - 1] = scr->cc_bo_r1; cut_bo_h[C_C - 1] = scr->cc_bo_r2;
    cut_nc_l[C_C - 1] = scr->cc_nc_r1; cut_nc_h[C_C - 1] = scr->cc_nc_r2;
    for (int q = 0; q < 10; q++) {
      double m = par->cut_in_h[q];
      if (cut_ar_h[q] > m) m = cut_ar_h[q];
      if (cut_bo_h[q] > m) m = cut_bo_h[q];
      if (cut_nc_h[q] > m) m = cut_nc_h[q];
      max_cut_sq[q] = m * m;
    }
    Cmin = scr->Cmin; Cmax = scr->Cmax;
    dC = Cmax - Cmin;
    C_dr_cut = Cmax * Cmax / (4 * (Cmax - 1)); /* rebo2_db.f90:105-108 */
  }
  long ntot = 0;
  int nebmax = 0;
  for (int i = 0; i < nat; i++) {
    int d = (int)(last[i] - seed[i] + 1);
    if (d > nebmax) nebmax = d;
    ntot += d;
  }
  if (nebmax < 1) nebmax = 1;
  long nebsize = ntot + nat + 1;
  int nm2 = nebmax * nebmax;

  int *neb = (int *)malloc(sizeof(int) * nebsize);
  long *nbb = (long *)malloc(sizeof(long) * nebsize);
  int *dcell = (int *)malloc(sizeof(int) * 3 * nebsize);
  int *bndtyp = (int *)malloc(sizeof(int) * nebsize);
  double *bndlen = (double *)malloc(sizeof(double) * nebsize);
  double *bndnm = (double *)malloc(sizeof(double) * 3 * nebsize);
  double *cutfcnar = (double *)malloc(sizeof(double) * nebsize);
  double *cutdrvar = (double *)malloc(sizeof(double) * nebsize);
  double *cutfcnbo = cutfcnar, *cutdrvbo = cutdrvar, *cutfcnnc = cutfcnar, *cutdrvnc = cutdrvar;
  long *sneb_seed = NULL, *sneb_last = NULL;
  if (scr) {
    cutfcnbo = (double *)malloc(sizeof(double) * nebsize);
    cutdrvbo = (double *)malloc(sizeof(double) * nebsize);
    cutfcnnc = (double *)malloc(sizeof(double) * nebsize);
    cutdrvnc = (double *)malloc(sizeof(double) * nebsize);
  }
  /* screening-neighbour bookkeeping (always allocated; empty ranges when unscreened) */
  sneb_seed = (long *)malloc(sizeof(long) * nebsize);
  sneb_last = (long *)malloc(sizeof(long) * nebsize);
  long scap = nebsize + 16, snebtot = 0;
  int *sneb = (int *)malloc(sizeof(int) * scap);
  long *sbnd = (long *)malloc(sizeof(long) * scap);
  double *cutdrarik = (double *)malloc(sizeof(double) * scap), *cutdrarjk = (double *)malloc(sizeof(double) * scap);
  double *cutdrboik = (double *)malloc(sizeof(double) * scap), *cutdrbojk = (double *)malloc(sizeof(double) * scap);
  double *cutdrncik = (double *)malloc(sizeof(double) * scap), *cutdrncjk = (double *)malloc(sizeof(double) * scap);
  long *neb_seed = (long *)malloc(sizeof(long) * (nat + 1));
  long *neb_last = (long *)malloc(sizeof(long) * (nat + 1));
  double *nn = (double *)calloc((size_t)typemax * (nat + 1), sizeof(double));
  double *pe = (double *)calloc(nat + 1, sizeof(double));
  double *f = (double *)calloc(3 * (nat + 1), sizeof(double));

#define DALLOC(n) ((double *)calloc((size_t)(n), sizeof(double)))
#define IALLOC(n) ((int *)calloc((size_t)(n), sizeof(int)))
  double *dbidk = DALLOC(3 * nebmax), *dbjdl = DALLOC(3 * nebmax);
  double *dnidk = DALLOC(3 * nebmax * typemax), *dnjdl = DALLOC(3 * nebmax * typemax);
  double *dnconjidxi = DALLOC(nebmax), *dnconjjdxj = DALLOC(nebmax);
  double *dncnidk = DALLOC(3 * nebmax), *dncnjdl = DALLOC(3 * nebmax);
  double *dncnidm = DALLOC(3 * nm2), *dncnjdn = DALLOC(3 * nm2);
  double *xikdm = DALLOC(3 * nebmax), *xjldn = DALLOC(3 * nebmax);
  double *fxik = DALLOC(nebmax), *fxjl = DALLOC(nebmax);
  int *nebofi = IALLOC(nebmax), *nebofj = IALLOC(nebmax);
  int *nebofk = IALLOC(nm2), *nebofl = IALLOC(nm2);
  long *slotofi = (long *)calloc(nebmax, sizeof(long));
  shift_t *dcofj = (shift_t *)calloc(nebmax, sizeof(shift_t));
  shift_t *dcofk = (shift_t *)calloc(nm2, sizeof(shift_t));
  shift_t *dcofl = (shift_t *)calloc(nm2, sizeof(shift_t));
  shift_t *dcofi = (shift_t *)calloc(nebmax, sizeof(shift_t));
  int *numnbk = IALLOC(nebmax + 2), *numnbl = IALLOC(nebmax + 2);
  long *seedi = (long *)calloc(nebmax + 1, sizeof(long)), *lasti = (long *)calloc(nebmax + 1, sizeof(long));
  long *seedj = (long *)calloc(nebmax + 1, sizeof(long)), *lastj = (long *)calloc(nebmax + 1, sizeof(long));
  long *seedk = (long *)calloc(nm2 + 1, sizeof(long)), *lastk = (long *)calloc(nm2 + 1, sizeof(long));
  long *seedl = (long *)calloc(nm2 + 1, sizeof(long)), *lastl = (long *)calloc(nm2 + 1, sizeof(long));
  double *zfaci = DALLOC(nebmax + 1), *zfacj = DALLOC(nebmax + 1);
  double *dri = DALLOC(3 * nebmax), *drj = DALLOC(3 * nebmax);
  double *drk = DALLOC(3 * nm2), *drl = DALLOC(3 * nm2);
/* dnidk(:, ikc, t) with t in 1..3 */
#define DN(a, c, ikc, t) (a)[(c) + 3 * ((ikc) + (long)nebmax * ((t)-1))]

  double wpot[9] = {0};
  int err = 0;

  /* loop 1 over ALL nat atoms: bop_kernel_rebo2.f90:700-1181 */
  long nebtot = 0;
  for (int i = 0; i < nat; i++) {
    int ktypi = ktyp[i];
    neb_seed[i] = nebtot;
    neb_last[i] = nebtot - 1;
    if (ktypi <= 0) continue;
    for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
      int j = neighbors[jn - 1] - 1;
      int ktypj = ktyp[j];
      if (ktypj <= 0) continue;
      double rij[3];
      for (int k = 0; k < 3; k++) {
        double s = 0.0;
        for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (jn - 1) + c];
        rij[k] = r[3 * j + k] - r[3 * i + k] - s;
      }
      double rlij = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
      int ijpot = Z2pair(ktypi, ktypj);
      double l = par->cut_in_l[ijpot - 1];
      sneb_seed[nebtot] = snebtot;
      sneb_last[nebtot] = snebtot - 1;
      if (rlij < l * l) {
        cutfcnar[nebtot] = 1.0;
        cutdrvar[nebtot] = 0.0;
        cutfcnbo[nebtot] = 1.0; cutdrvbo[nebtot] = 0.0;
        cutfcnnc[nebtot] = 1.0; cutdrvnc[nebtot] = 0.0;
        rlij = sqrt(rlij);
      } else if (scr) {
        /* bop_kernel_rebo2.f90:813-1101 */
        if (!(rlij < max_cut_sq[ijpot - 1])) continue;
        int screened = 0, need_derivative = 0;
        double sij = 0.0, dsijdrij = 0.0;
        long ineb = snebtot;
        if (ijpot == C_C) {
          intptr_t kn = seed[i];
          while (!(screened || sij < screening_threshold) && kn <= last[i]) {
            int k = neighbors[kn - 1] - 1;
            double rik[3];
            for (int a = 0; a < 3; a++) {
              double sh = 0.0;
              for (int c = 0; c < 3; c++) sh += M3(Abox, a, c) * (double)dc[3 * (kn - 1) + c];
              rik[a] = r[3 * k + a] - r[3 * i + a] - sh;
            }
            double rik2 = rik[0] * rik[0] + rik[1] * rik[1] + rik[2] * rik[2];
            if (rik2 < C_dr_cut * rlij) {
              int same = (k == j) && dc[3 * (kn - 1)] == dc[3 * (jn - 1)] &&
                         dc[3 * (kn - 1) + 1] == dc[3 * (jn - 1) + 1] && dc[3 * (kn - 1) + 2] == dc[3 * (jn - 1) + 2];
              if (!same) {
                double dot_ij_ik = rij[0] * rik[0] + rij[1] * rik[1] + rij[2] * rik[2];
                double rlik = rik2;
                double rjk[3] = {-rij[0] + rik[0], -rij[1] + rik[1], -rij[2] + rik[2]};
                double dot_ij_jk = rij[0] * rjk[0] + rij[1] * rjk[1] + rij[2] * rjk[2];
                double rljk = rjk[0] * rjk[0] + rjk[1] * rjk[1] + rjk[2] * rjk[2];
                if (dot_ij_ik > dot_threshold && dot_ij_jk < -dot_threshold) {
                  double xik = rlik / rlij, xjk = rljk / rlij;
                  double xm = xik - xjk, xp = xik + xjk;
                  double fac = 1.0 / (1 - xm * xm);
                  double C = (2 * xp - xm * xm - 1) * fac;
                  if (C <= Cmin) {
                    screened = 1;
                  } else if (C < Cmax) {
                    need_derivative = 1;
                    double Cmax_C = Cmax - C, C_Cmin = C - Cmin;
                    double q = Cmax_C / C_Cmin;
                    sij = sij - q * q;
                    double dCdrik = 4 * xik * fac * (1 + (C - 1) * xm);
                    double dCdrjk = 4 * xjk * fac * (1 - (C - 1) * xm);
                    double dCdrij = -(dCdrik + dCdrjk);
                    fac = 2 * Cmax_C * dC / (C_Cmin * C_Cmin * C_Cmin);
                    dsijdrij = dsijdrij + fac * dCdrij;
                    if (snebtot + 1 > scap) {
                      scap = scap * 2;
                      sneb = (int *)realloc(sneb, sizeof(int) * scap);
                      sbnd = (long *)realloc(sbnd, sizeof(long) * scap);
                      cutdrarik = (double *)realloc(cutdrarik, sizeof(double) * scap);
                      cutdrarjk = (double *)realloc(cutdrarjk, sizeof(double) * scap);
                      cutdrboik = (double *)realloc(cutdrboik, sizeof(double) * scap);
                      cutdrbojk = (double *)realloc(cutdrbojk, sizeof(double) * scap);
                      cutdrncik = (double *)realloc(cutdrncik, sizeof(double) * scap);
                      cutdrncjk = (double *)realloc(cutdrncjk, sizeof(double) * scap);
                    }
                    sneb[snebtot] = k;
                    sbnd[snebtot] = kn - 1;
                    cutdrarik[snebtot] = fac * dCdrik / rlik;
                    cutdrarjk[snebtot] = fac * dCdrjk / rljk;
                    sneb_last[nebtot] = snebtot;
                    snebtot++;
                  }
                }
              }
            }
            kn++;
          }
        }
        if ((screened || sij < screening_threshold) && rlij > par->cut_in_h2[ijpot - 1]) {
          snebtot = ineb;
          sneb_last[nebtot] = ineb - 1;
          continue; /* fully screened */
        }
        rlij = sqrt(rlij);
        double fcin, dfcin, fa, dfa, fb, dfb, fn, dfn;
        if (screened) {
          fCin(par, ijpot, rlij, &fcin, &dfcin);
          cutfcnar[nebtot] = fcin; cutdrvar[nebtot] = dfcin;
          cutfcnbo[nebtot] = fcin; cutdrvbo[nebtot] = dfcin;
          cutfcnnc[nebtot] = fcin; cutdrvnc[nebtot] = dfcin;
          snebtot = ineb;
          sneb_last[nebtot] = ineb - 1;
        } else if (need_derivative) {
          sij = exp(sij);
          fCin(par, ijpot, rlij, &fcin, &dfcin);
          trig_cut(cut_ar_l[ijpot - 1], cut_ar_h[ijpot - 1], rlij, &fa, &dfa);
          trig_cut(cut_bo_l[ijpot - 1], cut_bo_h[ijpot - 1], rlij, &fb, &dfb);
          trig_cut(cut_nc_l[ijpot - 1], cut_nc_h[ijpot - 1], rlij, &fn, &dfn);
          cutfcnar[nebtot] = (1.0 - fcin) * sij * fa + fcin;
          cutdrvar[nebtot] = (1.0 - fcin) * sij * (dfa + fa * dsijdrij / rlij) - dfcin * sij * fa + dfcin;
          cutfcnbo[nebtot] = (1.0 - fcin) * sij * fb + fcin;
          cutdrvbo[nebtot] = (1.0 - fcin) * sij * (dfb + fb * dsijdrij / rlij) - dfcin * sij * fb + dfcin;
          cutfcnnc[nebtot] = (1.0 - fcin) * sij * fn + fcin;
          cutdrvnc[nebtot] = (1.0 - fcin) * sij * (dfn + fn * dsijdrij / rlij) - dfcin * sij * fn + dfcin;
          for (long q = ineb; q < snebtot; q++) {
            cutdrboik[q] = cutdrarik[q] * sij * fb * (1.0 - fcin);
            cutdrbojk[q] = cutdrarjk[q] * sij * fb * (1.0 - fcin);
            cutdrncik[q] = cutdrarik[q] * sij * fn * (1.0 - fcin);
            cutdrncjk[q] = cutdrarjk[q] * sij * fn * (1.0 - fcin);
            cutdrarik[q] = cutdrarik[q] * sij * fa * (1.0 - fcin);
            cutdrarjk[q] = cutdrarjk[q] * sij * fa * (1.0 - fcin);
          }
        } else {
          trig_cut(cut_ar_l[ijpot - 1], cut_ar_h[ijpot - 1], rlij, &fa, &dfa);
          trig_cut(cut_bo_l[ijpot - 1], cut_bo_h[ijpot - 1], rlij, &fb, &dfb);
          trig_cut(cut_nc_l[ijpot - 1], cut_nc_h[ijpot - 1], rlij, &fn, &dfn);
          if (rlij < par->cut_in_h[ijpot - 1]) {
            fCin(par, ijpot, rlij, &fcin, &dfcin);
            cutfcnar[nebtot] = (1.0 - fcin) * fa + fcin;
            cutdrvar[nebtot] = (1.0 - fcin) * dfa - dfcin * fa + dfcin;
            cutfcnbo[nebtot] = (1.0 - fcin) * fb + fcin;
            cutdrvbo[nebtot] = (1.0 - fcin) * dfb - dfcin * fb + dfcin;
            cutfcnnc[nebtot] = (1.0 - fcin) * fn + fcin;
            cutdrvnc[nebtot] = (1.0 - fcin) * dfn - dfcin * fn + dfcin;
          } else {
            cutfcnar[nebtot] = fa; cutdrvar[nebtot] = dfa;
            cutfcnbo[nebtot] = fb; cutdrvbo[nebtot] = dfb;
            cutfcnnc[nebtot] = fn; cutdrvnc[nebtot] = dfn;
          }
        }
      } else if (rlij < par->cut_in_h2[ijpot - 1]) {
        rlij = sqrt(rlij);
        fCin(par, ijpot, rlij, &cutfcnar[nebtot], &cutdrvar[nebtot]);
      } else
        continue;
      neb[nebtot] = j;
      nbb[nebtot] = jn - 1;
      for (int c = 0; c < 3; c++) dcell[3 * nebtot + c] = dc[3 * (jn - 1) + c];
      bndlen[nebtot] = rlij;
      for (int c = 0; c < 3; c++) bndnm[3 * nebtot + c] = rij[c] / rlij;
      bndtyp[nebtot] = ijpot;
      neb_last[i] = nebtot;
      nebtot++;
    }
  }

  /* bop_kernel_rebo2.f90:686-688: screening force factors start at zero */
  double *sfacbo = DALLOC(snebtot + 1), *sfacnc = DALLOC(snebtot + 1);

  /* nn: bop_kernel_rebo2.f90:1189-1200 */
  for (int i = 0; i < nat; i++)
    for (long jn = neb_seed[i]; jn <= neb_last[i]; jn++) {
      int j = neb[jn];
      if (ktyp[j] > 0) nn[(ktyp[j] - 1) + typemax * i] += cutfcnnc[jn];
    }
#define NN(t, i) nn[((t)-1) + typemax * (i)]

  /* loop 2: bop_kernel_rebo2.f90:1209-2781 */
  for (int i = 0; i < natloc; i++) {
    int ktypi = ktyp[i];
    if (ktypi <= 0) continue;
    double fi[3] = {0, 0, 0};
    long istart = neb_seed[i], ifinsh = neb_last[i];
    double nconjit = 0.0;
    memset(dnidk, 0, sizeof(double) * 3 * nebmax * typemax);
    int ikc = 0, kmc = 0;
    numnbk[0] = 0;
    /* ik_loop1: :1231-1317 */
    for (long ik = istart; ik <= ifinsh; ik++, ikc++) {
      int k = neb[ik];
      shift_t kdc = {{dcell[3 * ik], dcell[3 * ik + 1], dcell[3 * ik + 2]}};
      nebofi[ikc] = k;
      dcofi[ikc] = kdc;
      slotofi[ikc] = ik;
      seedi[ikc] = sneb_seed[ik];
      lasti[ikc] = sneb_last[ik];
      int ktypk = ktyp[k];
      double rlik = bndlen[ik];
      const double *rnik = &bndnm[3 * ik];
      for (int c = 0; c < 3; c++) dri[3 * ikc + c] = rlik * rnik[c];
      double fcik = cutfcnnc[ik], dfcikr = cutdrvnc[ik];
      for (int c = 0; c < 3; c++) DN(dnidk, c, ikc, ktypk) = rnik[c] * dfcikr;
      if (ktypk == REBO2_C) {
        int nk = (int)(neb_last[k] - neb_seed[k] + 1);
        for (long km = neb_seed[k]; km <= neb_last[k]; km++) {
          int o = (int)(km - neb_seed[k]);
          nebofk[kmc + o] = neb[km];
          dcofk[kmc + o] = sadd(kdc, &dcell[3 * km]);
          seedk[kmc + o] = sneb_seed[km];
          lastk[kmc + o] = sneb_last[km];
          for (int c = 0; c < 3; c++) {
            drk[3 * (kmc + o) + c] = bndlen[km] * bndnm[3 * km + c];
            xikdm[3 * o + c] = cutdrvnc[km] * bndnm[3 * km + c];
          }
        }
        double xik = NN(REBO2_C, k) + NN(REBO2_H, k) - fcik;
        int first = kmc;
        kmc += nk;
        numnbk[ikc + 1] = kmc;
        double dfxikx;
        fconj(xik, &fxik[ikc], &dfxikx);
        double nconjdr = fxik[ikc] * dfcikr;
        double nconjdx = fcik * dfxikx;
        dnconjidxi[ikc] = nconjdx;
        nconjit = nconjit + fcik * fxik[ikc];
        for (int c = 0; c < 3; c++) dncnidk[3 * ikc + c] = nconjdr * rnik[c];
        for (int o = 0; o < nk; o++)
          for (int c = 0; c < 3; c++) dncnidm[3 * (first + o) + c] = nconjdx * xikdm[3 * o + c];
      } else {
        numnbk[ikc + 1] = kmc;
        fxik[ikc] = 0.0;
        for (int c = 0; c < 3; c++) dncnidk[3 * ikc + c] = 0.0;
      }
    }

    for (long ij = istart; ij <= ifinsh; ij++) {
      int j = neb[ij];
      shift_t jdc = {{dcell[3 * ij], dcell[3 * ij + 1], dcell[3 * ij + 2]}};
      /* j_gt_i :1332 */
      if (!((szero(jdc) && j > i) || spositive(jdc.s))) continue;
      int ijpot = bndtyp[ij];
      double rlij = bndlen[ij];
      if (!(rlij < cut_ar_h[ijpot - 1])) continue;

      double fj[3] = {0, 0, 0};
      int ktypj = ktyp[j];
      double rlijr = 1.0 / rlij;
      const double *rnij = &bndnm[3 * ij];
      double rij[3] = {rlij * rnij[0], rlij * rnij[1], rlij * rnij[2]};
      double fcarij = cutfcnar[ij], dfcarijr = cutdrvar[ij];
      double ni[4], nj[4]; /* index by type 1..3 */
      for (int t = 1; t <= 3; t++) { ni[t] = NN(t, i); nj[t] = NN(t, j); }
      ni[ktypj] = ni[ktypj] - cutfcnnc[ij];
      nj[ktypi] = nj[ktypi] - cutfcnnc[ij];
      double nconjj = 0.0, nconji = 0.0;
      memset(dnjdl, 0, sizeof(double) * 3 * nebmax * typemax);
      if (ni[REBO2_C] > 4.0) ni[REBO2_C] = 4.0;
      if (ni[REBO2_H] > 4.0) ni[REBO2_H] = 4.0;
      double nti = ni[REBO2_C] + ni[REBO2_H];
      if (nj[REBO2_C] > 4.0) nj[REBO2_C] = 4.0;
      if (nj[REBO2_H] > 4.0) nj[REBO2_H] = 4.0;
      double ntj = nj[REBO2_C] + nj[REBO2_H];
      double faij, dfaijr, frij, dfrijr;
      VA(par, ijpot, rlij, &faij, &dfaijr);
      VR(par, ijpot, rlij, &frij, &dfrijr);
      double wij[9] = {0}, wijb[9] = {0}, wjib[9] = {0};
      double zij = 0.0, dbidi[3] = {0, 0, 0}, dbidj[3] = {0, 0, 0}, dzdni = 0.0;

      /* ik_loop2 :1407-1587 */
      ikc = 0;
      for (long ik = istart; ik <= ifinsh; ik++, ikc++) {
        double fcik = cutfcnbo[ik];
        if (ik != ij) {
          int ikpot = bndtyp[ik];
          double rlik = bndlen[ik];
          if (rlik < cut_bo_h[ikpot - 1]) {
            const double *rnik = &bndnm[3 * ik];
            double rik[3] = {rlik * rnik[0], rlik * rnik[1], rlik * rnik[2]};
            double dfcikr = cutdrvbo[ik];
            double qfacan, qfadan, gfacan, gddan, dgdn;
            hfun(par, ijpot, ikpot, rlij - rlik, &qfacan, &qfadan);
            double costh = rnik[0] * rnij[0] + rnik[1] * rnij[1] + rnik[2] * rnij[2];
            gfun(par, ktypi, costh, nti, &gfacan, &gddan, &dgdn);
            double dkc[3];
            for (int c = 0; c < 3; c++) dkc[c] = rnik[c] * rlik - rnij[c] * rlij;
            double disjk = sqrt(dkc[0] * dkc[0] + dkc[1] * dkc[1] + dkc[2] * dkc[2]);
            for (int c = 0; c < 3; c++) dkc[c] = dkc[c] / disjk;
            double dcsdij = 1.0 / rlik - costh * rlijr;
            double dcsdik = rlijr - costh / rlik;
            double dcsdjk = -disjk * rlijr / rlik;
            dzdni = dzdni + fcik * dgdn * qfacan;
            double dzfac = fcik * gddan * qfacan;
            zfaci[ikc] = gfacan * qfacan; /* saved for the screening-function derivative */
            zij = zij + fcik * gfacan * qfacan;
            double dzdrij = gfacan * fcik * qfadan;
            double dzdrik = gfacan * (dfcikr * qfacan - fcik * qfadan);
            double df[3];
            for (int c = 0; c < 3; c++) {
              double dcsdi = -dcsdij * rnij[c] - dcsdik * rnik[c];
              double dcsdj = dcsdij * rnij[c] - dcsdjk * dkc[c];
              double dcsdk = dcsdik * rnik[c] + dcsdjk * dkc[c];
              dbidi[c] = dbidi[c] - dzdrij * rnij[c] - dzdrik * rnik[c] + dzfac * dcsdi;
              df[c] = dzdrij * rnij[c] + dzfac * dcsdj;
              dbidj[c] = dbidj[c] + df[c];
              dbidk[3 * ikc + c] = dzdrik * rnik[c] + dzfac * dcsdk;
            }
            outer_add(wijb, -1.0, rij, df);
            outer_add(wijb, -1.0, rik, &dbidk[3 * ikc]);
          } else {
            zfaci[ikc] = 0.0;
            for (int c = 0; c < 3; c++) dbidk[3 * ikc + c] = 0.0;
          }
        } else {
          fcik = cutfcnnc[ik];
          nconji = nconjit - fcik * fxik[ikc];
        }
      }
      int numnbi = ikc;

      double pij = 0.0, dpdnci = 0.0, dpdnhi = 0.0;
      if (ktypi == REBO2_C) {
        if (ijpot == C_C) orc_table2d_eval(&par->Pcc, ni[REBO2_H], ni[REBO2_C], &pij, &dpdnhi, &dpdnci);
        else orc_table2d_eval(&par->Pch, ni[REBO2_H], ni[REBO2_C], &pij, &dpdnhi, &dpdnci);
        zij = zij + pij;
        dpdnci = dpdnci + dzdni;
        dpdnhi = dpdnhi + dzdni;
      }
      double bij, dfbij;
      bo(par, ktypi, zij, fcarij, faij, &bij, &dfbij);

      /* jl_loop :1644-1887 */
      double zji = 0.0, dbjdi[3] = {0, 0, 0}, dbjdj[3] = {0, 0, 0}, dzdnj = 0.0;
      int jlc = 0, lnc = 0;
      numnbl[0] = 0;
      for (long jl = neb_seed[j]; jl <= neb_last[j]; jl++) {
        int l = neb[jl];
        shift_t ldc = sadd(jdc, &dcell[3 * jl]);
        if (!(l != i || !szero(ldc))) continue;
        nebofj[jlc] = l;
        dcofj[jlc] = ldc;
        seedj[jlc] = sneb_seed[jl];
        lastj[jlc] = sneb_last[jl];
        int ktypl = ktyp[l];
        int jlpot = bndtyp[jl];
        double rljl = bndlen[jl];
        const double *rnjl = &bndnm[3 * jl];
        for (int c = 0; c < 3; c++) drj[3 * jlc + c] = rljl * rnjl[c];
        double fcjl = cutfcnnc[jl], dfcjlr = cutdrvnc[jl];
        for (int c = 0; c < 3; c++) DN(dnjdl, c, jlc, ktypl) = rnjl[c] * dfcjlr;
        if (ktypl == REBO2_C) {
          int nl_ = (int)(neb_last[l] - neb_seed[l] + 1);
          for (long ln = neb_seed[l]; ln <= neb_last[l]; ln++) {
            int o = (int)(ln - neb_seed[l]);
            nebofl[lnc + o] = neb[ln];
            dcofl[lnc + o] = sadd(ldc, &dcell[3 * ln]);
            seedl[lnc + o] = sneb_seed[ln];
            lastl[lnc + o] = sneb_last[ln];
            for (int c = 0; c < 3; c++) {
              drl[3 * (lnc + o) + c] = bndlen[ln] * bndnm[3 * ln + c];
              xjldn[3 * o + c] = cutdrvnc[ln] * bndnm[3 * ln + c];
            }
          }
          double xjl = NN(REBO2_C, l) + NN(REBO2_H, l) - fcjl;
          int first = lnc;
          lnc += nl_;
          numnbl[jlc + 1] = lnc;
          double dfxjlx;
          fconj(xjl, &fxjl[jlc], &dfxjlx);
          double nconjdr = fxjl[jlc] * dfcjlr;
          double nconjdx = fcjl * dfxjlx;
          dnconjjdxj[jlc] = nconjdx;
          nconjj = nconjj + fcjl * fxjl[jlc];
          for (int c = 0; c < 3; c++) dncnjdl[3 * jlc + c] = nconjdr * rnjl[c];
          for (int o = 0; o < nl_; o++)
            for (int c = 0; c < 3; c++) dncnjdn[3 * (first + o) + c] = nconjdx * xjldn[3 * o + c];
        } else {
          numnbl[jlc + 1] = lnc;
          fxjl[jlc] = 0.0;
          for (int c = 0; c < 3; c++) dncnjdl[3 * jlc + c] = 0.0;
        }
        if (rljl < cut_bo_h[jlpot - 1]) {
          fcjl = cutfcnbo[jl];       /* the angular part uses the bond-order cutoff (:1755-1756) */
          dfcjlr = cutdrvbo[jl];
          double rjl[3] = {rljl * rnjl[0], rljl * rnjl[1], rljl * rnjl[2]};
          double qfacan, qfadan, gfacan, gddan, dgdn;
          hfun(par, ijpot, jlpot, rlij - rljl, &qfacan, &qfadan);
          double costh = -(rnjl[0] * rnij[0] + rnjl[1] * rnij[1] + rnjl[2] * rnij[2]);
          gfun(par, ktypj, costh, ntj, &gfacan, &gddan, &dgdn);
          double dlc[3];
          for (int c = 0; c < 3; c++) dlc[c] = rnjl[c] * rljl + rnij[c] * rlij;
          double disil = sqrt(dlc[0] * dlc[0] + dlc[1] * dlc[1] + dlc[2] * dlc[2]);
          for (int c = 0; c < 3; c++) dlc[c] = dlc[c] / disil;
          double dcsdji = 1.0 / rljl - costh * rlijr;
          double dcsdjl = rlijr - costh / rljl;
          double dcsdil = -disil * rlijr / rljl;
          dzdnj = dzdnj + fcjl * dgdn * qfacan;
          double dzfac = fcjl * gddan * qfacan;
          zfacj[jlc] = gfacan * qfacan;
          zji = zji + fcjl * gfacan * qfacan;
          double dzdrji = gfacan * fcjl * qfadan;
          double dzdrjl = gfacan * (dfcjlr * qfacan - fcjl * qfadan);
          double df[3];
          for (int c = 0; c < 3; c++) {
            double dcsdj = dcsdji * rnij[c] - dcsdjl * rnjl[c];
            double dcsdi = -dcsdji * rnij[c] - dcsdil * dlc[c];
            double dcsdl = dcsdjl * rnjl[c] + dcsdil * dlc[c];
            dbjdj[c] = dbjdj[c] + dzdrji * rnij[c] - dzdrjl * rnjl[c] + dzfac * dcsdj;
            df[c] = -dzdrji * rnij[c] + dzfac * dcsdi;
            dbjdi[c] = dbjdi[c] + df[c];
            dbjdl[3 * jlc + c] = dzdrjl * rnjl[c] + dzfac * dcsdl;
          }
          outer_add(wjib, 1.0, rij, df);
          outer_add(wjib, -1.0, rjl, &dbjdl[3 * jlc]);
        } else {
          zfacj[jlc] = 0.0;
          for (int c = 0; c < 3; c++) dbjdl[3 * jlc + c] = 0.0;
        }
        jlc++;
      }
      int numnbj = jlc;

      double pji = 0.0, dpdncj = 0.0, dpdnhj = 0.0;
      if (ktypj == REBO2_C) {
        if (ijpot == C_C) orc_table2d_eval(&par->Pcc, nj[REBO2_H], nj[REBO2_C], &pji, &dpdnhj, &dpdncj);
        else orc_table2d_eval(&par->Pch, nj[REBO2_H], nj[REBO2_C], &pji, &dpdnhj, &dpdncj);
        zji = zji + pji;
        dpdncj = dpdncj + dzdnj;
        dpdnhj = dpdnhj + dzdnj;
      }
      double bji, dfbji;
      bo(par, ktypj, zji, fcarij, faij, &bji, &dfbji);

      double nconj = nconji * nconji + nconjj * nconjj;
      if (nconj > 8.0) nconj = 8.0;
      if (nti > 3.0) nti = 3.0;
      if (ntj > 3.0) ntj = 3.0;

      double bdh = 0.0, tij = 0.0, dtdni = 0.0, dtdnj = 0.0, dtdncn = 0.0;
      if (scr && par->with_dihedral && ijpot == C_C) {
        /* ALT_DIHEDRAL (the dihedral term of the screened build, rebo2_scr.f90:62), :2089-2371: the angle
         * between the planes (r_ij, r_k1k2) and (r_ij, r_l1l2) over the PAIRS k1 < k2 of bond partners of i and
         * l1 < l2 of j; T_ij enters doubled (:2109-2112) */
        orc_table3d_eval(&par->Tcc, nti, ntj, nconj, &tij, &dtdni, &dtdnj, &dtdncn);
        tij = 2 * tij; dtdni = 2 * dtdni; dtdnj = 2 * dtdnj; dtdncn = 2 * dtdncn;
        double tije = tij * faij * fcarij;
        if (tij != 0) {
          const double rlijsq = rlij * rlij;
          const double rij[3] = {rlij * rnij[0], rlij * rnij[1], rlij * rnij[2]};
          for (long ik1 = istart; ik1 <= ifinsh - 1; ik1++) {
            if (ik1 == ij) continue;
            int k1 = neb[ik1];
            shift_t kdc1 = dcofi[ik1 - istart];
            double rlik1 = bndlen[ik1];
            if (!(rlik1 < cut_bo_h[bndtyp[ik1] - 1])) continue;
            const double *rnik1 = &bndnm[3 * ik1];
            double fcik1 = cutfcnbo[ik1], dfcik1r = cutdrvbo[ik1];
            for (long ik2 = ik1 + 1; ik2 <= ifinsh; ik2++) {
              if (ik2 == ij) continue;
              int k2 = neb[ik2];
              shift_t kdc2 = dcofi[ik2 - istart];
              double rlik2 = bndlen[ik2];
              if (!(rlik2 < cut_bo_h[bndtyp[ik2] - 1])) continue;
              const double *rnik2 = &bndnm[3 * ik2];
              double rk1k2[3];
              for (int c = 0; c < 3; c++) rk1k2[c] = rlik2 * rnik2[c] - rlik1 * rnik1[c];
              double dot_ij_k1k2 = rij[0] * rk1k2[0] + rij[1] * rk1k2[1] + rij[2] * rk1k2[2];
              double k1k2sq = rk1k2[0] * rk1k2[0] + rk1k2[1] * rk1k2[1] + rk1k2[2] * rk1k2[2];
              double dck1k2 = rlijsq * k1k2sq - dot_ij_k1k2 * dot_ij_k1k2;
              double fcik2 = cutfcnbo[ik2], dfcik2r = cutdrvbo[ik2];
              for (long jl1 = neb_seed[j]; jl1 <= neb_last[j] - 1; jl1++) {
                int l1 = neb[jl1];
                shift_t ldc1 = sadd(jdc, &dcell[3 * jl1]);
                if (!((l1 != i || !szero(ldc1)) && (l1 != k1 || !seq(ldc1, kdc1)) && (l1 != k2 || !seq(ldc1, kdc2))))
                  continue;
                double rljl1 = bndlen[jl1];
                if (!(rljl1 < cut_bo_h[bndtyp[jl1] - 1])) continue;
                const double *rnjl1 = &bndnm[3 * jl1];
                double fcjl1 = cutfcnbo[jl1], dfcjl1r = cutdrvbo[jl1];
                for (long jl2 = jl1 + 1; jl2 <= neb_last[j]; jl2++) {
                  int l2 = neb[jl2];
                  shift_t ldc2 = sadd(jdc, &dcell[3 * jl2]);
                  if (!((l2 != i || !szero(ldc2)) && (l2 != k1 || !seq(ldc2, kdc1)) && (l2 != k2 || !seq(ldc2, kdc2))))
                    continue;
                  double rljl2 = bndlen[jl2];
                  if (!(rljl2 < cut_bo_h[bndtyp[jl2] - 1])) continue;
                  const double *rnjl2 = &bndnm[3 * jl2];
                  double fcjl2 = cutfcnbo[jl2], dfcjl2r = cutdrvbo[jl2];
                  double rl1l2[3];
                  for (int c = 0; c < 3; c++) rl1l2[c] = rljl2 * rnjl2[c] - rljl1 * rnjl1[c];
                  double dot_ij_l1l2 = rij[0] * rl1l2[0] + rij[1] * rl1l2[1] + rij[2] * rl1l2[2];
                  double dot_k1k2_l1l2 = rk1k2[0] * rl1l2[0] + rk1k2[1] * rl1l2[1] + rk1k2[2] * rl1l2[2];
                  double l1l2sq = rl1l2[0] * rl1l2[0] + rl1l2[1] * rl1l2[1] + rl1l2[2] * rl1l2[2];
                  double dcl1l2 = rlijsq * l1l2sq - dot_ij_l1l2 * dot_ij_l1l2;
                  double abs_dc = sqrt(dck1k2 * dcl1l2);
                  double costijkl = (dot_ij_k1k2 * dot_ij_l1l2 - rlijsq * dot_k1k2_l1l2) / abs_dc;
                  double bdhij = 1 - costijkl * costijkl;
                  bdh = bdh + bdhij * fcik1 * fcik2 * fcjl1 * fcjl2;
                  bdhij = bdhij * tij * faij * fcarij / 2;
                  double dbdhij = -2 * costijkl * tije * fcik1 * fcik2 * fcjl1 * fcjl2 / 2;
                  double df[3], v[3];
                  for (int c = 0; c < 3; c++)
                    df[c] = dbdhij * ((dot_ij_l1l2 / abs_dc + costijkl * dot_ij_k1k2 / dck1k2) * rk1k2[c] +
                                      (dot_ij_k1k2 / abs_dc + costijkl * dot_ij_l1l2 / dcl1l2) * rl1l2[c] -
                                      (2 * dot_k1k2_l1l2 / abs_dc + costijkl * (k1k2sq / dck1k2 + l1l2sq / dcl1l2)) * rij[c]);
                  for (int c = 0; c < 3; c++) { fi[c] += df[c]; fj[c] -= df[c]; }
                  outer_add(wij, 1.0, rij, df);
                  for (int c = 0; c < 3; c++)
                    df[c] = dbdhij * (-(1.0 / dck1k2 * costijkl * rk1k2[c] + 1.0 / abs_dc * rl1l2[c]) * rlijsq +
                                      (dot_ij_l1l2 / abs_dc + costijkl * dot_ij_k1k2 / dck1k2) * rij[c]);
                  for (int c = 0; c < 3; c++) { f[3 * k1 + c] += df[c]; f[3 * k2 + c] += -df[c]; }
                  outer_add(wij, 1.0, rk1k2, df);
                  for (int c = 0; c < 3; c++)
                    df[c] = dbdhij * (-(1.0 / dcl1l2 * costijkl * rl1l2[c] + 1.0 / abs_dc * rk1k2[c]) * rlijsq +
                                      (dot_ij_k1k2 / abs_dc + costijkl * dot_ij_l1l2 / dcl1l2) * rij[c]);
                  for (int c = 0; c < 3; c++) { f[3 * l1 + c] += df[c]; f[3 * l2 + c] += -df[c]; }
                  outer_add(wij, 1.0, rl1l2, df);
                  for (int c = 0; c < 3; c++) df[c] = bdhij * dfcik1r * fcik2 * fcjl1 * fcjl2 * rnik1[c];
                  for (int c = 0; c < 3; c++) { fi[c] += df[c]; f[3 * k1 + c] += -df[c]; v[c] = rlik1 * rnik1[c]; }
                  outer_add(wij, 1.0, v, df);
                  for (int c = 0; c < 3; c++) df[c] = bdhij * dfcik2r * fcik1 * fcjl1 * fcjl2 * rnik2[c];
                  for (int c = 0; c < 3; c++) { fi[c] += df[c]; f[3 * k2 + c] += -df[c]; v[c] = rlik2 * rnik2[c]; }
                  outer_add(wij, 1.0, v, df);
                  for (int c = 0; c < 3; c++) df[c] = bdhij * dfcjl1r * fcjl2 * fcik1 * fcik2 * rnjl1[c];
                  for (int c = 0; c < 3; c++) { fj[c] += df[c]; f[3 * l1 + c] += -df[c]; v[c] = rljl1 * rnjl1[c]; }
                  outer_add(wij, 1.0, v, df);
                  for (int c = 0; c < 3; c++) df[c] = bdhij * dfcjl2r * fcjl1 * fcik1 * fcik2 * rnjl2[c];
                  for (int c = 0; c < 3; c++) { fj[c] += df[c]; f[3 * l2 + c] += -df[c]; v[c] = rljl2 * rnjl2[c]; }
                  outer_add(wij, 1.0, v, df);
                  /* screening neighbours of the four bonds (:2322-2354) */
                  double dffac = bdhij * fcik2 * fcjl1 * fcjl2;
                  for (long q = sneb_seed[ik1]; q <= sneb_last[ik1]; q++) sfacbo[q] += dffac;
                  dffac = bdhij * fcik1 * fcjl1 * fcjl2;
                  for (long q = sneb_seed[ik2]; q <= sneb_last[ik2]; q++) sfacbo[q] += dffac;
                  dffac = bdhij * fcjl2 * fcik1 * fcik2;
                  for (long q = sneb_seed[jl1]; q <= sneb_last[jl1]; q++) sfacbo[q] += dffac;
                  dffac = bdhij * fcjl1 * fcik1 * fcik2;
                  for (long q = sneb_seed[jl2]; q <= sneb_last[jl2]; q++) sfacbo[q] += dffac;
                }
              }
            }
          }
        }
      } else if (par->with_dihedral && ijpot == C_C) {
        /* :1950-2087 */
        orc_table3d_eval(&par->Tcc, nti, ntj, nconj, &tij, &dtdni, &dtdnj, &dtdncn);
        double tije = tij * faij * fcarij;
        if (tij != 0) {
          int ikc3 = 0;
          for (long ik = istart; ik <= ifinsh; ik++, ikc3++) {
            if (ik == ij) continue;
            int k = neb[ik];
            shift_t kdc = dcofi[ikc3];
            double rlik = bndlen[ik];
            const double *rnik = &bndnm[3 * ik];
            double fcik = cutfcnbo[ik], dfcikr = cutdrvbo[ik];
            double dot_ij_ik = rnij[0] * rnik[0] + rnij[1] * rnik[1] + rnij[2] * rnik[2];
            double dcik = 1.0 - dot_ij_ik * dot_ij_ik;
            for (long jl = neb_seed[j]; jl <= neb_last[j]; jl++) {
              int l = neb[jl];
              shift_t ldc = sadd(jdc, &dcell[3 * jl]);
              if ((l != i || !szero(ldc)) && (l != k || !seq(ldc, kdc))) {
                double rljl = bndlen[jl];
                const double *rnjl = &bndnm[3 * jl];
                double fcjl = cutfcnbo[jl], dfcjlr = cutdrvbo[jl];
                double dot_ij_jl = rnij[0] * rnjl[0] + rnij[1] * rnjl[1] + rnij[2] * rnjl[2];
                double dot_ik_jl = rnik[0] * rnjl[0] + rnik[1] * rnjl[1] + rnik[2] * rnjl[2];
                double dcjl = 1.0 - dot_ij_jl * dot_ij_jl;
                double abs_dc = sqrt(dcik * dcjl);
                double costijkl = (dot_ij_ik * dot_ij_jl - dot_ik_jl) / abs_dc;
                double bdhij = 1 - costijkl * costijkl;
                bdh = bdh + bdhij * fcik * fcjl;
                bdhij = bdhij * tij * faij * fcarij / 2;
                double dbdhij = -2 * costijkl * tije * fcik * fcjl / 2;
                double df[3], v[3];
                for (int c = 0; c < 3; c++)
                  df[c] = dbdhij *
                          ((dot_ij_jl / abs_dc + costijkl * dot_ij_ik / dcik) * rnik[c] +
                           (dot_ij_ik / abs_dc + costijkl * dot_ij_jl / dcjl) * rnjl[c] -
                           (2 * dot_ik_jl / abs_dc + costijkl * (1.0 / dcik + 1.0 / dcjl)) * rnij[c]) /
                          rlij;
                for (int c = 0; c < 3; c++) { fi[c] += df[c]; fj[c] -= df[c]; v[c] = rlij * rnij[c]; }
                outer_add(wij, 1.0, v, df);
                for (int c = 0; c < 3; c++)
                  df[c] = dbdhij *
                              (-1.0 / dcik * costijkl * rnik[c] - 1.0 / abs_dc * rnjl[c] +
                               (dot_ij_jl / abs_dc + costijkl * dot_ij_ik / dcik) * rnij[c]) /
                              rlik +
                          bdhij * dfcikr * fcjl * rnik[c];
                for (int c = 0; c < 3; c++) { fi[c] += df[c]; f[3 * k + c] += -df[c]; v[c] = rlik * rnik[c]; }
                outer_add(wij, 1.0, v, df);
                for (int c = 0; c < 3; c++)
                  df[c] = dbdhij *
                              (-1.0 / dcjl * costijkl * rnjl[c] - 1.0 / abs_dc * rnik[c] +
                               (dot_ij_ik / abs_dc + costijkl * dot_ij_jl / dcjl) * rnij[c]) /
                              rljl +
                          bdhij * fcik * dfcjlr * rnjl[c];
                for (int c = 0; c < 3; c++) { fj[c] += df[c]; f[3 * l + c] += -df[c]; v[c] = rljl * rnjl[c]; }
                outer_add(wij, 1.0, v, df);
              }
            }
          }
        }
      }

      double fij = 0.0, dfdni = 0.0, dfdnj = 0.0, dfdncn = 0.0;
      if (ijpot == C_C) orc_table3d_eval(&par->Fcc, nti, ntj, nconj, &fij, &dfdni, &dfdnj, &dfdncn);
      else if (ijpot == H_H) orc_table3d_eval(&par->Fhh, nti, ntj, nconj, &fij, &dfdni, &dfdnj, &dfdncn);
      else if (ktypi == REBO2_C) orc_table3d_eval(&par->Fch, ntj, nti, nconj, &fij, &dfdnj, &dfdni, &dfdncn);
      else if (ktypj == REBO2_C) orc_table3d_eval(&par->Fch, nti, ntj, nconj, &fij, &dfdni, &dfdnj, &dfdncn);

      dfdni = dfdni + dtdni * bdh;
      dfdnj = dfdnj + dtdnj * bdh;
      dfdncn = dfdncn + dtdncn * bdh;
      dfdni = 0.5 * fcarij * faij * dfdni;
      dfdnj = 0.5 * fcarij * faij * dfdnj;
      dfdncn = 0.5 * fcarij * faij * dfdncn;
      double dfdncni = 2 * dfdncn * nconji;
      double dfdncnj = 2 * dfdncn * nconjj;

      /* :2433-2517 */
      for (ikc = 0; ikc < numnbi; ikc++) {
        if (slotofi[ikc] == ij) continue; /* reference: k /= j .or. kdc /= jdc */
        int k = nebofi[ikc];
        double df[3];
        for (int c = 0; c < 3; c++) {
          df[c] = -(dfdni * (DN(dnidk, c, ikc, REBO2_C) + DN(dnidk, c, ikc, REBO2_H)) +
                    dfdncni * dncnidk[3 * ikc + c]) -
                  dfbij * (dpdnci * DN(dnidk, c, ikc, REBO2_C) + dpdnhi * DN(dnidk, c, ikc, REBO2_H));
          f[3 * k + c] += df[c];
          fi[c] = fi[c] - df[c];
        }
        outer_add(wij, -1.0, &dri[3 * ikc], df);
        for (int km = numnbk[ikc]; km < numnbk[ikc + 1]; km++) {
          int m = nebofk[km];
          if (m != i || !szero(dcofk[km])) {
            for (int c = 0; c < 3; c++) {
              df[c] = -dfdncni * dncnidm[3 * km + c];
              f[3 * m + c] += df[c];
              f[3 * k + c] += -df[c];
            }
            outer_add(wij, -1.0, &drk[3 * km], df);
          }
        }
      }
      for (jlc = 0; jlc < numnbj; jlc++) {
        int l = nebofj[jlc];
        double df[3];
        for (int c = 0; c < 3; c++) {
          df[c] = -(dfdnj * (DN(dnjdl, c, jlc, REBO2_C) + DN(dnjdl, c, jlc, REBO2_H)) +
                    dfdncnj * dncnjdl[3 * jlc + c]) -
                  dfbji * (dpdncj * DN(dnjdl, c, jlc, REBO2_C) + dpdnhj * DN(dnjdl, c, jlc, REBO2_H));
          f[3 * l + c] += df[c];
          fj[c] = fj[c] - df[c];
        }
        outer_add(wij, -1.0, &drj[3 * jlc], df);
        for (int ln = numnbl[jlc]; ln < numnbl[jlc + 1]; ln++) {
          int n = nebofl[ln];
          if (n != j || !seq(dcofl[ln], jdc)) {
            for (int c = 0; c < 3; c++) {
              df[c] = -dfdncnj * dncnjdn[3 * ln + c];
              f[3 * n + c] += df[c];
              f[3 * l + c] += -df[c];
            }
            outer_add(wij, -1.0, &drl[3 * ln], df);
          }
        }
      }

      /* :2525-2716 */
      double baveij = 0.5 * (bij + bji + fij + tij * bdh);
      double hlfvij = fcarij * (frij + baveij * faij) / 2;
      pe[i] += hlfvij;
      pe[j] += hlfvij;
      if (epot_per_bond) epot_per_bond[nbb[ij]] += 2 * hlfvij;
      double dffac = dfrijr * fcarij + baveij * dfaijr * fcarij + frij * dfcarijr + baveij * faij * dfcarijr;
      double df[3];
      for (int c = 0; c < 3; c++) {
        df[c] = dffac * rnij[c];
        fi[c] += df[c];
        fj[c] -= df[c];
      }
      outer_add(wij, 1.0, rij, df);
      for (int c = 0; c < 9; c++) wij[c] = wij[c] - dfbij * wijb[c] - dfbji * wjib[c];
      if (f_per_bond)
        for (int c = 0; c < 3; c++) f_per_bond[3 * nbb[ij] + c] += df[c];
      for (int c = 0; c < 3; c++) {
        fi[c] += -(dfbij * dbidi[c] + dfbji * dbjdi[c]);
        fj[c] += -(dfbij * dbidj[c] + dfbji * dbjdj[c]);
      }
      for (ikc = 0; ikc < numnbi; ikc++) {
        if (slotofi[ikc] == ij) continue;
        int k = nebofi[ikc];
        for (int c = 0; c < 3; c++) f[3 * k + c] += -dfbij * dbidk[3 * ikc + c];
        if (scr) {
          /* :2611-2648 forces due to screening of the bonds i-k */
          if (seedi[ikc] <= lasti[ikc]) {
            double dffac2;
            if (ktyp[k] == REBO2_C) dffac2 = dpdnci * dfbij + dfdni + dfdncni * fxik[ikc];
            else dffac2 = dpdnhi * dfbij + dfdni + dfdncni * fxik[ikc];
            for (long q = seedi[ikc]; q <= lasti[ikc]; q++) {
              sfacbo[q] = sfacbo[q] + zfaci[ikc] * dfbij;
              sfacnc[q] = sfacnc[q] + dffac2;
            }
          }
          double dffac3 = dfdncni * dnconjidxi[ikc];
          for (int km = numnbk[ikc]; km < numnbk[ikc + 1]; km++)
            if (nebofk[km] != i || !szero(dcofk[km]))
              for (long q = seedk[km]; q <= lastk[km]; q++) sfacnc[q] = sfacnc[q] + dffac3;
        }
      }
      for (jlc = 0; jlc < numnbj; jlc++) {
        int l = nebofj[jlc];
        for (int c = 0; c < 3; c++) f[3 * l + c] += -dfbji * dbjdl[3 * jlc + c];
        if (scr) {
          /* :2673-2712 */
          if (seedj[jlc] <= lastj[jlc]) {
            double dffac2;
            if (ktyp[l] == REBO2_C) dffac2 = dpdncj * dfbji + dfdnj + dfdncnj * fxjl[jlc];
            else dffac2 = dpdnhj * dfbji + dfdnj + dfdncnj * fxjl[jlc];
            for (long q = seedj[jlc]; q <= lastj[jlc]; q++) {
              sfacbo[q] = sfacbo[q] + zfacj[jlc] * dfbji;
              sfacnc[q] = sfacnc[q] + dffac2;
            }
          }
          double dffac3 = dfdncnj * dnconjjdxj[jlc];
          for (int ln = numnbl[jlc]; ln < numnbl[jlc + 1]; ln++)
            if (nebofl[ln] != j || !seq(dcofl[ln], jdc))
              for (long q = seedl[ln]; q <= lastl[ln]; q++) sfacnc[q] = sfacnc[q] + dffac3;
        }
      }
      if (scr) {
        /* :2717-2757 forces on the screening neighbours of bond i-j, attractive/repulsive part */
        double dffs = frij + baveij * faij;
        for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
          int k = sneb[q];
          double rik[3], rjk[3], d1[3];
          for (int a = 0; a < 3; a++) {
            double sh = 0.0;
            for (int c = 0; c < 3; c++) sh += M3(Abox, a, c) * (double)dc[3 * sbnd[q] + c];
            rik[a] = r[3 * k + a] - r[3 * i + a] - sh;
            rjk[a] = -rij[a] + rik[a];
          }
          for (int c = 0; c < 3; c++) {
            d1[c] = dffs * cutdrarik[q] * rik[c];
            fi[c] += d1[c];
            f[3 * k + c] += -d1[c];
          }
          outer_add(wij, 1.0, rik, d1);
          for (int c = 0; c < 3; c++) {
            d1[c] = dffs * cutdrarjk[q] * rjk[c];
            fj[c] += d1[c];
            f[3 * k + c] += -d1[c];
          }
          outer_add(wij, 1.0, rjk, d1);
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

  if (scr) {
    /* loop 3 over ALL atoms: forces due to screening via the bond-order and neighbour-count
     * cutoffs (bop_kernel_rebo2.f90:2783-2871) */
    for (int i = 0; i < nat; i++) {
      if (ktyp[i] <= 0) continue;
      double fi[3] = {0, 0, 0};
      for (long ij = neb_seed[i]; ij <= neb_last[i]; ij++) {
        int j = neb[ij];
        double fj[3] = {0, 0, 0}, wij[9] = {0};
        double rij[3] = {bndlen[ij] * bndnm[3 * ij], bndlen[ij] * bndnm[3 * ij + 1], bndlen[ij] * bndnm[3 * ij + 2]};
        for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
          cutdrboik[q] = sfacbo[q] * cutdrboik[q] + sfacnc[q] * cutdrncik[q];
          cutdrbojk[q] = sfacbo[q] * cutdrbojk[q] + sfacnc[q] * cutdrncjk[q];
        }
        for (long q = sneb_seed[ij]; q <= sneb_last[ij]; q++) {
          int k = sneb[q];
          double rik[3], rjk[3], d1[3];
          for (int a = 0; a < 3; a++) {
            double sh = 0.0;
            for (int c = 0; c < 3; c++) sh += M3(Abox, a, c) * (double)dc[3 * sbnd[q] + c];
            rik[a] = r[3 * k + a] - r[3 * i + a] - sh;
            rjk[a] = -rij[a] + rik[a];
          }
          for (int c = 0; c < 3; c++) {
            d1[c] = cutdrboik[q] * rik[c];
            fi[c] += d1[c];
            f[3 * k + c] += -d1[c];
          }
          outer_add(wij, 1.0, rik, d1);
          for (int c = 0; c < 3; c++) {
            d1[c] = cutdrbojk[q] * rjk[c];
            fj[c] += d1[c];
            f[3 * k + c] += -d1[c];
          }
          outer_add(wij, 1.0, rjk, d1);
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
  }

  double e = 0.0;
  for (int i = 0; i < nat; i++) e += pe[i];
  *epot += e;
  for (int i = 0; i < nat; i++) {
    if (epot_per_at) epot_per_at[i] += pe[i];
    for (int c = 0; c < 3; c++) f_inout[3 * i + c] += f[3 * i + c];
  }
  for (int c = 0; c < 9; c++) wpot_inout[c] += wpot[c];

  free(neb); free(nbb); free(dcell); free(bndtyp); free(bndlen); free(bndnm); free(cutfcnar);
  free(cutdrvar); free(neb_seed); free(neb_last); free(nn); free(pe); free(f);
  free(dbidk); free(dbjdl); free(dnidk); free(dnjdl); free(dnconjidxi); free(dnconjjdxj);
  free(dncnidk); free(dncnjdl); free(dncnidm); free(dncnjdn); free(xikdm); free(xjldn);
  free(fxik); free(fxjl); free(nebofi); free(nebofj); free(nebofk); free(nebofl); free(slotofi);
  free(dcofj); free(dcofk); free(dcofl); free(dcofi); free(numnbk); free(numnbl);
  free(dri); free(drj); free(drk); free(drl);
  if (scr) { free(cutfcnbo); free(cutdrvbo); free(cutfcnnc); free(cutdrvnc); }
  free(sneb_seed); free(sneb_last); free(sneb); free(sbnd); free(cutdrarik); free(cutdrarjk);
  free(cutdrboik); free(cutdrbojk); free(cutdrncik); free(cutdrncjk); free(sfacbo); free(sfacnc);
  free(seedi); free(lasti); free(seedj); free(lastj); free(seedk); free(lastk); free(seedl); free(lastl);
  free(zfaci); free(zfacj);
  return err;
}

int orc_rebo2_energy_and_forces(const orc_rebo2_params_t *par, int nat, int natloc,
                                const double *r, const double *Abox, const int *ktyp,
                                const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                const int *dc, double *epot, double *f_inout, double *wpot_inout,
                                double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                double *wpot_per_at, double *wpot_per_bond) {
  return rebo2_kernel(par, NULL, nat, natloc, r, Abox, ktyp, seed, last, neighbors, dc, epot, f_inout, wpot_inout,
                      epot_per_at, epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond);
}

int orc_rebo2_scr_energy_and_forces(const orc_rebo2_params_t *par, const orc_rebo2_scr_t *scr, int nat,
                                    int natloc, const double *r, const double *Abox, const int *ktyp,
                                    const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                    const int *dc, double *epot, double *f_inout, double *wpot_inout,
                                    double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                    double *wpot_per_at, double *wpot_per_bond) {
  if (!scr) return -4;
  return rebo2_kernel(par, scr, nat, natloc, r, Abox, ktyp, seed, last, neighbors, dc, epot, f_inout, wpot_inout,
                      epot_per_at, epot_per_bond, f_per_bond, wpot_per_at, wpot_per_bond);
}

/* test hook for the functions of rebo2_func.f90 (tests/test_func_vs_reference.py evaluates the reference's own
 * source next to it).  which: 0 fconj(x), 1 fCin(ijpot = i1, x), 2 VA(i1, x), 3 VR(i1, x),
 * 4 g(ktyp = i1, costh = x, n = y) -> val, d/dcosth, d/dN, 5 bo(ktypi = i1, zij = x, fcij = y, faij = z),
 * 6 h(ijpot = i1, ikpot = i2, dr = x), 7 Z2pair(i1, i2) -> out[0] */
void orc_rebo2_func(const orc_rebo2_params_t *p, int which, int i1, int i2, double x, double y, double z,
                    double *out) {
  out[0] = out[1] = out[2] = 0.0;
  switch (which) {
    case 0: fconj(x, &out[0], &out[1]); break;
    case 1: fCin(p, i1, x, &out[0], &out[1]); break;
    case 2: VA(p, i1, x, &out[0], &out[1]); break;
    case 3: VR(p, i1, x, &out[0], &out[1]); break;
    case 4: gfun(p, i1, x, y, &out[0], &out[1], &out[2]); break;
    case 5: bo(p, i1, x, y, z, &out[0], &out[1]); break;
    case 6: hfun(p, i1, i2, x, &out[0], &out[1]); break;
    default: out[0] = (double)Z2pair(i1, i2);
  }
}
