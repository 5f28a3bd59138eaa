// REBO2 device functions shared by the unscreened (atx_rebo2.cu) and screened (atx_rebo2_scr.cuh)
// kernels: parameters as the device sees them and the functions of rebo2_func.f90 / table2d.f90 /
// table3d.f90 (citations in atx_rebo2.cu).  Pure per-call math, no CUDA runtime types: the CPU
// emulation harness of the test-suite (tests/emu/) compiles this header with a host compiler.
#pragma once

#include "atomistica_b200.h"

#define RB_PI 3.14159265358979323846264338327950288
#define RB_C 1
#define RB_H 3
#define RB_CC 1
#define RB_CH 3
#define RB_HH 6
#define RB_NBL 12  // per-thread bond scratch; larger coordination raises the reference's nebmax error

struct Rebo2Dev {
  double cc_B1, cc_B2, cc_B3, cc_beta1, cc_beta2, cc_beta3, cc_Q, cc_A, cc_alpha;
  double ch_B1, ch_beta1, ch_Q, ch_A, ch_alpha;
  double hh_B1, hh_beta1, hh_Q, hh_A, hh_alpha;
  double cc_g_theta[6], g1c[18], g2c[18], spgh[18];
  int igh[25];
  double conalp, conear[36], conpe[3], conan[3], conpf[3];
  double cut_l[7], cut_h[7], cut_h2[7], cut_fac[7];
  int with_dihedral;
  int el2typ[32];
  const double *Fcc, *Fch, *Fhh, *Tcc, *Pcc, *Pch;
  double n37;  // (double)3.7f, single-precision literal of rebo2_func.f90:314
};

// host: parameters of the C ABI -> device view (tables and el2typ are set by the caller)
inline void rb_fill_dev(Rebo2Dev &D, const atx_rebo2_params *par) {
  D.cc_B1 = par->cc_B1; D.cc_B2 = par->cc_B2; D.cc_B3 = par->cc_B3;
  D.cc_beta1 = par->cc_beta1; D.cc_beta2 = par->cc_beta2; D.cc_beta3 = par->cc_beta3;
  D.cc_Q = par->cc_Q; D.cc_A = par->cc_A; D.cc_alpha = par->cc_alpha;
  D.ch_B1 = par->ch_B1; D.ch_beta1 = par->ch_beta1; D.ch_Q = par->ch_Q; D.ch_A = par->ch_A;
  D.ch_alpha = par->ch_alpha;
  D.hh_B1 = par->hh_B1; D.hh_beta1 = par->hh_beta1; D.hh_Q = par->hh_Q; D.hh_A = par->hh_A;
  D.hh_alpha = par->hh_alpha;
  for (int i = 0; i < 6; i++) D.cc_g_theta[i] = par->cc_g_theta[i];
  for (int i = 0; i < 18; i++) {
    D.g1c[i] = par->cc_g1_coeff[i];
    D.g2c[i] = par->cc_g2_coeff[i];
    D.spgh[i] = par->spgh[i];
  }
  for (int i = 0; i < 25; i++) D.igh[i] = par->igh[i];
  D.conalp = par->conalp;
  for (int i = 0; i < 36; i++) D.conear[i] = par->conear[i];
  for (int i = 0; i < 3; i++) {
    D.conpe[i] = par->conpe[i];
    D.conan[i] = par->conan[i];
    D.conpf[i] = par->conpf[i];
  }
  for (int i = 0; i < 7; i++) D.cut_l[i] = D.cut_h[i] = D.cut_h2[i] = D.cut_fac[i] = 0.0;
  for (int t : {RB_CC, RB_CH, RB_HH}) {
    D.cut_l[t] = par->cut_in_l[t - 1];
    D.cut_h[t] = par->cut_in_h[t - 1];
    D.cut_h2[t] = par->cut_in_h2[t - 1];
    D.cut_fac[t] = RB_PI / (D.cut_h[t] - D.cut_l[t]);
  }
  D.with_dihedral = par->with_dihedral;
  D.n37 = (double)3.7f;
}

__device__ __forceinline__ void rb_table2d(const double *__restrict__ coeff, int nx, int ny,
                                           double nhi, double nci, double &v, double &dvdh, double &dvdc) {
  const int nboxs = nx * ny;
  int nhbox = (int)nhi;
  if (nhbox < 0) nhbox = 0;
  if (nhbox >= nx) nhbox = nx - 1;
  int ncbox = (int)nci;
  if (ncbox < 0) ncbox = 0;
  if (ncbox >= ny) ncbox = ny - 1;
  const int ibox = ny * nhbox + ncbox;
  const double x1 = nhi - nhbox, x2 = nci - ncbox;
  v = 0.0; dvdh = 0.0; dvdc = 0.0;
  for (int i = 4; i >= 1; i--) {
    double s = 0.0, sdc = 0.0;
    for (int j = 4; j >= 1; j--) {
      double c = __ldg(&coeff[ibox + nboxs * ((i - 1) + 4 * (j - 1))]);
      s = s * x2 + c;
      if (j > 1) sdc = sdc * x2 + (j - 1) * c;
    }
    v = v * x1 + s;
    if (i > 1) dvdh = dvdh * x1 + (i - 1) * s;
    dvdc = dvdc * x1 + sdc;
  }
}

__device__ __forceinline__ void rb_table3d(const double *__restrict__ coeff, int nx, int ny, int nz,
                                           double nti, double ntj, double nc, double &v, double &dvdi,
                                           double &dvdj, double &dvdc) {
  const int nboxs = nx * ny * nz;
  int ib = (int)nti;
  if (ib < 0) ib = 0;
  if (ib >= nx) ib = nx - 1;
  int jb = (int)ntj;
  if (jb < 0) jb = 0;
  if (jb >= ny) jb = ny - 1;
  int cb = (int)nc;
  if (cb < 0) cb = 0;
  if (cb >= nz) cb = nz - 1;
  const int ibox = nx * (ny * cb + jb) + ib;
  const double x1 = nti - ib, x2 = ntj - jb, x3 = nc - cb;
  v = 0.0; dvdi = 0.0; dvdj = 0.0; dvdc = 0.0;
  for (int i = 4; i >= 1; i--) {
    double s = 0.0, sdj = 0.0, sdc = 0.0;
    for (int j = 4; j >= 1; j--) {
      double t = 0.0, tdc = 0.0;
      for (int k = 4; k >= 1; k--) {
        double c = __ldg(&coeff[ibox + nboxs * ((i - 1) + 4 * ((j - 1) + 4 * (k - 1)))]);
        t = t * x3 + c;
        if (k > 1) tdc = tdc * x3 + (k - 1) * c;
      }
      s = s * x2 + t;
      if (j > 1) sdj = sdj * x2 + (j - 1) * t;
      sdc = sdc * x2 + tdc;
    }
    v = v * x1 + s;
    if (i > 1) dvdi = dvdi * x1 + (i - 1) * s;
    dvdj = dvdj * x1 + sdj;
    dvdc = dvdc * x1 + sdc;
  }
}

__device__ __forceinline__ void rb_fconj(double x, double &fx, double &dfx) {
  if (x <= 2.0) { fx = 1.0; dfx = 0.0; }
  else if (x >= 3.0) { fx = 0.0; dfx = 0.0; }
  else {
    double sn, cs;
    sincos(RB_PI * (x - 2.0), &sn, &cs);
    fx = 0.5 * (1.0 + cs);
    dfx = -0.5 * RB_PI * sn;
  }
}

__device__ __forceinline__ void rb_VA(const Rebo2Dev &P, int ijpot, double dr, double &val, double &dval) {
  if (ijpot == RB_CC) {
    double e1 = P.cc_B1 * exp(-P.cc_beta1 * dr);
    double e2 = P.cc_B2 * exp(-P.cc_beta2 * dr);
    double e3 = P.cc_B3 * exp(-P.cc_beta3 * dr);
    val = -(e1 + e2 + e3);
    dval = -(-P.cc_beta1 * e1 - P.cc_beta2 * e2 - P.cc_beta3 * e3);
  } else if (ijpot == RB_CH) {
    double e1 = P.ch_B1 * exp(-P.ch_beta1 * dr);
    val = -e1;
    dval = P.ch_beta1 * e1;
  } else {
    double e1 = P.hh_B1 * exp(-P.hh_beta1 * dr);
    val = -e1;
    dval = P.hh_beta1 * e1;
  }
}

__device__ __forceinline__ void rb_VR(const Rebo2Dev &P, int ijpot, double dr, double &val, double &dval) {
  double A, Q, al;
  if (ijpot == RB_CC) { A = P.cc_A; Q = P.cc_Q; al = P.cc_alpha; }
  else if (ijpot == RB_CH) { A = P.ch_A; Q = P.ch_Q; al = P.ch_alpha; }
  else { A = P.hh_A; Q = P.hh_Q; al = P.hh_alpha; }
  double e1 = A * exp(-al * dr);
  double hlp1 = 1 + Q / dr;
  val = hlp1 * e1;
  dval = (-Q / (dr * dr) - hlp1 * al) * e1;
}

__device__ __forceinline__ void rb_poly5(const double *c, double x, double &h, double &dh) {
  // h = c1 + c2 x + sum_{i=3..6} c_i x^(i-1)
  double x2 = x * x, x3 = x2 * x, x4 = x3 * x, x5 = x4 * x;
  h = c[0] + c[1] * x;
  dh = c[1];
  h = h + c[2] * x2; dh = dh + 2 * c[2] * x;
  h = h + c[3] * x3; dh = dh + 3 * c[3] * x2;
  h = h + c[4] * x4; dh = dh + 4 * c[4] * x3;
  h = h + c[5] * x5; dh = dh + 5 * c[5] * x4;
}

__device__ __forceinline__ void rb_cc_g(const Rebo2Dev &P, const double *c, double costh, double &val,
                                        double &dval) {
  int j;
  if (costh < P.cc_g_theta[1]) j = 0;
  else if (costh < P.cc_g_theta[2]) j = 1;
  else j = 2;
  rb_poly5(&c[6 * j], costh, val, dval);
}

__device__ __forceinline__ void rb_g(const Rebo2Dev &P, int ktyp, double costh, double n, double &val,
                                     double &dval, double &dvaldN) {
  dvaldN = 0.0;
  if (ktyp == RB_C) {
    if (n < 3.2) rb_cc_g(P, P.g2c, costh, val, dval);
    else if (n > P.n37) rb_cc_g(P, P.g1c, costh, val, dval);
    else {
      double v1, v2, dv1, dv2;
      rb_cc_g(P, P.g1c, costh, v1, dv1);
      rb_cc_g(P, P.g2c, costh, v2, dv2);
      double sn, cs;
      sincos(2 * RB_PI * (n - 3.2), &sn, &cs);
      double s = (1 + cs) / 2, ds = -RB_PI * sn;
      val = v1 * (1 - s) + v2 * s;
      dval = dv1 * (1 - s) + dv2 * s;
      dvaldN = (v2 - v1) * ds;
    }
  } else {
    int ig = P.igh[(int)(-costh * 12.0) + 13 - 1];
    rb_poly5(&P.spgh[6 * (ig - 1)], costh, val, dval);
  }
}

__device__ __forceinline__ void rb_bo(const Rebo2Dev &P, int ktypi, double zij, double fcij, double faij,
                                      double &bij, double &dfbij) {
  double arg = 1.0 + zij;
  bij = pow(arg, P.conpe[ktypi - 1]);
  dfbij = P.conan[ktypi - 1] * fcij * faij * pow(arg, P.conpf[ktypi - 1]);
}

__device__ __forceinline__ void rb_h(const Rebo2Dev &P, int ijpot, int ikpot, double dr, double &val,
                                     double &dval) {
  if (ijpot + ikpot <= 4) { val = 1.0; dval = 0.0; }
  else {
    val = P.conear[(ijpot - 1) + 6 * (ikpot - 1)] * exp(P.conalp * dr);
    dval = P.conalp * val;
  }
}

__device__ __forceinline__ int rb_Z2pair(int a, int b) {
  if (a == RB_C) return b;
  if (b == RB_C) return a;
  return a + b;
}

