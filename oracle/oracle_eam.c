/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see atomistica_oracle.h).
 *
 * Tabulated alloy EAM, restated from
 *   src/potentials/eam/tabulated_alloy_eam.f90:423-627 (energy_and_forces_kernel)
 *   src/support/simple_spline.f90:373-434 (func), 467-528 (dfunc), 536-614 (f_and_df)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_oracle.h"

#define M3(M, i, j) (M)[(j) * 3 + (i)]

/* simple_spline_f; returns nonzero when x is outside the table (RAISE_ERROR in the reference) */
static int spl_func(const orc_spline_t *s, double x, double *f) {
  double xf;
  int i;
  if (x == s->cut) {
    xf = s->n;
    i = s->n - 1;
  } else {
    xf = (x - s->x0) / s->dx + 1;
    i = (int)floor(xf);
  }
  if (i < 1 || i >= s->n) { *f = 1.0; return 1; }
  double B = xf - i;
  i--;
  *f = s->y[i] + B * (s->c1[i] + B * (s->c2[i] + B * s->c3[i]));
  return 0;
}

static int spl_dfunc(const orc_spline_t *s, double x, double *df) {
  double xf;
  int i;
  if (x == s->cut) {
    xf = s->n;
    i = s->n - 1;
  } else {
    xf = (x - s->x0) / s->dx + 1;
    i = (int)floor(xf);
  }
  if (i < 1 || i >= s->n) { *df = 1.0; return 1; }
  double B = xf - i;
  i--;
  *df = s->d1[i] + B * (s->d2[i] + B * s->d3[i]);
  return 0;
}

/* simple_spline_f_and_df with extrapolate=.true. */
static void spl_f_and_df_extrap(const orc_spline_t *s, double x, double *f, double *df) {
  double xf = (x - s->x0) / s->dx + 1;
  int i = (int)floor(xf);
  if (i < 1) i = 1;
  else if (i >= s->n) i = s->n - 1;
  double B = xf - i;
  i--;
  *f = s->y[i] + B * (s->c1[i] + B * (s->c2[i] + B * s->c3[i]));
  *df = s->d1[i] + B * (s->d2[i] + B * s->d3[i]);
}

static int orc_eam_threads = 1;

/* number of OpenMP threads of the EAM kernel (the reference runs it under "!$omp parallel",
 * tabulated_alloy_eam.f90:473-486, with thread-local energy/force arrays tls_sca1/tls_vec1).
 * 1 (default) keeps the summation order of the serial reference build used for the KATs. */
void orc_eam_set_threads(int n) { orc_eam_threads = n > 0 ? n : 1; }

int orc_eam_energy_and_forces(int nat, int natloc, const double *r, const double *Abox,
                              const int *eldb, const intptr_t *seed, const intptr_t *last,
                              const int *neighbors, const int *dc, int ndb, const orc_spline_t *fF,
                              const orc_spline_t *frho, const orc_spline_t *fphi, double cutoff,
                              const int *mask, double *epot, double *f, double *wpot,
                              double *epot_per_at, double *wpot_per_at) {
  double cutoff_sq = cutoff * cutoff;
  double e = 0.0, w[9] = {0};
  int maxneb = 0;
  for (int i = 0; i < natloc; i++) {
    int d = (int)(last[i] - seed[i] + 1);
    if (d > maxneb) maxneb = d;
  }
  double *pe = (double *)calloc(nat > 0 ? nat : 1, sizeof(double));      /* tls_sca1 (summed) */
  double *fv = (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double));  /* tls_vec1 (summed) */
  int err = 0;
  const int nthreads = orc_eam_threads;

#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
  {
    int *neb = (int *)malloc(sizeof(int) * (maxneb + 1));
    double *neb_dr = (double *)malloc(sizeof(double) * 3 * (maxneb + 1));
    double *neb_abs = (double *)malloc(sizeof(double) * (maxneb + 1));
    /* thread-local force array; the serial path writes into fv directly */
    double *fl = nthreads > 1 ? (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double)) : fv;
    double wl[9] = {0};
    int errl = 0;

#pragma omp for schedule(static)
    for (int i = 0; i < natloc; i++) {
      if (errl) continue;
      if (mask && mask[i] == 0) continue;
      int dbi = eldb[i];
      if (dbi <= 0) continue;
      double rho = 0.0;
      int neb_n = 0;
      for (intptr_t ni = seed[i]; ni <= last[i]; ni++) {
        int j = neighbors[ni - 1] - 1;
        int dbj = eldb[j];
        if (dbj <= 0) continue;
        double dr[3];
        for (int k = 0; k < 3; k++) {
          double s = 0.0;
          for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (ni - 1) + c];
          dr[k] = r[3 * i + k] - r[3 * j + k] + s;
        }
        double abs_dr = 0.0;
        for (int k = 0; k < 3; k++) abs_dr += dr[k] * dr[k];
        if (abs_dr < cutoff_sq) {
          abs_dr = sqrt(abs_dr);
          double drho;
          if (spl_func(&frho[dbj - 1], abs_dr, &drho)) { errl = 1; break; }
          rho += drho;
          neb[neb_n] = j;
          neb_dr[3 * neb_n + 0] = dr[0];
          neb_dr[3 * neb_n + 1] = dr[1];
          neb_dr[3 * neb_n + 2] = dr[2];
          neb_abs[neb_n] = abs_dr;
          neb_n++;
        }
      }
      if (errl) continue;
      if (rho < 0.0) rho = 0.0;
      double Fi, dFi;
      spl_f_and_df_extrap(&fF[dbi - 1], rho, &Fi, &dFi);
      double pei = Fi;   /* pe(i) only ever receives contributions from its own centre */

      double fori[3] = {0, 0, 0};
      for (int ni = 0; ni < neb_n; ni++) {
        int j = neb[ni];
        int dbj = eldb[j];
        const double *dr = &neb_dr[3 * ni];
        double abs_dr = neb_abs[ni];
        double phi, dphi, fac;
        spl_f_and_df_extrap(&fphi[(dbi - 1) + ndb * (dbj - 1)], abs_dr, &phi, &dphi);
        double r_abs_dr = 1.0 / abs_dr;
        pei += phi * r_abs_dr;
        if (spl_dfunc(&frho[dbj - 1], abs_dr, &fac)) { errl = 1; break; }
        double pref = -(dFi * fac + (dphi - phi * r_abs_dr) * r_abs_dr) * r_abs_dr;
        double df[3] = {pref * dr[0], pref * dr[1], pref * dr[2]};
        double wij[9];
        for (int k = 0; k < 3; k++) {
          fori[k] += df[k];
          fl[3 * j + k] -= df[k];
        }
        for (int b = 0; b < 3; b++)
          for (int a = 0; a < 3; a++) {
            M3(wij, a, b) = -(dr[a] * df[b]);
            M3(wl, a, b) += M3(wij, a, b);
          }
        if (wpot_per_at)
          for (int k = 0; k < 9; k++) {
            if (nthreads > 1) {
#pragma omp atomic
              wpot_per_at[9 * i + k] += wij[k] / 2;
#pragma omp atomic
              wpot_per_at[9 * j + k] += wij[k] / 2;
            } else {
              wpot_per_at[9 * i + k] += wij[k] / 2;
              wpot_per_at[9 * j + k] += wij[k] / 2;
            }
          }
      }
      pe[i] += pei;
      for (int k = 0; k < 3; k++) fl[3 * i + k] += fori[k];
    }

#pragma omp critical
    {
      if (errl) err = 1;
      for (int k = 0; k < 9; k++) w[k] += wl[k];
      if (nthreads > 1)
        for (int i = 0; i < 3 * nat; i++) fv[i] += fl[i];
    }
    free(neb); free(neb_dr); free(neb_abs);
    if (nthreads > 1) free(fl);
  }

  if (!err) {
    for (int i = 0; i < natloc; i++) e += pe[i];
    for (int i = 0; i < nat; i++) {
      if (epot_per_at) epot_per_at[i] += pe[i];
      for (int k = 0; k < 3; k++) f[3 * i + k] += fv[3 * i + k];
    }
    *epot += e;
    for (int k = 0; k < 9; k++) wpot[k] += w[k];
  }
  free(pe); free(fv);
  return err ? -1 : 0;
}

/* ======================================================================================
 * TabulatedEAM (single-element funcfl tables), restated from
 *   src/potentials/eam/tabulated_eam.f90:334-511 (energy_and_forces_kernel)
 *   src/spline.inc (SPLINE_FUNC / SPLINE_DFUNC / SPLINE_F_AND_DF: x is mapped with the
 *   RECIPROCAL spacing, the index is the truncated value, no range check)
 * phi = Z(r)**2 / r with Z scaled by sqrt(0.5 Hartree Bohr) at init (tabulated_eam.f90:199).
 * `in` flags the atoms selected by the element filter (IS_EL2(els, el)).
 * ====================================================================================== */

static void inl_f(const orc_spline_t *s, double rdx, double x, double *f) {
  double xf = (x - s->x0) * rdx + 1.0;
  int i = (int)xf;
  double B = xf - i;
  i--;
  *f = s->y[i] + B * (s->c1[i] + B * (s->c2[i] + B * s->c3[i]));
}

static void inl_df(const orc_spline_t *s, double rdx, double x, double *df) {
  double xf = (x - s->x0) * rdx + 1.0;
  int i = (int)xf;
  double B = xf - i;
  i--;
  *df = s->d1[i] + B * (s->d2[i] + B * s->d3[i]);
}

int orc_eam_funcfl_energy_and_forces(int nat, const double *r, const double *Abox, const int *in,
                                     const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                     const int *dc, const orc_spline_t *fF, const orc_spline_t *fZ,
                                     const orc_spline_t *frho, double cutoff, double *epot, double *f,
                                     double *wpot, double *epot_per_at) {
  const double cutoff_sq = cutoff * cutoff;
  const double F_rdx = 1.0 / fF->dx, Z_rdx = 1.0 / fZ->dx, rho_rdx = 1.0 / frho->dx;
  int maxneb = 0;
  for (int i = 0; i < nat; i++) {
    int d = (int)(last[i] - seed[i] + 1);
    if (d > maxneb) maxneb = d;
  }
  int *neb = (int *)malloc(sizeof(int) * (maxneb + 1));
  double *neb_dr = (double *)malloc(sizeof(double) * 3 * (maxneb + 1));
  double *neb_abs = (double *)malloc(sizeof(double) * (maxneb + 1));
  double *sca = (double *)calloc(nat > 0 ? nat : 1, sizeof(double));
  double *vec = (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double));
  double e = 0.0, w[9] = {0};
  for (int i = 0; i < nat; i++) {
    if (!in[i]) continue;
    double rho = 0.0;
    int neb_n = 0;
    for (intptr_t ni = seed[i]; ni <= last[i]; ni++) {
      int j = neighbors[ni - 1] - 1;
      if (!in[j]) continue;
      double dr[3], abs_dr = 0.0;
      for (int k = 0; k < 3; k++) {
        double s = 0.0;
        for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (ni - 1) + c];
        dr[k] = r[3 * i + k] - r[3 * j + k] + s;
        abs_dr += dr[k] * dr[k];
      }
      if (abs_dr < cutoff_sq) {
        abs_dr = sqrt(abs_dr);
        double drho;
        inl_f(frho, rho_rdx, abs_dr, &drho);
        rho = rho + drho;
        neb[neb_n] = j;
        neb_dr[3 * neb_n] = dr[0]; neb_dr[3 * neb_n + 1] = dr[1]; neb_dr[3 * neb_n + 2] = dr[2];
        neb_abs[neb_n] = abs_dr;
        neb_n++;
      }
    }
    if (rho < 0.0) rho = 0.0;
    double Fi, dFi;
    inl_f(fF, F_rdx, rho, &Fi);
    inl_df(fF, F_rdx, rho, &dFi);
    sca[i] += Fi;
    double sumphi = 0.0, fori[3] = {0, 0, 0};
    for (int n = 0; n < neb_n; n++) {
      double a = neb_abs[n], z, dz, fac;
      inl_f(fZ, Z_rdx, a, &z);
      inl_df(fZ, Z_rdx, a, &dz);
      double dphi = 2 * z * dz / a;
      double phi = z * z / a;
      dphi = dphi - phi / a;
      sumphi += phi;
      inl_df(frho, rho_rdx, a, &fac);
      double pref = -((dFi * fac + dphi) / a);
      double df[3] = {pref * neb_dr[3 * n], pref * neb_dr[3 * n + 1], pref * neb_dr[3 * n + 2]};
      int j = neb[n];
      for (int k = 0; k < 3; k++) {
        fori[k] += df[k];
        vec[3 * j + k] -= df[k];
      }
      for (int b = 0; b < 3; b++)
        for (int aa = 0; aa < 3; aa++) M3(w, aa, b) += -(neb_dr[3 * n + aa] * df[b]);
    }
    sca[i] += sumphi;
    for (int k = 0; k < 3; k++) vec[3 * i + k] += fori[k];
  }
  for (int i = 0; i < nat; i++) {
    e += sca[i];
    if (epot_per_at) epot_per_at[i] += sca[i];
    for (int k = 0; k < 3; k++) f[3 * i + k] += vec[3 * i + k];
  }
  *epot += e;
  for (int k = 0; k < 9; k++) wpot[k] += w[k];
  free(neb); free(neb_dr); free(neb_abs); free(sca); free(vec);
  return 0;
}
