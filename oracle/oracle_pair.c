/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see atomistica_oracle.h).
 *
 * Pair potentials, restated loop by loop (each has its own traversal and weighting conventions):
 *   LJCut           src/potentials/pair_potentials/lj_cut.f90:190-330   (i <= j, mask weights)
 *   Harmonic        src/potentials/pair_potentials/harmonic.f90:150-225 (i > j, no mask)
 *   DoubleHarmonic  src/potentials/pair_potentials/double_harmonic.f90:140-230 (every directed
 *                   entry with a factor 1/2; epot_per_at receives en/2 per entry and atom, i.e. the
 *                   per-atom energies sum to twice epot -- kept as in the reference)
 *   r6              src/potentials/pair_potentials/r6.f90:150-215 (as Harmonic: i > j, no mask)
 *   BornMayer       src/potentials/pair_potentials/born_mayer.f90:200-275 (i <= j INCLUDING the
 *                   i == j image entries, asymmetric element test, energy and forces only:
 *                   wpot and the per-atom outputs are never touched)
 * Element filters are the bit masks of src/core/filter.f90 (bit k = particle element id k).
 * dr follows DIST_SQ / GET_DRJ (macros.inc:76): r_i - r_j + Abox.dc.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_oracle.h"

#define M3(M, i, j) (M)[(j) * 3 + (i)]
#define IS_EL2(f, e) (((f) >> (e)) & 1)

static void dist(const double *r, const double *Abox, const int *dc, int i, int j, intptr_t jn,
                 double *dr, double *d2) {
  *d2 = 0.0;
  for (int k = 0; k < 3; k++) {
    double s = 0.0;
    for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)dc[3 * (jn - 1) + c];
    dr[k] = r[3 * i + k] - r[3 * j + k] + s;
    *d2 += dr[k] * dr[k];
  }
}

static void add_virial(double *w, double *wpa, int i, int j, const double *dr, const double *df) {
  double dw[9];
  for (int b = 0; b < 3; b++)
    for (int a = 0; a < 3; a++) {
      M3(dw, a, b) = -(dr[a] * df[b]);
      M3(w, a, b) += M3(dw, a, b);
    }
  if (wpa)
    for (int c = 0; c < 9; c++) {
      wpa[9 * i + c] += dw[c] / 2;
      wpa[9 * j + c] += dw[c] / 2;
    }
}

int orc_pair_energy_and_forces(int kind, const double *par, int shift, int nat, const double *r,
                               const double *Abox, const int *el, int el1, int el2,
                               const intptr_t *seed, const intptr_t *last, const int *neighbors,
                               const int *dc, const int *mask, double *epot, double *f, double *wpot,
                               double *epot_per_at, double *wpot_per_at) {
  double w[9] = {0};
  if (kind == ORC_PAIR_LJCUT) {
    const double epsilon = par[0], sigma = par[1], cutoff = par[2];
    const double cut_sq = cutoff * cutoff;
    double offset = 0.0;
    if (shift) offset = 4 * epsilon * (pow(sigma / cutoff, 12) - pow(sigma / cutoff, 6));
    double *sca = (double *)calloc(nat > 0 ? nat : 1, sizeof(double));
    double *vec = (double *)calloc(nat > 0 ? 3 * nat : 1, sizeof(double));
    for (int i = 0; i < nat; i++) {
      int weighti = 1;
      if (mask && mask[i] == 0) weighti = 0;
      for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
        int j = neighbors[jn - 1] - 1;
        if (!(i <= j)) continue;
        int maskj = mask && mask[j] == 0;
        int weight = (i == j || maskj) ? weighti : weighti + 1;
        if (!(weight > 0 && ((IS_EL2(el1, el[i]) && IS_EL2(el2, el[j])) ||
                             (IS_EL2(el2, el[i]) && IS_EL2(el1, el[j])))))
          continue;
        double dr[3], abs_dr;
        dist(r, Abox, dc, i, j, jn, dr, &abs_dr);
        if (abs_dr < cut_sq) {
          abs_dr = sqrt(abs_dr);
          double fac12 = pow(sigma / abs_dr, 12), fac6 = pow(sigma / abs_dr, 6);
          double en = 0.5 * weight * (4 * epsilon * (fac12 - fac6) - offset);
          double fo = 0.5 * weight * 24 * epsilon * (2 * fac12 - fac6) / abs_dr;
          double df[3];
          for (int k = 0; k < 3; k++) {
            df[k] = fo * dr[k] / abs_dr;
            vec[3 * i + k] += df[k];
            vec[3 * j + k] -= df[k];
          }
          en = en / 2;
          sca[i] += en;
          sca[j] += en;
          add_virial(w, wpot_per_at, i, j, dr, df);
        }
      }
    }
    double e = 0.0;
    for (int i = 0; i < nat; i++) {
      e += sca[i];
      if (epot_per_at) epot_per_at[i] += sca[i];
      for (int k = 0; k < 3; k++) f[3 * i + k] += vec[3 * i + k];
    }
    *epot += e;
    free(sca); free(vec);
  } else if (kind == ORC_PAIR_HARMONIC) {
    if (mask) return -1;
    const double kk = par[0], r0 = par[1], cutoff = par[2];
    const double cut_sq = cutoff * cutoff;
    double offset = shift ? 0.5 * kk * (cutoff - r0) * (cutoff - r0) : 0.0;
    for (int i = 0; i < nat; i++)
      for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
        int j = neighbors[jn - 1] - 1;
        if (!(i > j)) continue;
        if (!((IS_EL2(el1, el[i]) && IS_EL2(el2, el[j])) || (IS_EL2(el2, el[i]) && IS_EL2(el1, el[j])))) continue;
        double dr[3], abs_dr;
        dist(r, Abox, dc, i, j, jn, dr, &abs_dr);
        if (abs_dr < cut_sq) {
          abs_dr = sqrt(abs_dr);
          double fo = kk * (r0 - abs_dr);
          double en = 0.5 * fo * (r0 - abs_dr) - offset;
          *epot += en;
          double df[3];
          for (int k = 0; k < 3; k++) {
            df[k] = fo * dr[k] / abs_dr;
            f[3 * i + k] += df[k];
            f[3 * j + k] -= df[k];
          }
          if (epot_per_at) {
            epot_per_at[i] += en / 2;
            epot_per_at[j] += en / 2;
          }
          add_virial(w, wpot_per_at, i, j, dr, df);
        }
      }
  } else if (kind == ORC_PAIR_DOUBLE_HARMONIC) {
    if (mask) return -1;
    const double k1 = par[0], r1 = par[1], k2 = par[2], r2 = par[3], cutoff = par[4];
    const double cut_sq = cutoff * cutoff, rm = (r1 + r2) / 2;
    for (int i = 0; i < nat; i++)
      for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
        int j = neighbors[jn - 1] - 1;
        if (!((IS_EL2(el1, el[i]) && IS_EL2(el2, el[j])) || (IS_EL2(el2, el[i]) && IS_EL2(el1, el[j])))) continue;
        double dr[3], abs_dr;
        dist(r, Abox, dc, i, j, jn, dr, &abs_dr);
        if (abs_dr < cut_sq) {
          abs_dr = sqrt(abs_dr);
          double fo, en;
          if (abs_dr < rm) { fo = k1 * (r1 - abs_dr); en = 0.5 * fo * (r1 - abs_dr); }
          else { fo = k2 * (r2 - abs_dr); en = 0.5 * fo * (r2 - abs_dr); }
          *epot += 0.5 * en;
          double df[3];
          for (int k = 0; k < 3; k++) {
            df[k] = 0.5 * fo * dr[k] / abs_dr;
            f[3 * i + k] += df[k];
            f[3 * j + k] -= df[k];
          }
          if (epot_per_at) {
            epot_per_at[i] += en / 2;
            epot_per_at[j] += en / 2;
          }
          add_virial(w, wpot_per_at, i, j, dr, df);
        }
      }
  } else if (kind == ORC_PAIR_R6) {
    if (mask) return -1;
    const double A = par[0], r0 = par[1], cutoff = par[2];
    const double cut_sq = cutoff * cutoff;
    for (int i = 0; i < nat; i++)
      for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
        int j = neighbors[jn - 1] - 1;
        if (!(i > j)) continue;
        if (!((IS_EL2(el1, el[i]) && IS_EL2(el2, el[j])) || (IS_EL2(el2, el[i]) && IS_EL2(el1, el[j])))) continue;
        double dr[3], abs_dr;
        dist(r, Abox, dc, i, j, jn, dr, &abs_dr);
        if (abs_dr < cut_sq) {
          abs_dr = sqrt(abs_dr);
          double en = A / pow(r0 + abs_dr, 6);
          double fo = 6 * en / (r0 + abs_dr);
          *epot += en;
          double df[3];
          for (int k = 0; k < 3; k++) {
            df[k] = fo * dr[k] / abs_dr;
            f[3 * i + k] += df[k];
            f[3 * j + k] -= df[k];
          }
          if (epot_per_at) {
            epot_per_at[i] += en / 2;
            epot_per_at[j] += en / 2;
          }
          add_virial(w, wpot_per_at, i, j, dr, df);
        }
      }
  } else if (kind == ORC_PAIR_BORN_MAYER) {
    if (mask) return -1;
    const double A = par[0], rho = par[1], cutoff = par[2];
    const double shift_e = A * exp(-cutoff / rho); /* always shifted, born_mayer.f90:169 */
    double e = 0.0;
    for (int i = 0; i < nat; i++) {
      int want; /* filter the partner must match */
      if (IS_EL2(el1, el[i])) want = el2;
      else if (IS_EL2(el2, el[i])) want = el1;
      else continue;
      for (intptr_t jn = seed[i]; jn <= last[i]; jn++) {
        int j = neighbors[jn - 1] - 1;
        double dr[3], abs_dr;
        dist(r, Abox, dc, i, j, jn, dr, &abs_dr);
        if (i <= j && IS_EL2(want, el[j])) {
          if (abs_dr < cutoff * cutoff) {
            abs_dr = sqrt(abs_dr);
            double exp_r = exp(-abs_dr / rho);
            e = e + A * exp_r - shift_e;
            for (int k = 0; k < 3; k++) {
              double fk = (A / rho) * exp_r * dr[k] / abs_dr;
              f[3 * i + k] += fk;
              f[3 * j + k] -= fk;
            }
          }
        }
      }
    }
    *epot += e;
  } else {
    return -2;
  }
  for (int c = 0; c < 9; c++) wpot[c] += w[c];
  return 0;
}
