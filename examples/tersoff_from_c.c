/* Using the C ABI from plain C: Tersoff silicon on a diamond lattice.
 *
 *   gcc -std=c99 -Iinclude examples/tersoff_from_c.c -o tersoff_from_c \
 *       -Latomistica_b200 -latomistica_b200 -Wl,-rpath,$PWD/atomistica_b200 -lm
 *
 * This is the call sequence the Fortran shims of atomistica_b200/fortran/ make on behalf of
 * tersoff_bind_to / tersoff_energy_and_forces.  Prints the cohesive
 * energy per atom (-4.62959501 eV for the Tersoff 1989 Si parameters at a0 = 5.432 A) and the
 * largest force component of the rattled crystal.  Without a CUDA device atx_ctx_create fails and
 * the library's error text is printed: there is no CPU fallback.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_b200.h"

#define CHECK(call)                                           \
  do {                                                        \
    int err_ = (call);                                        \
    if (err_ != 0) {                                          \
      char msg[1024];                                         \
      atx_last_error(msg, (int)sizeof(msg));                  \
      fprintf(stderr, "%s failed (%d): %s\n", #call, err_, msg); \
      return 2;                                               \
    }                                                         \
  } while (0)

int main(void) {
  /* diamond Si, 3 x 3 x 3 conventional cells */
  const double a0 = 5.432;
  const int nc = 3, nat = 8 * nc * nc * nc;
  static const double basis[8][3] = {{0, 0, 0}, {0, .5, .5}, {.5, 0, .5}, {.5, .5, 0},
                                     {.25, .25, .25}, {.25, .75, .75}, {.75, .25, .75}, {.75, .75, .25}};
  double *r = (double *)malloc(sizeof(double) * 3 * nat);
  double *f = (double *)calloc(3 * (size_t)nat, sizeof(double));
  int *el = (int *)malloc(sizeof(int) * nat);
  int n = 0;
  for (int i = 0; i < nc; i++)
    for (int j = 0; j < nc; j++)
      for (int k = 0; k < nc; k++)
        for (int b = 0; b < 8; b++, n++) {
          r[3 * n] = (i + basis[b][0]) * a0;
          r[3 * n + 1] = (j + basis[b][1]) * a0;
          r[3 * n + 2] = (k + basis[b][2]) * a0;
          el[n] = 1; /* particle element id 1 = Si, see el2Z below */
        }

  /* Abox: columns are the cell vectors (column-major 3x3), Bbox its inverse */
  double Abox[9] = {0}, Bbox[9] = {0};
  for (int k = 0; k < 3; k++) {
    Abox[4 * k] = nc * a0;
    Bbox[4 * k] = 1.0 / (nc * a0);
  }
  const int pbc[3] = {1, 1, 1};

  /* Tersoff, Phys. Rev. B 39, 5566 (1989), silicon only (tersoff_params.f90:85-131) */
  atx_bop_params par;
  memset(&par, 0, sizeof(par));
  par.kind = ATX_BOP_TERSOFF;
  par.nel = 1;
  par.Z[0] = 14;
  par.A[0] = 1.8308e3; par.B[0] = 4.7118e2; par.xi[0] = 1.0;
  par.lambda[0] = 2.4799; par.mu[0] = 1.7322; par.omega[0] = 1.0; par.mubo[0] = 0.0; par.m[0] = 1;
  par.beta[0] = 1.1000e-6; par.n[0] = 7.8734e-1; par.c[0] = 1.0039e5; par.d[0] = 1.6217e1;
  par.h[0] = -5.9825e-1;
  par.r1[0] = 2.70; par.r2[0] = 3.00;

  atx_ctx *ctx = NULL;
  atx_particles *p = NULL;
  atx_neighbors *nl = NULL;
  atx_bop *pot = NULL;
  printf("%s\n", atx_version());
  CHECK(atx_ctx_create(0, &ctx));
  CHECK(atx_particles_create(ctx, &p));
  CHECK(atx_particles_set_cell(p, Abox, Bbox, pbc));
  CHECK(atx_particles_set_elements(p, nat, el));
  CHECK(atx_particles_set_positions(p, nat, r));
  CHECK(atx_neighbors_create(ctx, 100, &nl));
  CHECK(atx_bop_create(ctx, &par, &pot));
  const int el2Z[1] = {14};
  CHECK(atx_bop_bind_to(pot, p, nl, 1, el2Z));

  double epot = 0.0, wpot[9] = {0};
  CHECK(atx_bop_energy_and_forces(pot, p, nl, NULL, &epot, f, wpot, NULL, NULL, NULL, NULL, NULL));
  printf("perfect crystal: epot/atom = %.8f eV\n", epot / nat);

  /* move the atoms (deterministic pseudo-random displacement) and evaluate again: outputs are
   * ADDED to the caller's arrays like in the reference, so they are cleared first */
  unsigned s = 12345u;
  for (int k = 0; k < 3 * nat; k++) {
    s = s * 1664525u + 1013904223u;
    r[k] += 0.1 * ((double)(s >> 8) / (double)(1u << 24) - 0.5);
  }
  CHECK(atx_particles_set_positions(p, nat, r));
  epot = 0.0;
  memset(f, 0, sizeof(double) * 3 * (size_t)nat);
  memset(wpot, 0, sizeof(wpot));
  CHECK(atx_bop_energy_and_forces(pot, p, nl, NULL, &epot, f, wpot, NULL, NULL, NULL, NULL, NULL));
  double fmax = 0.0, fsum[3] = {0, 0, 0};
  for (int k = 0; k < 3 * nat; k++) {
    if (fabs(f[k]) > fmax) fmax = fabs(f[k]);
    fsum[k % 3] += f[k];
  }
  printf("rattled crystal: epot/atom = %.8f eV, max |f| = %.6f eV/A, sum f = %.2e %.2e %.2e\n", epot / nat,
         fmax, fsum[0], fsum[1], fsum[2]);

  CHECK(atx_bop_destroy(pot));
  CHECK(atx_neighbors_destroy(nl));
  CHECK(atx_particles_destroy(p));
  CHECK(atx_ctx_destroy(ctx));
  free(r); free(f); free(el);
  return 0;
}
