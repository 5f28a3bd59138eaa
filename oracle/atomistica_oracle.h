/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the Atomistica hot path (neighbour list, generic
 * bond-order potentials, REBO2, tabulated alloy EAM).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (atomistica_b200/) never links, imports
 * or calls it.
 *
 * Parity status: the Fortran reference cannot be compiled in the build
 * container (no Fortran compiler), so this oracle is pinned against the
 * reference's own known-answer tests (tests/golden/kat.json, generated from
 * /root/reference/tests/*.py) and against the reference's Fortran kernels
 * executed statement by statement through tests/fortran_subset.py
 * (tests/test_func_vs_reference.py) rather than against output of a compiled
 * reference binary.
 *
 * Array conventions are the Fortran host's: r(3,nat) contiguous per atom,
 * Abox(3,3) column-major (columns are the cell vectors), 1-based atom indices
 * in `neighbors`, 1-based inclusive seed/last slot ranges, dc(3,slot).
 */
#ifndef ATOMISTICA_ORACLE_H
#define ATOMISTICA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- neighbour list: src/python/f90/python_neighbors.f90:570-959 ---- */

typedef struct {
  int n_cells[3];
  int dx, dy, dz;          /* stencil half widths */
  double rec_cell_size[9]; /* column-major (3,3) */
} orc_binning_t;

int orc_binning_init(const double *Abox, const double *Bbox, double cutoff, double bin_size,
                     orc_binning_t *b);

/* OpenMP threads of the build (default 1 = the serial loop) */
void orc_nl_set_threads(int n);
/* returns number of pairs, or -1 on "Neighbor list overflow" (capacity = size of neighbors[]) */
long orc_nl_build(int nat, const double *r, const double *Abox, const double *Bbox, const int *pbc,
                  double cutoff, long capacity, intptr_t *seed, intptr_t *last, int *neighbors,
                  int *dc);

/* ---- simple_spline: src/support/simple_spline.f90 ---- */

typedef struct {
  int n;
  double x0, dx, cut;
  const double *y, *c1, *c2, *c3, *d1, *d2, *d3;
} orc_spline_t;

/* ---- tabulated alloy EAM: src/potentials/eam/tabulated_alloy_eam.f90:423-627 ---- */

/* OpenMP threads of the EAM kernel (default 1 = serial summation order) */
void orc_eam_set_threads(int n);
int orc_eam_energy_and_forces(int nat, int natloc, const double *r, const double *Abox,
                              const int *eldb, const intptr_t *seed, const intptr_t *last,
                              const int *neighbors, const int *dc, int ndb, const orc_spline_t *fF,
                              const orc_spline_t *frho, const orc_spline_t *fphi, double cutoff,
                              const int *mask, double *epot, double *f, double *wpot,
                              double *epot_per_at, double *wpot_per_at);

/* ---- TabulatedEAM (funcfl, one element): src/potentials/eam/tabulated_eam.f90:334-511 ---- */
/* fZ must already carry the factor sqrt(0.5 Hartree Bohr); `in[i]` = atom selected by the element
 * filter.  No mask and no per-atom virial in the reference. */
int orc_eam_funcfl_energy_and_forces(int nat, const double *r, const double *Abox, const int *in,
                                     const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                     const int *dc, const orc_spline_t *fF, const orc_spline_t *fZ,
                                     const orc_spline_t *frho, double cutoff, double *epot, double *f,
                                     double *wpot, double *epot_per_at);

/* ---- generic bond-order potentials: src/potentials/bop/bop_kernel.f90 ---- */

enum { ORC_TERSOFF = 1, ORC_KUMAGAI = 2, ORC_BRENNER = 3, ORC_JUSLIN = 4 };

/* pair-parameter rows */
enum {
  /* Tersoff */ OT_A = 0, OT_B, OT_XI, OT_LAMBDA, OT_MU, OT_OMEGA, OT_MUBO,
  /* Kumagai */ OK_A = 0, OK_B, OK_LAMBDA1, OK_LAMBDA2, OK_ALPHA,
  /* Brenner */ OB_D0 = 0, OB_R0, OB_S, OB_BETA, OB_GAMMA, OB_C, OB_D, OB_H, OB_MU, OB_N
};
/* element-parameter rows */
enum {
  /* Tersoff */ OTE_BETA = 0, OTE_N, OTE_C, OTE_D, OTE_H,
  /* Kumagai */ OKE_ETA = 0, OKE_DELTA, OKE_C1, OKE_C2, OKE_C3, OKE_C4, OKE_C5, OKE_H
};

/* Juslin (W-C-H, Fe-C-H; src/potentials/bop/juslin/): Brenner's functional form with
 * NON-symmetric pair indices PAIR_INDEX_NS (nel**2 entries, macros.inc:139) in the OB_* rows and
 * triplet-indexed alpha/omega/m (TRIPLET_INDEX_NS, macros.inc:146) for h(). */
typedef struct {
  int kind;
  int nel;
  double pp[12][9]; /* pair parameters [row][pair-1] */
  double ep[8][3];  /* element parameters [row][el-1] */
  int ip[9];        /* integer pair parameter: Tersoff/Brenner m, Kumagai beta */
  double r1[9], r2[9];
  double t_alpha[27], t_omega[27]; /* Juslin triplet parameters */
  int t_m[27];
} orc_bop_params_t;

int orc_bop_energy_and_forces(const orc_bop_params_t *par, int nat, int natloc, const double *r,
                              const double *Abox, const int *el, const intptr_t *seed,
                              const intptr_t *last, const int *neighbors, const int *dc,
                              const int *mask, double *epot, double *f, double *wpot,
                              double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                              double *wpot_per_at, double *wpot_per_bond);

/* Screened variants (TersoffScr, KumagaiScr, BrennerScr; JuslinScr with trigonometric cutoffs): outer / bond-order cutoffs and the
 * Baskes screening bounds of the *_Scr parameter sets (parameters.py), pair-indexed.  In these
 * variants r1/r2 of orc_bop_params_t are the INNER cutoff and every cutoff is exp_cutoff_t. */
typedef struct {
  double or1[9], or2[9], bor1[9], bor2[9], Cmin[9], Cmax[9]; /* 9: Juslin's el x el pair index */
} orc_bop_scr_t;

int orc_bop_scr_energy_and_forces(const orc_bop_params_t *par, const orc_bop_scr_t *scr, int nat,
                                  int natloc, const double *r, const double *Abox, const int *el,
                                  const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                  const int *dc, const int *mask, double *epot, double *f,
                                  double *wpot, double *epot_per_at, double *epot_per_bond,
                                  double *f_per_bond, double *wpot_per_at, double *wpot_per_bond);

/* ---- pair potentials: src/potentials/pair_potentials/{lj_cut,harmonic,double_harmonic}.f90 ---- */

enum { ORC_PAIR_LJCUT = 1, ORC_PAIR_HARMONIC = 2, ORC_PAIR_DOUBLE_HARMONIC = 3, ORC_PAIR_BORN_MAYER = 4,
       ORC_PAIR_R6 = 5 };
/* par: LJCut {epsilon, sigma, cutoff}; Harmonic {k, r0, cutoff}; DoubleHarmonic {k1, r1, k2, r2,
 * cutoff}; BornMayer {A, rho, cutoff}; r6 {A, r0, cutoff}.  el[] are PARTICLE element ids (1-based), el1/el2 the filter bit masks of filter.f90.
 * Returns -1 when a mask is passed to a potential that has none. */
int orc_pair_energy_and_forces(int kind, const double *par, int shift, int nat, const double *r,
                               const double *Abox, const int *el, int el1, int el2,
                               const intptr_t *seed, const intptr_t *last, const int *neighbors,
                               const int *dc, const int *mask, double *epot, double *f, double *wpot,
                               double *epot_per_at, double *wpot_per_at);

/* cutoff functions of src/support/cutoff.f90 (0 = trig_off :152-196, 1 = exp_cutoff :232-293) */
void orc_cutoff_eval(int kind, double r1, double r2, double r, double *val, double *dval);

/* one evaluation of a function of {tersoff,kumagai,brenner,juslin}_func.f90 (test hook; 1-based indices).
 * which: 0 VA(dr), 1 VR(dr), 2 g(costh), 3 bo(zij; fcij, faij), 4 h(dr) */
void orc_bop_func(const orc_bop_params_t *p, int which, int ktypj, int ktypi, int ktypk, int ijpot,
                  int ikpot, double x, double fcij, double faij, double *val, double *dval);

/* ---- REBO2: src/potentials/bop/rebo2/ ---- */

typedef struct {
  int nx, ny;
  const double *coeff; /* Fortran coeff(nboxs,4,4) */
} orc_table2d_t;

typedef struct {
  int nx, ny, nz;
  const double *coeff; /* Fortran coeff(nboxs,4,4,4) */
} orc_table3d_t;

typedef struct {
  double cc_B1, cc_B2, cc_B3, cc_beta1, cc_beta2, cc_beta3, cc_Q, cc_A, cc_alpha;
  double ch_B1, ch_beta1, ch_Q, ch_A, ch_alpha;
  double hh_B1, hh_beta1, hh_Q, hh_A, hh_alpha;
  double cc_g_theta[6];
  double cc_g1_coeff[18]; /* Fortran c(6,3) */
  double cc_g2_coeff[18];
  double spgh[18];        /* Fortran SPGH(6,3) */
  int igh[25];
  double conalp;
  double conear[36];      /* Fortran conear(6,6) */
  double conpe[3], conan[3], conpf[3];
  double cut_in_l[10], cut_in_h[10], cut_in_h2[10];
  int with_dihedral;
  orc_table3d_t Fcc, Fch, Fhh, Tcc;
  orc_table2d_t Pcc, Pch;
} orc_rebo2_params_t;

void orc_table2d_eval(const orc_table2d_t *t, double nhi, double nci, double *v, double *dvdh,
                      double *dvdc);
void orc_table3d_eval(const orc_table3d_t *t, double nti, double ntj, double nconj, double *v,
                      double *dvdi, double *dvdj, double *dvdc);

/* Rebo2Scr (rebo2_scr.f90): cut_in_* of orc_rebo2_params_t hold the INNER cutoff (cc_in_r1/r2 =
 * 1.95/2.25, rebo2_type.f90:205-206); the attractive/repulsive, bond-order and neighbour-count
 * cutoffs of C-C and the screening bounds come here (rebo2_type.f90:65-66, 208-213).  C-H and H-H
 * use their single cutoff for all three families (rebo2_db.f90:194-219). */
typedef struct {
  double cc_ar_r1, cc_ar_r2, cc_bo_r1, cc_bo_r2, cc_nc_r1, cc_nc_r2, Cmin, Cmax;
} orc_rebo2_scr_t;

int orc_rebo2_scr_energy_and_forces(const orc_rebo2_params_t *par, const orc_rebo2_scr_t *scr, int nat,
                                    int natloc, const double *r, const double *Abox, const int *ktyp,
                                    const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                    const int *dc, double *epot, double *f, double *wpot,
                                    double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                    double *wpot_per_at, double *wpot_per_bond);

/* one evaluation of a function of rebo2_func.f90 (test hook).  which: 0 fconj(x), 1 fCin(ijpot = i1, x),
 * 2 VA(i1, x), 3 VR(i1, x), 4 g(ktyp = i1, costh = x, n = y) -> val, d/dcosth, d/dN,
 * 5 bo(ktypi = i1, zij = x, fcij = y, faij = z), 6 h(ijpot = i1, ikpot = i2, dr = x), 7 Z2pair(i1, i2) */
void orc_rebo2_func(const orc_rebo2_params_t *p, int which, int i1, int i2, double x, double y, double z,
                    double *out);

int orc_rebo2_energy_and_forces(const orc_rebo2_params_t *par, int nat, int natloc,
                                const double *r, const double *Abox, const int *ktyp,
                                const intptr_t *seed, const intptr_t *last, const int *neighbors,
                                const int *dc, double *epot, double *f, double *wpot,
                                double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                double *wpot_per_at, double *wpot_per_bond);

#ifdef __cplusplus
}
#endif
#endif
