/*
 * atomistica_b200 -- C-ABI of the B200-native hot path of Atomistica.
 *
 * This is "seam 0" of SURVEY.md 8(b): the entry points the reference's Fortran
 * host (src/python/f90/python_neighbors.f90, src/potentials/...) binds with
 * ISO_C_BINDING in place of its CPU loops.  Everything is extern "C", plain
 * pointers and sizes, Fortran array conventions:
 *   - r(3,nat), f(3,nat): contiguous, 3 doubles per atom
 *   - Abox(3,3), Bbox(3,3), wpot(3,3), wpot_per_at(3,3,nat): column-major
 *   - atom indices in host-visible neighbour arrays are 1-based; seed/last are
 *     1-based inclusive slot ranges of type intptr_t (NEIGHPTR_T) with one
 *     0-terminator slot after every atom's range
 * Every function returns 0 on success or a negative ERROR_* code
 * (src/support/error.f90:81-85); the message is read with atx_last_error().
 * energy_and_forces entry points ADD into epot/f/wpot/per-atom outputs, like
 * the reference (tls_reduce, bop_kernel.f90:1613-1628).  Optional outputs are
 * NULL when not requested.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with ATX_ERROR_DEVICE.
 */
#ifndef ATOMISTICA_B200_H
#define ATOMISTICA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATX_ERROR_NONE 0
#define ATX_ERROR_UNSPECIFIED (-1) /* ERROR_UNSPECIFIED, src/support/error.f90:82 */
#define ATX_ERROR_IO (-2)
#define ATX_ERROR_MPI (-4)
#define ATX_ERROR_DEVICE (-100) /* CUDA failure / no device */

typedef struct atx_ctx atx_ctx;
typedef struct atx_particles atx_particles;
typedef struct atx_neighbors atx_neighbors;
typedef struct atx_eam atx_eam;
typedef struct atx_bop atx_bop;
typedef struct atx_rebo2 atx_rebo2;
typedef struct atx_md atx_md;

/* ---- library / context ------------------------------------------------- */

/* replaces atomistica_startup/shutdown (src/support/atomistica.f90) for the device side */
int atx_ctx_create(int device, atx_ctx **ctx);
int atx_ctx_destroy(atx_ctx *ctx);
int atx_ctx_synchronize(atx_ctx *ctx);
/* get_full_error_string (src/support/error.f90) equivalent */
int atx_last_error(char *buf, int len);
const char *atx_version(void);
/* number of kernel launches issued by this library since the last reset */
long long atx_kernel_launches(int reset);

/* ---- particles_t: src/python/f90/python_particles.f90:84-170 ------------ */

int atx_particles_create(atx_ctx *ctx, atx_particles **p);
int atx_particles_destroy(atx_particles *p);
/* particles_set_cell (python_particles.f90:286-346); Bbox is the host's inverse (dgesv) */
int atx_particles_set_cell(atx_particles *p, const double *Abox, const double *Bbox, const int *pbc);
/* host r_non_cyc(3,nat) -> device; call when pos_rev changed (I_changed_positions) */
int atx_particles_set_positions(atx_particles *p, int nat, const double *r);
/* host el(nat) (1..nel element ids from particles_update_elements, :617-658) */
int atx_particles_set_elements(atx_particles *p, int nat, const int *el);
/* bare-metal variants: r / el already resident on the device (no copy) */
int atx_particles_set_positions_device(atx_particles *p, int nat, const double *r_dev);

/* ---- neighbors_t: src/python/f90/python_neighbors.f90:51-139 ------------ */

int atx_neighbors_create(atx_ctx *ctx, int avgn, atx_neighbors **nl);
int atx_neighbors_destroy(atx_neighbors *nl);
/* neighbors_request_interaction_range (:381-423): keeps the maximum, invalidates the list */
int atx_neighbors_request_interaction_range(atx_neighbors *nl, double cutoff);
/* neighbors_set (:337-379): verlet_shell ("skin"), default 0 like the Python host */
int atx_neighbors_set_verlet_shell(atx_neighbors *nl, double verlet_shell);
/* neighbors_update -> binning_update + fill_neighbor_list (:430-754, 904-959).
 * Fails with "Neighbor list overflow" when pairs + nat exceed nat*avgn (:716-718). */
int atx_neighbors_update(atx_neighbors *nl, atx_particles *p);
/* rebuild unconditionally (positions already on the device); used by the rebuild sweeps */
int atx_neighbors_rebuild(atx_neighbors *nl, atx_particles *p);
/* number of pairs, max neighbours per atom, n_cells(3), stencil half widths(3) */
int atx_neighbors_get_info(atx_neighbors *nl, long long *npairs, int *nebmax, int *n_cells,
                           int *stencil);
/* LAMMPS host (src/lammps/pair_style/pair_atomistica.cpp:385-460, lammps_neighbors.f90:177-217):
 * the host builds and communicates the list; ghosts are separate atoms (indices natloc..nat-1 of
 * the particles object), there are no periodic shifts, and -- as the pair style requests with
 * REQ_FULL|REQ_GHOST -- ghost atoms carry lists of their own.  ilist/numneigh/firstneigh are
 * LAMMPS's NeighList arrays for the inum+gnum atoms that have a list (0-based indices).  After this
 * call atx_neighbors_update is a no-op apart from following the positions, and
 * atx_{eam,bop,pair}_energy_and_forces return epot / wpot summed over the OWNED atoms' bonds and
 * forces on owned atoms only (ghost rows of f receive zero: the gather formulation needs no
 * reverse communication; the ghost shell must be 2 x cutoff wide, which is what
 * particles_get_border already asks LAMMPS for).  atx_rebo2_energy_and_forces works on such a list
 * too (every bond among the local atoms is evaluated, energy / virial counted for owned ends, ghost
 * rows cleared) and needs the ghosts within 5 bond cutoffs; per-bond outputs and the screened variant
 * are refused there.  Call again after every reneighbouring. */
int atx_neighbors_set_external(atx_neighbors *nl, atx_particles *p, int natloc, int inum,
                               const int *ilist, const int *numneigh, const int *const *firstneigh);
/* number of list builds and of updates answered without a rebuild (Verlet shell, neighbors.f90:552-590) */
int atx_neighbors_get_counters(atx_neighbors *nl, long long *nbuilds, long long *nreused);
/* largest interaction range requested so far (neighbors_request_interaction_range,
 * python_neighbors.f90:381-423 / lammps_neighbors.f90:223-251) and the Verlet shell */
int atx_neighbors_get_interaction_range(atx_neighbors *nl, double *range, double *verlet_shell);
/* request_interaction_range(nl, cutoff, el1, el2) with the element pair (particle element ids, 1-based):
 * same effect on the list as atx_neighbors_request_interaction_range, and the per-pair maximum is
 * kept for hosts that build the list themselves -- neighbors_get_cutoff(nl, i, j)
 * (lammps_neighbors.f90:223-251) feeds LAMMPS' cutsq(i,j) from it; 0 for pairs nobody asked for.
 * The bind_to entry points of the potentials use it. */
int atx_neighbors_request_interaction_range_pair(atx_neighbors *nl, double cutoff, int el1, int el2);
int atx_neighbors_get_pair_range(atx_neighbors *nl, int el1, int el2, double *range);
/* List post-processing on the device-resident list (no copy-back of the list):
 * f_get_coordination_numbers (src/python/f90/neighbors_wrap.f90:271-302; c[nat], original atom order) and
 * the helpers of src/python/c/analysis.c -- pair_distribution (:29-106), angle_distribution (:108-206),
 * bond_angles (:208-318) -- over ALL entries of the list (the reference applies them to the (i, j, r)
 * arrays of get_neighbors).  h / h2: mean and variance per bin (nbins doubles each), m[nat]. */
int atx_neighbors_coordination_numbers(atx_neighbors *nl, atx_particles *p, double cutoff, int *c);
int atx_neighbors_pair_distribution(atx_neighbors *nl, atx_particles *p, int nbins, double cutoff,
                                    double *h, double *h2);
int atx_neighbors_angle_distribution(atx_neighbors *nl, atx_particles *p, int nbins, double cutoff,
                                     double *h, double *h2);
int atx_neighbors_bond_angles(atx_neighbors *nl, atx_particles *p, int moment, double cutoff, double *m);
/* host view of the list in the reference's layout and order (seed(nat+1), last(nat+1),
 * neighbors(capacity), dc(3,capacity)); used by f_get_all_neighbors & co
 * (src/python/f90/neighbors_wrap.f90:211-550) */
int atx_neighbors_copy_to_host(atx_neighbors *nl, intptr_t *seed, intptr_t *last, int *neighbors,
                               int *dc, long long capacity);

/* ---- TabulatedAlloyEAM: src/potentials/eam/tabulated_alloy_eam.f90 ------- */

/* one simple_spline_t (src/support/simple_spline.f90:42-61): n points, n-1 intervals */
typedef struct {
  int n;
  double x0, dx;
  const double *y, *coeff1, *coeff2, *coeff3, *dcoeff1, *dcoeff2, *dcoeff3;
} atx_spline;

/* ndb database elements; fF[ndb], frho[ndb], fphi[ndb*ndb] (column-major (i,j), symmetric) */
int atx_eam_create(atx_ctx *ctx, int ndb, const atx_spline *fF, const atx_spline *frho,
                   const atx_spline *fphi, double cutoff, atx_eam **pot);
/* TabulatedEAM (one element, funcfl tables; src/potentials/eam/tabulated_eam.f90:141-511): fZ is
 * the effective-charge spline AFTER scale_y_axis(sqrt(0.5 Hartree Bohr)) (:199); pair term Z**2/r.
 * Used through atx_eam_bind_to / atx_eam_energy_and_forces (no mask, no wpot_per_at). */
int atx_eam_create_funcfl(atx_ctx *ctx, const atx_spline *fF, const atx_spline *frho,
                          const atx_spline *fZ, double cutoff, atx_eam **pot);
int atx_eam_destroy(atx_eam *pot);
/* bind_to (:297-350): el2db[nel] maps particle element ids (1..nel) to database ids (1..ndb, <=0:
 * ignored); requests the interaction range */
int atx_eam_bind_to(atx_eam *pot, atx_particles *p, atx_neighbors *nl, int nel, const int *el2db);
/* energy_and_forces (:360-415) + kernel (:423-627) */
int atx_eam_energy_and_forces(atx_eam *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                              double *epot, double *f, double *wpot, double *epot_per_at,
                              double *wpot_per_at);

/* ---- Tersoff / Kumagai / Brenner: src/potentials/bop/bop_kernel.f90 ------ */

#define ATX_BOP_TERSOFF 1
#define ATX_BOP_KUMAGAI 2
#define ATX_BOP_BRENNER 3
#define ATX_BOP_JUSLIN 4
#define ATX_BOP_MAX_EL 3
#define ATX_BOP_MAX_PAIRS 6

/* Parameter database in the layout of the Fortran BOP_DB_TYPEs (pair arrays in PAIR_INDEX
 * order, src/macros.inc:123).  Unused fields are ignored for a given kind.
 *   Tersoff  (tersoff_params.f90:33-84):  A B xi lambda mu omega mubo m | beta n c d h | r1 r2
 *   Kumagai  (kumagai_params.f90:30-118): A B lambda1 lambda2 alpha ibeta | eta delta c1..c5 h | r1 r2
 *   Brenner  (brenner_params.f90:33-70):  D0 r0 S beta gamma c d h mu n m | r1 r2 */
typedef struct {
  int kind;
  int nel;
  int Z[ATX_BOP_MAX_EL]; /* atomic numbers of db%el */
  /* pair parameters */
  double A[ATX_BOP_MAX_PAIRS], B[ATX_BOP_MAX_PAIRS], xi[ATX_BOP_MAX_PAIRS];
  double lambda[ATX_BOP_MAX_PAIRS], mu[ATX_BOP_MAX_PAIRS], omega[ATX_BOP_MAX_PAIRS];
  double mubo[ATX_BOP_MAX_PAIRS]; /* Kumagai alpha */
  int m[ATX_BOP_MAX_PAIRS];       /* Kumagai integer beta */
  double D0[ATX_BOP_MAX_PAIRS], r0[ATX_BOP_MAX_PAIRS], S[ATX_BOP_MAX_PAIRS];
  double pbeta[ATX_BOP_MAX_PAIRS], gamma[ATX_BOP_MAX_PAIRS];
  double pc[ATX_BOP_MAX_PAIRS], pd[ATX_BOP_MAX_PAIRS], ph[ATX_BOP_MAX_PAIRS];
  double pn[ATX_BOP_MAX_PAIRS];
  double r1[ATX_BOP_MAX_PAIRS], r2[ATX_BOP_MAX_PAIRS];
  /* element parameters */
  double beta[ATX_BOP_MAX_EL], n[ATX_BOP_MAX_EL], c[ATX_BOP_MAX_EL], d[ATX_BOP_MAX_EL];
  double h[ATX_BOP_MAX_EL];
  double eta[ATX_BOP_MAX_EL], delta[ATX_BOP_MAX_EL];
  double c1[ATX_BOP_MAX_EL], c2[ATX_BOP_MAX_EL], c3[ATX_BOP_MAX_EL], c4[ATX_BOP_MAX_EL];
  double c5[ATX_BOP_MAX_EL];
} atx_bop_params;

int atx_bop_create(atx_ctx *ctx, const atx_bop_params *par, atx_bop **pot);

/* Screened variants TersoffScr / KumagaiScr / BrennerScr: the same modules compiled with SCREENING
 * and CUTOFF_T = exp_cutoff_t (tersoff_scr.f90:46-47, kumagai_scr.f90, brenner_scr.f90).  r1/r2 of
 * atx_bop_params are then the INNER cutoff; or1/or2 the outer (attractive/repulsive) cutoff,
 * bor1/bor2 the bond-order cutoff, Cmin/Cmax the bounds of the Baskes screening function
 * (tersoff_params.f90:85-102, bop_kernel.f90:682-995), all pair-indexed.  bind_to and
 * energy_and_forces are the unscreened entry points; bind_to requests the longer list cutoff
 * of default_bind_to_func.f90:106-130. */
typedef struct {
  double or1[ATX_BOP_MAX_PAIRS], or2[ATX_BOP_MAX_PAIRS];
  double bor1[ATX_BOP_MAX_PAIRS], bor2[ATX_BOP_MAX_PAIRS];
  double Cmin[ATX_BOP_MAX_PAIRS], Cmax[ATX_BOP_MAX_PAIRS];
} atx_bop_screening;
int atx_bop_create_screened(atx_ctx *ctx, const atx_bop_params *par, const atx_bop_screening *scr,
                            atx_bop **pot);
/* Juslin (src/potentials/bop/juslin/: W-C-H of Juslin et al., Fe-C-H of Kuopanportti et al.):
 * Brenner's functional form with NON-symmetric pair parameters (nel**2 entries, index
 * j + (i-1)*nel, macros.inc:139) and triplet-indexed alpha/omega/m of h() (index
 * k + nel*(j-1 + nel*(i-1)), macros.inc:146).  The caller passes the database AFTER the mirroring
 * of BIND_TO_FUNC (juslin_module.f90:283-312).  The object is used through atx_bop_bind_to /
 * atx_bop_energy_and_forces / atx_bop_destroy. */
typedef struct {
  int nel;
  int Z[ATX_BOP_MAX_EL];
  double D0[9], r0[9], S[9], beta[9], gamma[9], c[9], d[9], h[9], n[9], r1[9], r2[9];
  double alpha[27], omega[27];
  int m[27];
} atx_juslin_params;
int atx_bop_create_juslin(atx_ctx *ctx, const atx_juslin_params *par, atx_bop **pot);
/* JuslinScr (src/potentials/bop/juslin_scr/juslin_scr.f90 = juslin_module.f90 with SCREENING): r1/r2
 * of atx_juslin_params are the inner cutoff; or1/or2, bor1/bor2, Cmin/Cmax as in atx_bop_screening but
 * with the nel**2 non-symmetric pair index and AFTER the mirroring of BIND_TO_FUNC
 * (juslin_module.f90:285-298).  All three switching functions are the cosine form
 * (juslin_func.f90:27-122); C_dr_cut = Cmax**2/(4 (Cmax-1)) for every pair (:301-307). */
typedef struct {
  double or1[9], or2[9], bor1[9], bor2[9], Cmin[9], Cmax[9];
} atx_juslin_screening;
int atx_bop_create_juslin_screened(atx_ctx *ctx, const atx_juslin_params *par,
                                   const atx_juslin_screening *scr, atx_bop **pot);
int atx_bop_destroy(atx_bop *pot);
/* BIND_TO_FUNC (default_bind_to_func.f90:25-146): el2Z[nel] are the atomic numbers of the
 * particle element ids; builds Z2db and requests r2 of every present pair */
int atx_bop_bind_to(atx_bop *pot, atx_particles *p, atx_neighbors *nl, int nel, const int *el2Z);
/* COMPUTE_FUNC (default_compute_func.f90:25-102) + BOP_KERNEL.  Per-bond outputs are indexed
 * by host neighbour-list slot (this%nbb, bop_kernel.f90:1371,1407,1512). */
int atx_bop_energy_and_forces(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                              double *epot, double *f, double *wpot, double *epot_per_at,
                              double *epot_per_bond, double *f_per_bond, double *wpot_per_at,
                              double *wpot_per_bond);

/* ---- Rebo2: src/potentials/bop/rebo2/ -------------------------------------- */

/* Everything rebo2_db_init_with_parameters (rebo2_db.f90:81-303) leaves in rebo2_t that the
 * kernel reads.  Table coefficients are the Fortran coeff(nboxs,4,4[,4]) arrays
 * (table2d.f90:84-226, table3d.f90:85-284), nboxs = 4*4*9 (F, T) and 5*5 (P). */
typedef struct {
  double cc_B1, cc_B2, cc_B3, cc_beta1, cc_beta2, cc_beta3, cc_Q, cc_A, cc_alpha;
  double ch_B1, ch_beta1, ch_Q, ch_A, ch_alpha;
  double hh_B1, hh_beta1, hh_Q, hh_A, hh_alpha;
  double cc_g_theta[6];
  double cc_g1_coeff[18], cc_g2_coeff[18]; /* g_coeff_t%c(6,3) */
  double spgh[18];                         /* SPGH(6,3) */
  int igh[25];
  double conalp, conear[36];               /* conear(6,6) */
  double conpe[3], conan[3], conpf[3];
  double cut_in_l[10], cut_in_h[10], cut_in_h2[10];
  int with_dihedral;
  const double *Fcc, *Fch, *Fhh, *Tcc; /* 144*64 doubles each */
  const double *Pcc, *Pch;             /* 25*16 doubles each */
} atx_rebo2_params;

int atx_rebo2_create(atx_ctx *ctx, const atx_rebo2_params *par, atx_rebo2 **pot);
/* Rebo2Scr (src/potentials/bop/rebo2/rebo2_scr.f90:60-64, parameters rebo2_type.f90 /
 * rebo2_db.f90:92-112): screened C-C bonds with separate attractive-repulsive, bond-order and
 * neighbour-count cutoffs; the inner C-C cutoff is par->cut_in_*[0].  The dihedral term of this
 * variant is not available (par->with_dihedral must be 0).  The returned object is used with the
 * atx_rebo2_* entry points below. */
typedef struct {
  double cc_ar_r1, cc_ar_r2, cc_bo_r1, cc_bo_r2, cc_nc_r1, cc_nc_r2, Cmin, Cmax;
} atx_rebo2_screening;
int atx_rebo2_create_screened(atx_ctx *ctx, const atx_rebo2_params *par,
                              const atx_rebo2_screening *scr, atx_rebo2 **pot);
int atx_rebo2_destroy(atx_rebo2 *pot);
/* BIND_TO_FUNC (rebo2_module.f90:70-135) */
int atx_rebo2_bind_to(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, int nel,
                      const int *el2Z);
/* COMPUTE_FUNC (rebo2_module.f90:143-223) + kernel (bop_kernel_rebo2.f90) */
int atx_rebo2_energy_and_forces(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, double *epot,
                                double *f, double *wpot, double *epot_per_at,
                                double *epot_per_bond, double *f_per_bond, double *wpot_per_at,
                                double *wpot_per_bond);

/* ---- pair potentials: src/potentials/pair_potentials/{lj_cut,harmonic,double_harmonic}.f90 ---- */

#define ATX_PAIR_LJCUT 1            /* p = {epsilon, sigma, cutoff};        shift: lj_cut.f90:176-181 */
#define ATX_PAIR_HARMONIC 2         /* p = {k, r0, cutoff};                 shift: harmonic.f90:134-137 */
#define ATX_PAIR_DOUBLE_HARMONIC 3  /* p = {k1, r1, k2, r2, cutoff} */
#define ATX_PAIR_BORN_MAYER 4       /* p = {A, rho, cutoff}; epot and f only (born_mayer.f90:200-275) */
#define ATX_PAIR_R6 5               /* p = {A, r0, cutoff}: A/(r0+r)**6 (r6.f90:150-215) */
typedef struct atx_pair atx_pair;
typedef struct {
  int kind;
  double p[8];
  int shift; /* shift the potential to zero at the cutoff (LJCut, Harmonic) */
} atx_pair_params;
int atx_pair_create(atx_ctx *ctx, const atx_pair_params *par, atx_pair **pot);
int atx_pair_destroy(atx_pair *pot);
/* *_bind_to: el1/el2 are the element filters of src/core/filter.f90 (bit k set = particle element
 * id k takes part; filter_from_string builds them from "*" or a list of symbols) */
int atx_pair_bind_to(atx_pair *pot, atx_particles *p, atx_neighbors *nl, int el1, int el2);
/* mask: LJCut only (lj_cut.f90:232-262); the other two have no mask argument in the reference */
int atx_pair_energy_and_forces(atx_pair *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                               double *epot, double *f, double *wpot, double *epot_per_at,
                               double *wpot_per_at);
int atx_pair_set_store_outputs(atx_pair *pot, int on);

/* ---- output mode ------------------------------------------------------------ */
/* The reference ADDS into f / epot_per_at / wpot_per_at (the Python host hands in fresh zeroed
 * arrays, LAMMPS its live force array) and that is the default here.  A host that always passes a
 * fresh buffer can switch a potential object to STORE mode: the three per-atom outputs are then
 * written (=) straight from the device into the caller's arrays -- no zero-initialisation and no
 * host-side accumulation pass.  epot, wpot and the per-bond outputs keep accumulating. */
int atx_eam_set_store_outputs(atx_eam *pot, int on);
int atx_bop_set_store_outputs(atx_bop *pot, int on);
int atx_rebo2_set_store_outputs(atx_rebo2 *pot, int on);

/* ---- device-resident MD driver -------------------------- */
/* velocity-Verlet (src/standalone/verlet.f90:100-235) with the Verlet-shell rebuild rule
 * 2*accum_max_dr >= verlet_shell (src/standalone/neighbors.f90:552-590); positions,
 * velocities and forces never leave the GPU between atx_md_run calls. */

#define ATX_POT_EAM 1
#define ATX_POT_BOP 2
#define ATX_POT_REBO2 3

int atx_md_create(atx_ctx *ctx, int pot_kind, void *pot, atx_particles *p, atx_neighbors *nl,
                  const double *mass /* per atom, amu */, const double *v /* (3,nat) */, double dt,
                  atx_md **md);
int atx_md_destroy(atx_md *md);
/* advance nsteps; epot / ekin of the last step are returned */
int atx_md_run(atx_md *md, int nsteps, double *epot, double *ekin);
/* copy r, v, f (3,nat each, original atom order) back to the host; any pointer may be NULL */
int atx_md_get_state(atx_md *md, double *r, double *v, double *f);
/* number of neighbour-list rebuilds so far and device milliseconds spent in the last run */
int atx_md_get_stats(atx_md *md, long long *nrebuilds, double *last_run_ms);

/* ---- spatial domain decomposition over the GPUs of one node -------------------- */
/* One process per GPU.  Slab decomposition along the first cell vector, ghost-atom halo exchange of
 * positions with the two slab neighbours and the global rebuild decision run over NCCL
 * (src/standalone/domain_decomposition.f90:494-970 is the behavioural model; MPI_context.f90 the
 * backend it replaces).  NCCL is loaded lazily with dlopen. */
typedef struct atx_dd atx_dd;
typedef struct atx_ddmd atx_ddmd;
/* rank 0 creates the 128-byte NCCL unique id and hands it to the other ranks out of band */
int atx_dd_get_unique_id(char *id128);
int atx_dd_create(atx_ctx *ctx, int rank, int nranks, const char *id128, atx_dd **dd);
int atx_dd_destroy(atx_dd *dd);
/* Domain-decomposed NVE driver.  Abox/Bbox/pbc describe the GLOBAL cell; every rank passes the
 * nown atoms it owns (fractional coordinate along the first cell vector in [rank/P, (rank+1)/P)):
 * global ids, particle element ids, positions (global frame), velocities (A/fs), masses (amu).
 * rc is the interaction range of the potential (which must already be bound to the element map with
 * atx_<pot>_bind_to(pot, NULL, NULL, nel, ...)), skin the Verlet shell. */
int atx_dd_md_create(atx_dd *dd, int pot_kind, void *pot, const double *Abox, const double *Bbox,
                     const int *pbc, double rc, double skin, int avgn, int nown, const long long *id,
                     const int *el, const double *r, const double *v, const double *mass, double dt,
                     atx_ddmd **md);
int atx_dd_md_destroy(atx_ddmd *md);
/* advance nsteps on all ranks (collective); GLOBAL potential and kinetic energy of the last step */
int atx_dd_md_run(atx_ddmd *md, int nsteps, double *epot, double *ekin);
int atx_dd_md_get_count(atx_ddmd *md, int *nown, int *nghost);
/* owned atoms of this rank: global ids, positions (global frame), velocities, forces */
int atx_dd_md_get_state(atx_ddmd *md, long long *id, double *r, double *v, double *f);
int atx_dd_md_get_stats(atx_ddmd *md, long long *nrebuilds, double *last_run_ms);
/* Host wall time (ms, accumulated) of the eight phases of the decomposition rebuild (migration,
 * ghost construction, local list -- communicate_particles / communicate_ghosts,
 * domain_decomposition.f90:494-826) and whether the peer-to-peer step path is active. */
int atx_dd_md_get_profile(atx_ddmd *md, double *ms8, int *p2p);

/* ---- measurement hooks (bench.py) -------------------------------------------- */
/* CUDA-event timing of the library's own kernels on the launching stream.  When enabled every
 * launch of a profiled kernel is bracketed by an event pair; atx_profile_read resolves them. */
int atx_profile_enable(atx_ctx *ctx, int on);
/* total milliseconds and launch count of kernel `name` since the last enable; names:
 * "eam_density", "eam_force", "bop_force", "rebo2_force", "nl_pairs_count", "nl_pairs_fill" */
int atx_profile_read(atx_ctx *ctx, const char *name, double *total_ms, long long *count);
/* FP64 FMA throughput of the device (TFLOP/s), measured with a register-resident DFMA chain */
int atx_measure_fp64_peak(atx_ctx *ctx, double *tflops);
/* device copy bandwidth (GB/s, read+write bytes), measured with a 1 GiB grid-stride copy */
int atx_measure_copy_bandwidth(atx_ctx *ctx, double *gbs);

/* page-locked host memory for the arrays the host hands to the library every call (r_non_cyc, f):
 * copies from/to such buffers run at full PCIe speed and need no staging */
int atx_host_alloc_pinned(size_t bytes, void **ptr);
int atx_host_free_pinned(void *ptr);
/* Page-lock an array the HOST owns (cudaHostRegister): the reference's hosts keep their positions
 * in their own storage -- particles_t%r_non_cyc (src/python/f90/python_particles.f90:120-135),
 * LAMMPS' atom->x aliased by lammps_particles.f90:176-200, the numpy array behind ase.Atoms -- and
 * hand the same buffer to every call; registered once, atx_particles_set_positions DMAs straight out
 * of it with no staging copy.  Returns an error (and leaves the array usable as ordinary pageable
 * memory) when the range cannot be locked; a range that is already page-locked is not an error
 * (*already = 1, do not unregister it).  `already` may be NULL. */
int atx_host_register(void *ptr, size_t bytes, int *already);
int atx_host_unregister(void *ptr);

/* ---- host-side init helpers ------------------------------------------------ */
/* The reference computes these on the host in Fortran; a Fortran host keeps doing so and passes
 * the results into atx_*_create.  They are exported so that non-Fortran hosts (the Python mirror
 * in this repo) can build the same inputs. */

/* simple_spline_init (simple_spline.f90:127-195); out arrays: y[n], d2y[n], coeff*[n-1], dcoeff*[n-1] */
int atx_host_spline_init(int n, double x0, double dx, const double *y_in, double *y, double *d2y,
                         double *coeff1, double *coeff2, double *coeff3, double *dcoeff1,
                         double *dcoeff2, double *dcoeff3);
/* gaussn (f_linearalgebra.f90:599-637): solve A X = B, A(n,n), B(n,m) column-major, in place */
int atx_host_gaussn(int n, double *A, int m, double *B);
/* table2d_init / table3d_init; values/derivatives are Fortran arrays (0:nx,0:ny[,0:nz]) */
int atx_host_table2d_init(int nx, int ny, const double *values, const double *dvdx,
                          const double *dvdy, double *coeff);
int atx_host_table3d_init(int nx, int ny, int nz, const double *values, const double *dvdx,
                          const double *dvdy, const double *dvdz, double *coeff);
/* rebo2_db_make_cc_g_spline (rebo2_db.f90:405-524) */
int atx_host_rebo2_g_spline(const double *theta, const double *g1, const double *dg1,
                            const double *d2g1, const double *g2, double *g1_coeff,
                            double *g2_coeff);

#ifdef __cplusplus
}
#endif
#endif
