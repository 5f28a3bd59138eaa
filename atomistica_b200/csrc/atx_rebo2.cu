// REBO2 (Brenner et al. 2002) on the device -- non-screened build (DIHEDRAL + NUM_NEIGHBORS).
//
// Replaces rebo2_kernel (src/potentials/bop/rebo2/bop_kernel_rebo2.f90:700-2883, SCREENING
// undefined) and the functions it inlines (rebo2_func.f90: fconj :31-57, fCin :63-85, VA :173-221,
// VR :230-277, g :289-347, cc_g_from_spline :351-397, bo :403-425, h :431-461, Z2pair :467-485)
// plus table2d_eval / table3d_eval (src/special/table2d.f90:255-318, table3d.f90:313-389).
//
//   k_rebo2_bonds  loop 1 + the neighbour counts nn(C,H) of the reference: one thread per atom
//                  writes a fixed-stride bond table (unit vector, length, fc, fc', pair type,
//                  neighbour, periodic shift) -- the reference's neb/bndnm/bndlen/cutfcnar arrays.
//   k_rebo2_force  loop 2: one thread per atom i, one undirected bond at a time (j_gt_i).  The
//                  nebmax^2 second-neighbour caches of the reference (nebofk, dncnidm, drk, ...)
//                  are not materialised: the conjugation terms of k's and l's neighbours are
//                  recomputed from the bond table when their forces are applied.
// Forces on i, j, k, l, m, n are accumulated with native FP64 atomicAdd (RED.ADD.F64); periodic
// image identity along neighbour paths is tracked with 3-vector shift sums.
#include <cstdlib>

#include "atx_potential_common.cuh"

#include "atx_rebo2_func.cuh"

#define RBS_ADD(p, v) atomicAdd((p), (v))
#define RBS_OR(p, v) atomicOr((p), (v))
#include "atx_rebo2_scr.cuh"
#include "atx_rebo2_atom.cuh"

struct atx_rebo2 {
  atx_ctx *ctx = nullptr;
  Rebo2Dev dev{};
  DevBuf<double> tables;
  bool bound = false;
  // bond table (stride nbs per atom)
  int nbs = 0;
  DevBuf<int> b_cnt, b_nb, b_typ, b_shift, b_slot;
  DevBuf<double4> b_vec;   // rnx, rny, rnz, rl
  DevBuf<double2> b_cut;   // fc, dfc
  DevBuf<RbBond> b_tab;    // unscreened path: the same fields as one 64-byte record per bond
  DevBuf<double2> nn;      // nn(C), nn(H)
  DevBuf<double> epb, fpb, wpb, epa_out;
  DevBuf<int> flag;
  PotScratch sc;
  // one-thread-per-bond force kernel (ATX_REBO2_PERBOND=1, experimental): compact list of the bonds
  // each atom is responsible for
  int per_bond = 3;   // 0: thread per atom; 1 / 2 / 3 / 4 / 5: thread per bond at 4 / 6 / 8 / 12 / 16 resident blocks
                      // per SM.  C3 on a B200 (round 2): 1.58 ms per atom, 1.05 / 0.92 / 0.885 ms at 4 / 6 / 8 blocks
  DevBuf<int> own_cnt, own_off;
  DevBuf<int2> own;
  // screened variant (Rebo2Scr): b_cut holds the attractive/repulsive cutoff, b_cbo / b_cnc the
  // bond-order and neighbour-count cutoffs, s_* the screening neighbours (stride nss per atom)
  bool screened = false;
  RbsCut scr{};
  int nss = 0;
  DevBuf<int> b_sseed, b_scnt, s_ent;
  DevBuf<double2> b_cbo, b_cnc;
  DevBuf<double> s_arik, s_arjk, s_boik, s_bojk, s_ncik, s_ncjk, s_facbo, s_facnc;
};

// ---- kernel 1: bond table + nn -------------------------------------------------------------

__global__ void k_rebo2_bonds(int nat, int nbs, Mat3 A, Rebo2Dev P, const double4 *__restrict__ pos4,
                              const long long *__restrict__ seed, const int2 *__restrict__ list,
                              int *__restrict__ b_cnt, RbBond *__restrict__ b_tab,
                              double2 *__restrict__ nn, int *__restrict__ flag,
                              const int *__restrict__ stop) {
  if (stop && *stop) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  rb_bonds_atom(nbs, A, P, pos4, seed, list, b_cnt, b_tab, nn, flag, s);
}

// ---- kernel 2: energies and forces ------------------------------------------------------------

#define RB_BLOCK 64

__global__ void __launch_bounds__(RB_BLOCK)
k_rebo2_force(int nat, int nbs, Rebo2Dev P, const long long *__restrict__ seed,
              const int *__restrict__ b_cnt, const RbBond *__restrict__ b_tab,
              const double2 *__restrict__ nn, const double4 *__restrict__ pos4,
              const int *__restrict__ order, double *__restrict__ f,
              double *__restrict__ epa, double *__restrict__ wpa, double *__restrict__ epb,
              double *__restrict__ fpb, double *__restrict__ wpb, double *__restrict__ partials,
              const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (RB_BLOCK / 32)];
  const int i = blockIdx.x * RB_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;

  rb_force_atom(nat, nbs, P, seed, b_cnt, b_tab, nn, pos4, order, f, epa,
                wpa, epb, fpb, wpb, i, acc);
  atx_block_sum<ATX_NSUM, RB_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

// ---- one thread per bond (default; ATX_REBO2_PERBOND=0 selects thread per atom) ----------------------------------------
// k_rebo2_force gives every thread the bonds its atom is responsible for: 0 .. 4 in amorphous carbon,
// 1.9 on average, so a warp waits for its busiest lane (lane utilisation ~ 47 %).  Here the
// responsible (atom, slot) pairs are compacted first (count, exclusive scan, fill) and every thread
// evaluates exactly one bond with the same per-atom source (rb_force_atom<ROLES, ONE>).

__global__ void k_rebo2_own_count(int nat, int nbs, Rebo2Dev P, const int *__restrict__ b_cnt,
                                  const RbBond *__restrict__ b_tab,
                                  const double4 *__restrict__ pos4, const int *__restrict__ order,
                                  int *__restrict__ cnt, const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nat) return;
  cnt[s] = s < nat ? rb_owned_bonds(nbs, P, b_cnt, b_tab, pos4, order, s, nullptr) : 0;
}

__global__ void k_rebo2_own_fill(int nat, int nbs, Rebo2Dev P, const int *__restrict__ b_cnt,
                                 const RbBond *__restrict__ b_tab,
                                 const double4 *__restrict__ pos4, const int *__restrict__ order,
                                 const int *__restrict__ off, int2 *__restrict__ own,
                                 const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  rb_owned_bonds(nbs, P, b_cnt, b_tab, pos4, order, s, own + off[s]);
}

// MINB: resident blocks per SM asked from the register allocator (4: 255 registers, 6: 170, 8: 128 with
// spills) -- occupancy against spills is to be measured (ATX_REBO2_PERBOND = 1 / 2 / 3)
template <int MINB, int NBL>
__global__ void __launch_bounds__(RB_BLOCK, MINB)
k_rebo2_force_bond(int nat, int nbs, Rebo2Dev P, const long long *__restrict__ seed,
                   const int *__restrict__ b_cnt, const RbBond *__restrict__ b_tab,
                   const double2 *__restrict__ nn, const double4 *__restrict__ pos4,
                   const int *__restrict__ order, double *__restrict__ f, double *__restrict__ epa,
                   double *__restrict__ wpa, double *__restrict__ epb, double *__restrict__ fpb,
                   double *__restrict__ wpb, double *__restrict__ partials, const int *__restrict__ off,
                   const int2 *__restrict__ own, const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (RB_BLOCK / 32)];
  const int t = blockIdx.x * RB_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (t < off[nat]) {
    const int2 e = own[t];
    rb_force_atom<false, true, NBL>(nat, nbs, P, seed, b_cnt, b_tab, nn, pos4, order,
                               f, epa, wpa, epb, fpb, wpb, e.x, acc, nullptr, e.y);
  }
  atx_block_sum<ATX_NSUM, RB_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

// Caller-supplied (LAMMPS-style) lists: same loop 2 with the energy / virial of a bond counted for
// its owned ends only, then the ghost rows are cleared (see rb_force_atom<ROLES>).
__global__ void __launch_bounds__(RB_BLOCK)
k_rebo2_force_roles(int nat, int nbs, Rebo2Dev P, const long long *__restrict__ seed,
                    const int *__restrict__ b_cnt, const RbBond *__restrict__ b_tab, const double2 *__restrict__ nn,
                    const double4 *__restrict__ pos4, const int *__restrict__ order,
                    double *__restrict__ f, double *__restrict__ epa, double *__restrict__ wpa,
                    double *__restrict__ partials, const unsigned char *__restrict__ role,
                    const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (RB_BLOCK / 32)];
  const int i = blockIdx.x * RB_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  rb_force_atom<true>(nat, nbs, P, seed, b_cnt, b_tab, nn, pos4, order, f,
                      epa, wpa, nullptr, nullptr, nullptr, i, acc, role);
  atx_block_sum<ATX_NSUM, RB_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

__global__ void k_rebo2_clear_ghosts(int nat, const unsigned char *__restrict__ role, double *__restrict__ f,
                                     double *__restrict__ epa, double *__restrict__ wpa,
                                     const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat || role[s] >= 2) return;
  f[3 * (size_t)s] = 0.0; f[3 * (size_t)s + 1] = 0.0; f[3 * (size_t)s + 2] = 0.0;
  if (epa) epa[s] = 0.0;
  if (wpa) {
#pragma unroll
    for (int q = 0; q < 9; q++) wpa[9 * (size_t)s + q] = 0.0;
  }
}

// ---- screened variant: thin kernels around the per-atom functions of atx_rebo2_scr.cuh ---------

__global__ void k_rbs_bonds(RbsTab T, Mat3 A, Rebo2Dev P, RbsCut S, const double4 *__restrict__ pos4,
                            const long long *__restrict__ seed, const int2 *__restrict__ list,
                            const int *__restrict__ stop) {
  if (stop && *stop) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= T.nat) return;
  rbs_bonds_atom(T, A, P, S, pos4, seed, list, s);
}

// partials layout [RBS_NSUM][nbtot]; this launch fills columns off .. off + gridDim.x
__global__ void __launch_bounds__(RB_BLOCK)
k_rbs_force(RbsTab T, Mat3 A, Rebo2Dev P, RbsCut S, const double4 *__restrict__ pos4,
            const long long *__restrict__ seed, const int2 *__restrict__ list,
            const int *__restrict__ order, double *__restrict__ f, double *__restrict__ epa,
            double *__restrict__ wpa, double *__restrict__ epb, double *__restrict__ fpb,
            double *__restrict__ wpb, double *__restrict__ partials, int nbtot, int off,
            const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (RB_BLOCK / 32)];
  const int i = blockIdx.x * RB_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (i < T.nat) rbs_force_atom(T, A, P, S, pos4, seed, list, order, f, epa, wpa, epb, fpb, wpb, i, acc);
  atx_block_sum<ATX_NSUM, RB_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * nbtot + off + blockIdx.x] = acc[k];
  }
}

__global__ void __launch_bounds__(RB_BLOCK)
k_rbs_scr(RbsTab T, Mat3 A, Rebo2Dev P, const double4 *__restrict__ pos4,
          const long long *__restrict__ seed, const int2 *__restrict__ list, double *__restrict__ f,
          double *__restrict__ wpa, double *__restrict__ wpb, double *__restrict__ partials, int nbtot,
          int off, const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (RB_BLOCK / 32)];
  const int i = blockIdx.x * RB_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (i < T.nat) rbs_scr_atom(T, A, P, pos4, seed, list, f, wpa, wpb, i, acc);
  atx_block_sum<ATX_NSUM, RB_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * nbtot + off + blockIdx.x] = acc[k];
  }
}

// ---------------------------------------------------------------------------

extern "C" int atx_rebo2_create(atx_ctx *ctx, const atx_rebo2_params *par, atx_rebo2 **out) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (!ctx || !par || !out) return ATX_ERROR_UNSPECIFIED;
  atx_rebo2 *pot = new atx_rebo2();
  pot->ctx = ctx;
  Rebo2Dev &D = pot->dev;
  rb_fill_dev(D, par);
  const size_t n3 = 144 * 64, n2 = 25 * 16;
  ATX_PASS(pot->tables.reserve(4 * n3 + 2 * n2));
  double *t = pot->tables.ptr;
  const double *src3[4] = {par->Fcc, par->Fch, par->Fhh, par->Tcc};
  for (int k = 0; k < 4; k++)
    ATX_CUDA(cudaMemcpy(t + k * n3, src3[k], sizeof(double) * n3, cudaMemcpyHostToDevice));
  ATX_CUDA(cudaMemcpy(t + 4 * n3, par->Pcc, sizeof(double) * n2, cudaMemcpyHostToDevice));
  ATX_CUDA(cudaMemcpy(t + 4 * n3 + n2, par->Pch, sizeof(double) * n2, cudaMemcpyHostToDevice));
  D.Fcc = t; D.Fch = t + n3; D.Fhh = t + 2 * n3; D.Tcc = t + 3 * n3;
  D.Pcc = t + 4 * n3; D.Pch = t + 4 * n3 + n2;
  for (int k = 0; k < 32; k++) D.el2typ[k] = 0;
  ATX_PASS(pot->flag.reserve(4));
  if (const char *v = getenv("ATX_REBO2_PERBOND")) pot->per_bond = atoi(v);
  // guarded (batched MD) steps never clear the flag: it has to start from zero
  ATX_CUDA(cudaMemset(pot->flag.ptr, 0, 4 * sizeof(int)));
  *out = pot;
  return 0;
}

// Rebo2Scr (rebo2_scr.f90, rebo2_db.f90:92-112, :170-253): the C-C cutoffs of the three families
// come from `scr`, all other pairs keep the inner cutoff in all three.
extern "C" int atx_rebo2_create_screened(atx_ctx *ctx, const atx_rebo2_params *par,
                                         const atx_rebo2_screening *scr, atx_rebo2 **out) {
  if (!ctx || !par || !scr || !out) return ATX_ERROR_UNSPECIFIED;
  if (!(scr->Cmax > scr->Cmin) || !(scr->Cmax > 1.0)) {
    atx_set_error("Rebo2Scr: need Cmax > Cmin and Cmax > 1.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_rebo2 *pot = nullptr;
  ATX_PASS(atx_rebo2_create(ctx, par, &pot));
  pot->screened = true;
  rbs_fill_cut(pot->scr, pot->dev, scr);
  *out = pot;
  return 0;
}

extern "C" int atx_rebo2_destroy(atx_rebo2 *pot) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  delete pot;
  return 0;
}

extern "C" int atx_rebo2_bind_to(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, int nel,
                                 const int *el2Z) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (nel > 31) {
    atx_set_error("Too many particle element ids.");
    return ATX_ERROR_UNSPECIFIED;
  }
  Rebo2Dev &D = pot->dev;
  for (int k = 0; k < 32; k++) D.el2typ[k] = 0;
  for (int k = 0; k < nel; k++) {
    if (el2Z[k] == 6) D.el2typ[k + 1] = RB_C;
    else if (el2Z[k] == 1) D.el2typ[k + 1] = RB_H;
  }
  // rebo2_module.f90:96-125
  if (nl) {
    // the screened variant needs every atom that can screen a bond of the longest cutoff
    // (rebo2_module.f90:96-125 with SCREENING: sqrt(C_dr_cut) * max cutoff)
    const double scr_range = pot->screened
        ? sqrt(pot->scr.max_cut_sq[RB_CC]) * (pot->scr.C_dr_cut > 1.0 ? sqrt(pot->scr.C_dr_cut) : 1.0) : 0.0;
    for (int i = 1; i <= nel; i++)
      for (int j = i; j <= nel; j++) {
        const int ti = D.el2typ[i], tj = D.el2typ[j];
        if (!ti || !tj) continue;
        const int pt = (ti == RB_C && tj == RB_C) ? RB_CC : ((ti == RB_H && tj == RB_H) ? RB_HH : RB_CH);
        double c = D.cut_h[pt];
        if (pt == RB_CC && scr_range > c) c = scr_range;
        ATX_PASS(atx_neighbors_request_interaction_range_pair(nl, c, i, j));
      }
  }
  pot->bound = true;
  return 0;
}

static int rebo2_scr_compute(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, const PotOut &o,
                             double *epb, double *fpb, double *wpb) {
  atx_ctx *ctx = pot->ctx;
  cudaStream_t st = ctx->stream;
  const int nat = nl->nat;
  int nbs = nl->nebmax < 1 ? 1 : nl->nebmax;
  if (nbs > RBS_NBL) nbs = RBS_NBL;
  if (pot->nss < 1) pot->nss = 32;
  // guarded (batched MD) steps cannot repeat a pass with a larger screening table: start wider; an
  // overflow raises the sticky flag that atx_rebo2_check_overflow reports after the batch
  if (o.stop && pot->nss < 64) pot->nss = 64;
  pot->nbs = nbs;
  const size_t nt = (size_t)nat * nbs + 1, ns = (size_t)nat * pot->nss + 1;
  ATX_PASS(pot->b_cnt.reserve(nat + 1));
  ATX_PASS(pot->b_nb.reserve(nt));
  ATX_PASS(pot->b_typ.reserve(nt));
  ATX_PASS(pot->b_shift.reserve(nt));
  ATX_PASS(pot->b_slot.reserve(nt));
  ATX_PASS(pot->b_sseed.reserve(nt));
  ATX_PASS(pot->b_scnt.reserve(nt));
  ATX_PASS(pot->b_vec.reserve(nt));
  ATX_PASS(pot->b_cut.reserve(nt));
  ATX_PASS(pot->b_cbo.reserve(nt));
  ATX_PASS(pot->b_cnc.reserve(nt));
  ATX_PASS(pot->nn.reserve(nat + 1));
  ATX_PASS(pot->s_ent.reserve(ns));
  ATX_PASS(pot->s_arik.reserve(ns));
  ATX_PASS(pot->s_arjk.reserve(ns));
  ATX_PASS(pot->s_boik.reserve(ns));
  ATX_PASS(pot->s_bojk.reserve(ns));
  ATX_PASS(pot->s_ncik.reserve(ns));
  ATX_PASS(pot->s_ncjk.reserve(ns));
  ATX_PASS(pot->s_facbo.reserve(ns));
  ATX_PASS(pot->s_facnc.reserve(ns));
  RbsTab T;
  T.nat = nat; T.nbs = nbs; T.nss = pot->nss;
  T.b_cnt = pot->b_cnt.ptr; T.b_nb = pot->b_nb.ptr; T.b_typ = pot->b_typ.ptr;
  T.b_shift = pot->b_shift.ptr; T.b_slot = pot->b_slot.ptr; T.b_sseed = pot->b_sseed.ptr;
  T.b_scnt = pot->b_scnt.ptr; T.b_vec = pot->b_vec.ptr; T.b_car = pot->b_cut.ptr;
  T.b_cbo = pot->b_cbo.ptr; T.b_cnc = pot->b_cnc.ptr; T.nn = pot->nn.ptr;
  T.s_ent = pot->s_ent.ptr; T.s_arik = pot->s_arik.ptr; T.s_arjk = pot->s_arjk.ptr;
  T.s_boik = pot->s_boik.ptr; T.s_bojk = pot->s_bojk.ptr; T.s_ncik = pot->s_ncik.ptr;
  T.s_ncjk = pot->s_ncjk.ptr; T.s_facbo = pot->s_facbo.ptr; T.s_facnc = pot->s_facnc.ptr;
  T.flag = pot->flag.ptr;
  int nblocks = (nat + RB_BLOCK - 1) / RB_BLOCK;
  if (nblocks < 1) nblocks = 1;
  const int nbtot = 2 * nblocks;  // loop 2 and loop 3 each contribute one column per block
  ATX_PASS(pot->sc.partials.reserve((size_t)nbtot * ATX_NSUM));
  if (!o.stop) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr, 0, sizeof(int), st));
  ATX_CUDA(cudaMemsetAsync(o.f, 0, sizeof(double) * 3 * (size_t)nat, st));
  if (o.epa) ATX_CUDA(cudaMemsetAsync(o.epa, 0, sizeof(double) * (size_t)nat, st));
  if (o.wpa) ATX_CUDA(cudaMemsetAsync(o.wpa, 0, sizeof(double) * 9 * (size_t)nat, st));
  if (nat > 0) {
    k_rbs_bonds<<<(nat + 127) / 128, 128, 0, st>>>(T, p->Abox, pot->dev, pot->scr, nl->pos4.ptr,
                                                   nl->seed.ptr, nl->list.ptr, o.stop);
    ATX_LAUNCHED();
  }
  {
    ProfScope ps_(ctx, "rebo2_scr_force");
    k_rbs_force<<<nblocks, RB_BLOCK, 0, st>>>(T, p->Abox, pot->dev, pot->scr, nl->pos4.ptr, nl->seed.ptr,
                                              nl->list.ptr, nl->order.ptr, o.f, o.epa, o.wpa, epb, fpb,
                                              wpb, pot->sc.partials.ptr, nbtot, 0, o.stop);
    ATX_LAUNCHED();
    k_rbs_scr<<<nblocks, RB_BLOCK, 0, st>>>(T, p->Abox, pot->dev, nl->pos4.ptr, nl->seed.ptr,
                                            nl->list.ptr, o.f, o.wpa, wpb, pot->sc.partials.ptr, nbtot,
                                            nblocks, o.stop);
    ATX_LAUNCHED();
  }
  if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nbtot, o.sums, o.stop));
  return 0;
}

static int rebo2_compute(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, const PotOut &o,
                         double *epb, double *fpb, double *wpb) {
  atx_ctx *ctx = pot->ctx;
  if (pot->screened) return rebo2_scr_compute(pot, p, nl, o, epb, fpb, wpb);
  cudaStream_t st = ctx->stream;
  int nat = nl->nat;
  int nbs = nl->nebmax < 1 ? 1 : nl->nebmax;
  if (nbs > RB_NBL) nbs = RB_NBL;
  pot->nbs = nbs;
  size_t nt = (size_t)nat * nbs + 1;
  ATX_PASS(pot->b_cnt.reserve(nat + 1));
  ATX_PASS(pot->b_tab.reserve(nt));
  ATX_PASS(pot->nn.reserve(nat + 1));
  int nblocks = (nat + RB_BLOCK - 1) / RB_BLOCK;
  if (nblocks < 1) nblocks = 1;
  ATX_PASS(pot->sc.partials.reserve((size_t)nblocks * ATX_NSUM));
  if (!o.stop) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr, 0, sizeof(int), st));
  // forces are accumulated with atomics: start from zero (guarded runs zero inside the MD driver)
  ATX_CUDA(cudaMemsetAsync(o.f, 0, sizeof(double) * 3 * (size_t)nat, st));
  if (o.epa) ATX_CUDA(cudaMemsetAsync(o.epa, 0, sizeof(double) * (size_t)nat, st));
  if (o.wpa) ATX_CUDA(cudaMemsetAsync(o.wpa, 0, sizeof(double) * 9 * (size_t)nat, st));
  if (nat > 0) {
    ProfScope ps_(ctx, "rebo2_bonds");
    k_rebo2_bonds<<<(nat + 127) / 128, 128, 0, st>>>(nat, nbs, p->Abox, pot->dev, nl->pos4.ptr,
                                                     nl->seed.ptr, nl->list.ptr, pot->b_cnt.ptr,
                                                     pot->b_tab.ptr,
                                                     pot->nn.ptr, pot->flag.ptr, o.stop);
    ATX_LAUNCHED();
  }
  if (o.role) {
    ProfScope ps_(ctx, "rebo2_force");
    k_rebo2_force_roles<<<nblocks, RB_BLOCK, 0, st>>>(
        nat, nbs, pot->dev, nl->seed.ptr, pot->b_cnt.ptr, pot->b_tab.ptr, pot->nn.ptr, nl->pos4.ptr, nl->order.ptr, o.f, o.epa,
        o.wpa, pot->sc.partials.ptr, o.role, o.stop);
    ATX_LAUNCHED();
    k_rebo2_clear_ghosts<<<(nat + 127) / 128, 128, 0, st>>>(nat, o.role, o.f, o.epa, o.wpa, o.stop);
    ATX_LAUNCHED();
  } else if (pot->per_bond && nat > 0) {
    ProfScope ps_(ctx, "rebo2_force");
    // every undirected bond has two directed table entries and one responsible end
    const size_t bound = (size_t)nat * nbs / 2 + 1;
    const int nbb = (int)((bound + RB_BLOCK - 1) / RB_BLOCK);
    ATX_PASS(pot->own_cnt.reserve(nat + 2));
    ATX_PASS(pot->own_off.reserve(nat + 2));
    ATX_PASS(pot->own.reserve(bound + 1));
    ATX_PASS(pot->sc.partials.reserve((size_t)nbb * ATX_NSUM));
    k_rebo2_own_count<<<(nat + 1 + 127) / 128, 128, 0, st>>>(nat, nbs, pot->dev, pot->b_cnt.ptr, pot->b_tab.ptr,
                                                             nl->pos4.ptr, nl->order.ptr, pot->own_cnt.ptr, o.stop);
    ATX_LAUNCHED();
    ATX_PASS(atx_scan_int(ctx, pot->own_cnt.ptr, pot->own_off.ptr, (size_t)nat + 1));
    k_rebo2_own_fill<<<(nat + 127) / 128, 128, 0, st>>>(nat, nbs, pot->dev, pot->b_cnt.ptr, pot->b_tab.ptr,
                                                        nl->pos4.ptr, nl->order.ptr, pot->own_off.ptr, pot->own.ptr,
                                                        o.stop);
    ATX_LAUNCHED();
#define RB_FORCE_BOND(MINB)                                                                                   \
  do {                                                                                                        \
  if (nbs <= 6)                                                                                               \
    k_rebo2_force_bond<MINB, 6><<<nbb, RB_BLOCK, 0, st>>>(                                                    \
      nat, nbs, pot->dev, nl->seed.ptr, pot->b_cnt.ptr, pot->b_tab.ptr, pot->nn.ptr, nl->pos4.ptr, nl->order.ptr, o.f, o.epa,  \
      o.wpa, epb, fpb, wpb, pot->sc.partials.ptr, pot->own_off.ptr, pot->own.ptr, o.stop);                    \
  else                                                                                                        \
  k_rebo2_force_bond<MINB, RB_NBL><<<nbb, RB_BLOCK, 0, st>>>(                                                         \
      nat, nbs, pot->dev, nl->seed.ptr, pot->b_cnt.ptr, pot->b_tab.ptr, pot->nn.ptr, nl->pos4.ptr, nl->order.ptr, o.f, o.epa,  \
      o.wpa, epb, fpb, wpb, pot->sc.partials.ptr, pot->own_off.ptr, pot->own.ptr, o.stop);                    \
  } while (0)
    if (pot->per_bond == 2) RB_FORCE_BOND(6);
    else if (pot->per_bond == 3) RB_FORCE_BOND(8);
    else if (pot->per_bond == 4) RB_FORCE_BOND(12);
    else if (pot->per_bond == 5) RB_FORCE_BOND(16);
    else RB_FORCE_BOND(4);
#undef RB_FORCE_BOND
    ATX_LAUNCHED();
    if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nbb, o.sums, o.stop));
    return 0;
  } else {
    ProfScope ps_(ctx, "rebo2_force");
    k_rebo2_force<<<nblocks, RB_BLOCK, 0, st>>>(nat, nbs, pot->dev, nl->seed.ptr, pot->b_cnt.ptr,
                                                pot->b_tab.ptr,
                                                pot->nn.ptr, nl->pos4.ptr, nl->order.ptr, o.f, o.epa, o.wpa, epb, fpb,
                                                wpb, pot->sc.partials.ptr, o.stop);
    ATX_LAUNCHED();
  }
  if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nblocks, o.sums, o.stop));
  return 0;
}

// overflow of the bond table during guarded (batched MD) steps; called by the MD driver after a sync
int atx_rebo2_check_overflow(atx_rebo2 *pot) {
  int h = 0;
  ATX_CUDA(cudaMemcpyAsync(&h, pot->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, pot->ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(pot->ctx->stream));
  if (h) {
    atx_set_error(h & 1 ? "Internal neighbor list exhausted, *nebmax* too small (bond table overflow during MD)."
                        : "Rebo2Scr: table of screening neighbours exhausted during MD.");
    return ATX_ERROR_UNSPECIFIED;
  }
  return 0;
}

int atx_rebo2_compute_device(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, const PotOut &o) {
  if (o.role && pot->screened) {
    atx_set_error("Rebo2Scr is not available under domain decomposition.");
    return ATX_ERROR_UNSPECIFIED;
  }
  return rebo2_compute(pot, p, nl, o, nullptr, nullptr, nullptr);
}

int atx_perbond_to_host(atx_ctx *ctx, atx_neighbors *nl, int ncomp, const double *dev_slots,
                        double *host, PinBuf<double> &stage);

extern "C" int atx_rebo2_energy_and_forces(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl,
                                           double *epot, double *f, double *wpot,
                                           double *epot_per_at, double *epot_per_bond,
                                           double *f_per_bond, double *wpot_per_at,
                                           double *wpot_per_bond) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (!pot->bound) {
    atx_set_error("bind_to has not been called on this potential.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_ctx *ctx = pot->ctx;
  if (nl->external && (pot->screened || epot_per_bond || f_per_bond || wpot_per_bond)) {
    atx_set_error("Rebo2 with an external neighbour list: the screened variant and per-bond outputs are not "
                  "available.");
    return ATX_ERROR_UNSPECIFIED;
  }
  ATX_PASS(atx_neighbors_update(nl, p));
  PotOut o;
  // caller-supplied list with explicit ghosts: needs the ghosts within 5 bond cutoffs of the owned
  // atoms and list rows for the ghosts
  if (nl->external) o.role = nl->role_ext.ptr;
  ATX_PASS(pot->sc.f.reserve(3 * (size_t)nl->nat + 3));
  ATX_PASS(pot->sc.sums.reserve(ATX_NSUM));
  o.f = pot->sc.f.ptr;
  o.sums = pot->sc.sums.ptr;
  if (epot_per_at) {
    ATX_PASS(pot->epa_out.reserve((size_t)nl->nat + 1));
    o.epa = pot->epa_out.ptr;
  }
  if (wpot_per_at) {
    ATX_PASS(pot->sc.wpa.reserve(9 * (size_t)nl->nat + 9));
    o.wpa = pot->sc.wpa.ptr;
  }
  size_t nslots = (size_t)nl->npairs + 1;
  double *epb = nullptr, *fpb = nullptr, *wpb = nullptr;
  if (epot_per_bond) {
    ATX_PASS(pot->epb.reserve(nslots));
    ATX_CUDA(cudaMemsetAsync(pot->epb.ptr, 0, sizeof(double) * nslots, ctx->stream));
    epb = pot->epb.ptr;
  }
  if (f_per_bond) {
    ATX_PASS(pot->fpb.reserve(3 * nslots));
    ATX_CUDA(cudaMemsetAsync(pot->fpb.ptr, 0, sizeof(double) * 3 * nslots, ctx->stream));
    fpb = pot->fpb.ptr;
  }
  if (wpot_per_bond) {
    ATX_PASS(pot->wpb.reserve(9 * nslots));
    ATX_CUDA(cudaMemsetAsync(pot->wpb.ptr, 0, sizeof(double) * 9 * nslots, ctx->stream));
    wpb = pot->wpb.ptr;
  }
  int h = 0;
  for (int attempt = 0;; attempt++) {
    ATX_PASS(rebo2_compute(pot, p, nl, o, epb, fpb, wpb));
    ATX_CUDA(cudaMemcpyAsync(&h, pot->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(ctx->stream));
    // screened variant: the per-atom table of screening neighbours was too small; the outputs of
    // the failed pass are overwritten (per-bond arrays are zeroed again) by the repeat
    if (pot->screened && (h & 2) && !(h & 1) && attempt < 4) {
      pot->nss *= 2;
      if (epb) ATX_CUDA(cudaMemsetAsync(epb, 0, sizeof(double) * nslots, ctx->stream));
      if (fpb) ATX_CUDA(cudaMemsetAsync(fpb, 0, sizeof(double) * 3 * nslots, ctx->stream));
      if (wpb) ATX_CUDA(cudaMemsetAsync(wpb, 0, sizeof(double) * 9 * nslots, ctx->stream));
      continue;
    }
    break;
  }
  if (h) {
    atx_set_error("Internal neighbor list exhausted, *nebmax* too small.");
    return ATX_ERROR_UNSPECIFIED;
  }
  if (epb) ATX_PASS(atx_perbond_to_host(ctx, nl, 1, epb, epot_per_bond, pot->sc.stage));
  if (fpb) ATX_PASS(atx_perbond_to_host(ctx, nl, 3, fpb, f_per_bond, pot->sc.stage));
  if (wpb) ATX_PASS(atx_perbond_to_host(ctx, nl, 9, wpb, wpot_per_bond, pot->sc.stage));
  return atx_finish_to_host(ctx, nl, pot->sc, o, epot, f, wpot, epot_per_at, wpot_per_at);
}

extern "C" int atx_rebo2_set_store_outputs(atx_rebo2 *pot, int on) {
  if (!pot) return ATX_ERROR_UNSPECIFIED;
  pot->sc.store_outputs = on != 0;
  return 0;
}
