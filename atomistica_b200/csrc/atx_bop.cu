// Tersoff / Kumagai / Brenner bond-order potentials on the device.
//
// Replaces BOP_KERNEL (src/potentials/bop/bop_kernel.f90:563-1630, SCREENING undefined) together
// with the potential-specific functions it inlines:
//   tersoff_func.f90:32-225, kumagai_func.f90:32-225, brenner_func.f90:27-215 (+ the derived
//   constants of brenner_module.f90:269-288) and trig_off_f (src/support/cutoff.f90:152-196).
//
// Scatter-free, deterministic formulation:
//   k_bop_center  one thread per central atom i.  Loop 1 of the reference (bond table: unit
//                 vector, length, cutoff value/derivative) is staged in shared memory, loop 2
//                 (per bond ij: zeta, bond order, pair terms) runs on top of it.  The force the
//                 bonds of i exert on each neighbour (the reference's "f(j) += fj", "f(k) += df")
//                 is accumulated PER LIST SLOT into G[slot] instead of being scattered.
//   k_bop_gather  one thread per atom: f_i = f_i(own) + sum over its slots a of G[rev[a]], where
//                 rev[a] is the slot of the reverse pair (j -> i, -shift) precomputed per list build.
// No floating-point atomics are used for forces and energies.
#include "atx_potential_common.cuh"

#define BOP_PI 3.14159265358979323846264338327950288

struct BopDev {
  int kind, nel;
  int el2db[32];  // particle element id -> db element (1..nel) or -1
  double r1[9], r2[9], r1sq[9], r2sq[9], cfac[9];
  // Tersoff: A B xi lambda mu omega mubo | beta n c d h (element)
  // Kumagai: A B lambda1(lambda) lambda2(mu) alpha(mubo) | eta delta c1..c5 h
  // Brenner: derived VR_f expR VA_f expA r0 gamma c_sq d_sq c_d h mu(mubo) n bo_exp bo_fac bo_exp1
  double A[9], B[9], xi[9], lambda[9], mu[9], omega[9], mubo[9];
  int m[9];
  double beta[3], n[3], c[3], d[3], h[3];
  double eta[3], delta[3], c1[3], c2[3], c3[3], c4[3], c5[3];
  double VR_f[9], expR[9], VA_f[9], expA[9], r0[9], gamma[9], c_sq[9], d_sq[9], c_d[9], ph[9],
      pn[9], bo_exp[9], bo_fac[9], bo_exp1[9];
  // precomputed Tersoff element constants
  // Juslin: triplet-indexed h() parameters (TRIPLET_INDEX_NS)
  double talpha[27], tomega[27];
  int tm[27];
  double tb[3];          // beta**n
  double te[3];          // -1/(2n)
  double c_sq_e[3], d_sq_e[3], one_p_c2d2[3];  // c^2, d^2, 1 + c^2/d^2
};

// exp_cutoff_t (src/support/cutoff.f90:232-293): the cutoff of every *_scr module
struct ExpCut {
  double r1, r2, fac1, fac2, c, d, off;
};

// Screened variants (SCREENING defined): cutoffs and Baskes screening bounds per pair,
// default_bind_to_func.f90:44-104
struct BopScrDev {   // 9 = nel**2 pairs of JuslinScr (6 for the symmetric pair index of the others)
  ExpCut cin[9], cout[9], cbo[9];
  double cut_in_l[9], cut_in_h[9], cut_in_h2[9], cut_out_l[9], cut_out_h[9], cut_bo_h[9], max_cut_sq[9];
  double Cmin[9], Cmax[9], dC[9], C_dr_cut[9];
  double screening_threshold, dot_threshold;
  int trig;   // JuslinScr: trigonometric switching functions (juslin_func.f90:27-122) instead of exp_cutoff_t
};

// per-atom tables of the screened kernels, bond-major / entry-major: element (b, s) at b*nat + s
struct BopScrTab {
  int nat, NB, NS;
  double *rnx, *rny, *rnz, *rl, *fcar, *dfcar, *fcbo, *dfcbo, *kx, *ky, *kz, *zf;  // [NB][nat]
  int *jidx, *slot, *typ, *sseed, *scnt;                                             // [NB][nat]
  int *nbond;                                                                        // [nat]
  double *arik, *arjk, *boik, *bojk, *sfac;                                          // [NS][nat]
  int *kslot;                                                                        // [NS][nat]
};

struct atx_bop {
  atx_ctx *ctx = nullptr;
  atx_bop_params par{};
  BopDev dev{};
  bool bound = false;
  DevBuf<double4> G;
  DevBuf<double> epb, fpb, wpb, epa_out;
  DevBuf<int> flag;   // [0] an atom exceeded BOP_NB_MAX bonds, [1] length of the queue, [8..40) bond histogram
  DevBuf<int> queue;  // atoms deferred to the queued pass
  PinBuf<int> hflag;
  int nb_cap = 0;     // bond-table capacity (template NB) in use
  const atx_neighbors *sized_nl = nullptr;  // list + build number nb_cap was sized for
  long long sized_build = -1;
  // screened variants
  bool screened = false;
  atx_juslin_screening scr_par{};   // 9 entries per row; the symmetric classes use the first 6
  DevBuf<BopScrDev> scr_dev;
  DevBuf<double> scr_d;   // backing store of the BopScrTab double fields
  DevBuf<int> scr_i;      // backing store of the BopScrTab int fields
  BopScrTab tab{};
  PotScratch sc;
};

__device__ __forceinline__ int bop_pair_index(int i, int j, int maxval) {
  // macros.inc:123, 1-based in, 0-based out
  int a = (i - 1) + (j - 1) * maxval, b = (j - 1) + (i - 1) * maxval;
  int c = (i - 1) * i / 2, d = (j - 1) * j / 2;
  return (a < b ? a : b) - (c < d ? c : d);
}

// pair index of the potential, 0-based: PAIR_INDEX for Tersoff/Kumagai/Brenner, PAIR_INDEX_NS
// (macros.inc:139; juslin_func.f90:300-313) for Juslin.  Elements 1-based.
__device__ __forceinline__ int bop_pidx(const BopDev &P, int i, int j) {
  if (P.kind == ATX_BOP_JUSLIN) return (j - 1) + (i - 1) * P.nel;
  return bop_pair_index(i, j, P.nel);
}

template <int KIND>
__device__ __forceinline__ void bop_VA(const BopDev &P, int ij, double dr, double &val, double &dval) {
  if (KIND == ATX_BOP_BRENNER || KIND == ATX_BOP_JUSLIN) {
    double e = exp(-P.expA[ij] * (dr - P.r0[ij]));
    val = -P.VA_f[ij] * e;
    dval = P.VA_f[ij] * P.expA[ij] * e;
  } else {
    double e = exp(-P.mu[ij] * dr);
    val = -P.B[ij] * e;
    dval = P.B[ij] * P.mu[ij] * e;
  }
}

template <int KIND>
__device__ __forceinline__ void bop_VR(const BopDev &P, int ij, double dr, double &val, double &dval) {
  if (KIND == ATX_BOP_BRENNER || KIND == ATX_BOP_JUSLIN) {
    double e = exp(-P.expR[ij] * (dr - P.r0[ij]));
    val = P.VR_f[ij] * e;
    dval = -P.VR_f[ij] * P.expR[ij] * e;
  } else {
    double e = exp(-P.lambda[ij] * dr);
    val = P.A[ij] * e;
    dval = -P.A[ij] * P.lambda[ij] * e;
  }
}

template <int KIND>
__device__ __forceinline__ void bop_g(const BopDev &P, int ti, int ik, double costh, double &val,
                                      double &dval) {
  if (KIND == ATX_BOP_TERSOFF) {
    const double omega = P.omega[ik];
    const double h_c = P.h[ti] - costh;
    const double c_sq = P.c_sq_e[ti];
    const double inv_h = 1.0 / (P.d_sq_e[ti] + h_c * h_c);
    val = omega * (P.one_p_c2d2[ti] - c_sq * inv_h);
    dval = -2 * omega * c_sq * h_c * inv_h * inv_h;
  } else if (KIND == ATX_BOP_KUMAGAI) {
    double h_cos = P.h[ti] - costh;
    double h_cos_sq = h_cos * h_cos;
    double tmp = h_cos / (P.c3[ti] + h_cos_sq);
    double go = P.c2[ti] * tmp;
    double ga1 = P.c4[ti] * exp(-P.c5[ti] * h_cos_sq);
    double v = go * (1.0 + ga1);
    dval = -2 * (1.0 - h_cos * tmp) * v + 2 * P.c5[ti] * h_cos_sq * go * ga1;
    val = P.c1[ti] + h_cos * v;
  } else {
    const double hc = P.ph[ik] + costh;
    const double inv_h = 1.0 / (P.d_sq[ik] + hc * hc);
    val = P.gamma[ik] * (1 + P.c_d[ik] - P.c_sq[ik] * inv_h);
    dval = 2 * P.gamma[ik] * P.c_sq[ik] * hc * inv_h * inv_h;
  }
}

template <int KIND>
__device__ __forceinline__ void bop_bo(const BopDev &P, int ti, int ij, double zij, double fcij,
                                       double faij, double &bij, double &dfbij) {
  // z**(n-1) and arg**(e-1) are formed as z**n / z and arg**e / arg: two pow() instead of four
  if (KIND == ATX_BOP_TERSOFF) {
    if (zij > 0.0) {
      const double n = P.n[ti], e = P.te[ti], b = P.tb[ti];
      const double zn = pow(zij, n);
      const double arg = 1.0 + b * zn;
      const double ae = pow(arg, e);
      bij = P.xi[ij] * ae;
      dfbij = -0.25 * fcij * faij * P.xi[ij] * b * (zn / zij) * (ae / arg);
    } else {
      bij = 1.0;
      dfbij = 0.0;
    }
  } else if (KIND == ATX_BOP_KUMAGAI) {
    if (zij > 0.0) {
      const double eta = P.eta[ti], delta = -P.delta[ti];
      const double zn = eta == 1.0 ? zij : pow(zij, eta);
      const double arg = 1.0 + zn;
      const double ad = pow(arg, delta);
      bij = ad;
      dfbij = 0.5 * fcij * faij * eta * (zn / zij) * delta * (ad / arg);
    } else {
      bij = 1.0;
      dfbij = 0.0;
    }
  } else {
    if (P.pn[ij] == 1.0) {
      const double arg = 1.0 + zij;
      const double ae = pow(arg, P.bo_exp[ij]);
      bij = ae;
      dfbij = P.bo_fac[ij] * fcij * faij * (ae / arg);
    } else if (zij > 0.0) {
      const double zn = pow(zij, P.pn[ij]);
      const double arg = 1.0 + zn;
      const double ae = pow(arg, P.bo_exp[ij]);
      bij = ae;
      dfbij = P.bo_fac[ij] * fcij * faij * (zn / zij) * (ae / arg);
    } else {
      bij = 1.0;
      dfbij = 0.0;
    }
  }
}

template <int KIND>
__device__ __forceinline__ void bop_h(const BopDev &P, int ti, int ij, int ik, double dr, double &val,
                                      double &dval) {
  if (KIND == ATX_BOP_JUSLIN) {
    // juslin_func.f90:255-294: with the non-symmetric pair index the partner element is ij % nel
    const int t = (ik % P.nel) + P.nel * ((ij % P.nel) + P.nel * ti);
    const double alpha = P.talpha[t], omega = P.tomega[t];
    const int m = P.tm[t];
    if (m == 1) {
      val = omega * exp(alpha * dr);
      dval = alpha * val;
    } else if (m == 3) {
      const double arg = alpha * dr;
      val = omega * exp(arg * arg * arg);
      dval = 3 * alpha * arg * arg * val;
    } else {
      const double arg = alpha * dr;
      val = omega * exp(pow(arg, (double)m));
      dval = m * pow(arg, (double)(m - 1)) * alpha * val;
    }
    return;
  }
  double mu = P.mubo[ik];
  if (mu == 0.0) {
    val = 1.0;
    dval = 0.0;
    return;
  }
  int m = P.m[ik];
  if (KIND == ATX_BOP_KUMAGAI) {
    if (m == 1) {
      val = exp(mu * dr);
      dval = mu * val;
    } else if (m == 3) {
      val = exp(dr * dr * dr);  // sic: kumagai_func.f90:211-213 drops alpha
      dval = 3 * mu * dr * dr * val;
    } else {
      val = exp(mu * pow(dr, (double)m));
      dval = m * mu * pow(dr, (double)(m - 1)) * val;
    }
  } else {
    if (m == 1) {
      val = exp(2 * mu * dr);
      dval = 2 * mu * val;
    } else if (m == 3) {
      double arg = 2 * mu * dr;
      val = exp(arg * arg * arg);
      dval = 2 * mu * m * arg * arg * val;
    } else {
      val = exp(pow(2 * mu * dr, (double)m));
      dval = 2 * mu * m * pow(2 * mu * dr, (double)(m - 1)) * val;
    }
  }
}

#define BOP_BLOCK 64
#define BOP_NB_MAX 24   // deepest bond table (queued pass); more bonds on one atom is an error

// Histogram of bonds (list entries inside the potential's r2) per atom -> hist[0..31] (last bin =
// 31 or more).  Sizes the shared-memory bond table by BONDS, not by list entries: with a Verlet
// shell the list of an atom can hold several times more entries than it has bonds.
__global__ void k_bop_count_bonds(int nat, Mat3 A, BopDev P, const double4 *__restrict__ pos4,
                                  const long long *__restrict__ seed, const int2 *__restrict__ list,
                                  const unsigned char *__restrict__ role, int *__restrict__ hist) {
  __shared__ int sh[32];
  if (threadIdx.x < 32) sh[threadIdx.x] = 0;
  __syncthreads();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nat && (!role || role[s] >= 1)) {
    int nb = 0;
    double4 pi = pos4[s];
    const int eli = P.el2db[(int)pi.w];
    if (eli > 0)
      for (long long a = seed[s]; a < seed[s + 1]; a++) {
        int2 en = list[a];
        int elj = P.el2db[ATX_ENTRY_EL(en.y)];
        if (elj <= 0) continue;
        double4 pj = pos4[en.x];
        double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        if (ATX_NONZERO_SHIFT(en.y)) {
          int sx, sy, sz;
          atx_unpack_shift(en.y, sx, sy, sz);
          double ax, ay, az;
          atx_image_vector(A, sx, sy, sz, ax, ay, az);
          dx -= ax; dy -= ay; dz -= az;
        }
        if (dx * dx + dy * dy + dz * dz < P.r2sq[bop_pidx(P, eli, elj)]) nb++;
      }
    atomicAdd(&sh[nb < 31 ? nb : 31], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// shared-memory bond table, field-major so that consecutive threads hit consecutive banks
template <int NB, bool LEAN = false>
struct BondSmem {
  double rnx[NB][BOP_BLOCK], rny[NB][BOP_BLOCK], rnz[NB][BOP_BLOCK];
  double rl[NB][BOP_BLOCK], ri[NB][BOP_BLOCK], fc[NB][BOP_BLOCK], dfc[NB][BOP_BLOCK];
  double kx[NB][BOP_BLOCK], ky[NB][BOP_BLOCK], kz[NB][BOP_BLOCK];  // dbidk of the current ij
  double gx[NB][BOP_BLOCK], gy[NB][BOP_BLOCK], gz[NB][BOP_BLOCK], ge[NB][BOP_BLOCK];  // per-slot G
  int slot[NB][BOP_BLOCK];
  int typ[NB][BOP_BLOCK];
  double red[ATX_NSUM * (BOP_BLOCK / 32)];
};
// MD steps whose energy nobody reads (no mask, no per-atom / per-bond outputs, no virial): no per-slot
// energy, pair type and list slot in one word -- 108 instead of 120 bytes per bond and thread, which is
// what lets an eighth block of the depth-4 kernel fit on an SM
template <int NB>
struct BondSmem<NB, true> {
  double rnx[NB][BOP_BLOCK], rny[NB][BOP_BLOCK], rnz[NB][BOP_BLOCK];
  double rl[NB][BOP_BLOCK], ri[NB][BOP_BLOCK], fc[NB][BOP_BLOCK], dfc[NB][BOP_BLOCK];
  double kx[NB][BOP_BLOCK], ky[NB][BOP_BLOCK], kz[NB][BOP_BLOCK];
  double gx[NB][BOP_BLOCK], gy[NB][BOP_BLOCK], gz[NB][BOP_BLOCK];
  int st[NB][BOP_BLOCK];   // pair type | list slot << 4
  double red[ATX_NSUM * (BOP_BLOCK / 32)];
};

// One centre atom: bond table in shared memory, then all bonds ij with their k-sums.  Returns false
// (having written nothing but zeroed G slots) when the atom has more than NB bonds.
template <int KIND, int NB, bool VIRIAL, bool LEAN = false>
__device__ __forceinline__ bool bop_center_atom(
    BondSmem<NB, LEAN> &S, const int t, const int s, const Mat3 &A, const BopDev &P,
    const double4 *__restrict__ pos4, const long long *__restrict__ seed, const int2 *__restrict__ list,
    const int *__restrict__ mask, double4 *__restrict__ G, double *__restrict__ f,
    double *__restrict__ pe_own, double *__restrict__ wpa, double *__restrict__ epb,
    double *__restrict__ fpb, double *__restrict__ wpb, double (&acc)[ATX_NSUM], const bool own) {
  // own: the centre belongs to this process / rank; bonds of ghost centres still produce the forces
  // they exert on owned atoms, but their energy and virial are counted by their owner
  double4 pi = pos4[s];
  const int eli = P.el2db[(int)pi.w];
  const long long b0 = seed[s], b1 = seed[s + 1];
  int nb = 0;
  bool ovf = false;
  // ---- loop 1: bond table (bop_kernel.f90:563-1068) ----
  for (long long a = b0; a < b1; a++) {
    int2 en = list[a];
    bool bond = false;
    if (eli > 0) {
      double4 pj = pos4[en.x];
      int elj = P.el2db[(int)pj.w];
      if (elj > 0) {
        double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        if (ATX_NONZERO_SHIFT(en.y)) {
          int sx, sy, sz;
          atx_unpack_shift(en.y, sx, sy, sz);
          double ax, ay, az;
          atx_image_vector(A, sx, sy, sz, ax, ay, az);
          dx -= ax; dy -= ay; dz -= az;
        }
        double r2 = dx * dx + dy * dy + dz * dz;
        int ij = bop_pidx(P, eli, elj);
        if (r2 < P.r2sq[ij]) {
          if (nb >= NB) {
            ovf = true;
          } else {
            double rl = sqrt(r2);
            double fc = 1.0, dfc = 0.0;
            if (!(r2 < P.r1sq[ij])) {
              // trig_off_f
              if (rl <= P.r1[ij]) { fc = 1.0; dfc = 0.0; }
              else if (rl >= P.r2[ij]) { fc = 0.0; dfc = 0.0; }
              else {
                double sn, cs;
                sincos(P.cfac[ij] * (rl - P.r1[ij]), &sn, &cs);
                fc = 0.5 * (1.0 + cs);
                dfc = -0.5 * P.cfac[ij] * sn;
              }
            }
            const double ri = 1.0 / rl;
            S.rnx[nb][t] = dx * ri; S.rny[nb][t] = dy * ri; S.rnz[nb][t] = dz * ri;
            S.rl[nb][t] = rl; S.ri[nb][t] = ri; S.fc[nb][t] = fc; S.dfc[nb][t] = dfc;
            S.gx[nb][t] = 0.0; S.gy[nb][t] = 0.0; S.gz[nb][t] = 0.0;
            if constexpr (LEAN) {
              S.st[nb][t] = ij | ((int)(a - b0) << 4);
            } else {
              S.ge[nb][t] = 0.0;
              S.slot[nb][t] = (int)(a - b0);
              S.typ[nb][t] = ij | (en.x << 4);
            }
            nb++;
            bond = true;
          }
        }
      }
    }
    if (!bond) G[a] = make_double4(0.0, 0.0, 0.0, 0.0);
  }

  if (ovf) return false;  // more bonds than this instantiation holds: the caller defers the atom

  // ---- loop 2 (bop_kernel.f90:1075-1529) ----
  double fix = 0.0, fiy = 0.0, fiz = 0.0, pei = 0.0;
  const int mi = mask ? mask[s] : 1;
  for (int ij = 0; ij < nb; ij++) {
    int tij, j = 0, maskfac = 2;
    if constexpr (LEAN) {
      tij = S.st[ij][t] & 15;
    } else {
      tij = S.typ[ij][t] & 15;
      j = S.typ[ij][t] >> 4;
      if (mask) {
        int mj = mask[j];
        if (mi == 0 && mj == 0) maskfac = 0;
        else if (mi == 0 || mj == 0) maskfac = 1;
      }
    }
    const double rlij = S.rl[ij][t];
    if (!(maskfac > 0 && rlij < P.r2[tij])) continue;
    const double rlijr = S.ri[ij][t];
    const double nx = S.rnx[ij][t], ny = S.rny[ij][t], nz = S.rnz[ij][t];
    const double rijx = rlij * nx, rijy = rlij * ny, rijz = rlij * nz;
    const double fcarij = S.fc[ij][t], dfcarijr = S.dfc[ij][t];
    double VAij, dVAij, VRij, dVRij;
    bop_VA<KIND>(P, tij, rlij, VAij, dVAij);
    bop_VR<KIND>(P, tij, rlij, VRij, dVRij);
    const double mf = 0.5 * maskfac;
    VAij *= mf; dVAij *= mf; VRij *= mf; dVRij *= mf;

    double zij = 0.0;
    double dix = 0, diy = 0, diz = 0, djx = 0, djy = 0, djz = 0;
    double wb[9];
    if (VIRIAL) {
#pragma unroll
      for (int q = 0; q < 9; q++) wb[q] = 0.0;
    }

    for (int ik = 0; ik < nb; ik++) {
      if (ik == ij) continue;
      int tik;
      if constexpr (LEAN) tik = S.st[ik][t] & 15; else tik = S.typ[ik][t] & 15;
      const double rlik = S.rl[ik][t];
      if (!(rlik < P.r2[tik])) {
        S.kx[ik][t] = 0.0; S.ky[ik][t] = 0.0; S.kz[ik][t] = 0.0;
        continue;
      }
      const double kx = S.rnx[ik][t], ky = S.rny[ik][t], kz = S.rnz[ik][t];
      const double fcik = S.fc[ik][t], dfcikr = S.dfc[ik][t];
      double h_Dr, dh_dDr, g_costh, dg_dcosth;
      bop_h<KIND>(P, eli - 1, tij, tik, rlij - rlik, h_Dr, dh_dDr);
      const double costh = kx * nx + ky * ny + kz * nz;
      bop_g<KIND>(P, eli - 1, tik, costh, g_costh, dg_dcosth);
      double ex = kx * rlik - nx * rlij, ey = ky * rlik - ny * rlij, ez = kz * rlik - nz * rlij;
      const double disjk = sqrt(ex * ex + ey * ey + ez * ez);
      const double idis = 1.0 / disjk, rlikr = S.ri[ik][t];
      ex *= idis; ey *= idis; ez *= idis;
      const double dcsdij = rlikr - costh * rlijr;
      const double dcsdik = rlijr - costh * rlikr;
      const double dcsdjk = -disjk * rlijr * rlikr;
      const double dzfac = fcik * dg_dcosth * h_Dr;
      zij += fcik * g_costh * h_Dr;
      const double dzdrij = g_costh * fcik * dh_dDr;
      const double dzdrik = g_costh * (dfcikr * h_Dr - fcik * dh_dDr);
      // x
      double dfx, dfy, dfz, dkx, dky, dkz;
      {
        double ci = -dcsdij * nx - dcsdik * kx, cj = dcsdij * nx - dcsdjk * ex, ck = dcsdik * kx + dcsdjk * ex;
        dix += -dzdrij * nx - dzdrik * kx + dzfac * ci;
        dfx = dzdrij * nx + dzfac * cj;
        dkx = dzdrik * kx + dzfac * ck;
      }
      {
        double ci = -dcsdij * ny - dcsdik * ky, cj = dcsdij * ny - dcsdjk * ey, ck = dcsdik * ky + dcsdjk * ey;
        diy += -dzdrij * ny - dzdrik * ky + dzfac * ci;
        dfy = dzdrij * ny + dzfac * cj;
        dky = dzdrik * ky + dzfac * ck;
      }
      {
        double ci = -dcsdij * nz - dcsdik * kz, cj = dcsdij * nz - dcsdjk * ez, ck = dcsdik * kz + dcsdjk * ez;
        diz += -dzdrij * nz - dzdrik * kz + dzfac * ci;
        dfz = dzdrij * nz + dzfac * cj;
        dkz = dzdrik * kz + dzfac * ck;
      }
      djx += dfx; djy += dfy; djz += dfz;
      S.kx[ik][t] = dkx; S.ky[ik][t] = dky; S.kz[ik][t] = dkz;
      if (VIRIAL) {
        const double rikx = rlik * kx, riky = rlik * ky, rikz = rlik * kz;
        // wijb(a,b) -= rij(a)*df(b) + rik(a)*dbidk(b); column-major index a + 3b
        wb[0] -= rijx * dfx + rikx * dkx; wb[1] -= rijy * dfx + riky * dkx; wb[2] -= rijz * dfx + rikz * dkx;
        wb[3] -= rijx * dfy + rikx * dky; wb[4] -= rijy * dfy + riky * dky; wb[5] -= rijz * dfy + rikz * dky;
        wb[6] -= rijx * dfz + rikx * dkz; wb[7] -= rijy * dfz + riky * dkz; wb[8] -= rijz * dfz + rikz * dkz;
      }
    }

    double bij, dfb;
    bop_bo<KIND>(P, eli - 1, tij, zij, fcarij, VAij, bij, dfb);
    const double e_bond = 0.5 * fcarij * (VRij + bij * VAij);
    pei += e_bond;
    if constexpr (!LEAN) S.ge[ij][t] += e_bond;
    const double dffac = 0.5 * (dVRij * fcarij + bij * dVAij * fcarij + VRij * dfcarijr + bij * VAij * dfcarijr);
    const double dfx = dffac * nx, dfy = dffac * ny, dfz = dffac * nz;
    fix += dfx - dfb * dix; fiy += dfy - dfb * diy; fiz += dfz - dfb * diz;
    S.gx[ij][t] += -dfx - dfb * djx; S.gy[ij][t] += -dfy - dfb * djy; S.gz[ij][t] += -dfz - dfb * djz;
    for (int ik = 0; ik < nb; ik++) {
      if (ik == ij) continue;
      S.gx[ik][t] -= dfb * S.kx[ik][t]; S.gy[ik][t] -= dfb * S.ky[ik][t]; S.gz[ik][t] -= dfb * S.kz[ik][t];
    }
    if (own) acc[0] += e_bond;
    long long a;
    if constexpr (LEAN) a = b0 + (S.st[ij][t] >> 4); else a = b0 + S.slot[ij][t];
    if (epb) epb[a] = e_bond;
    if (fpb) { fpb[3 * a] = dfx; fpb[3 * a + 1] = dfy; fpb[3 * a + 2] = dfz; }
    if (VIRIAL) {
      double w[9];
      w[0] = rijx * dfx - dfb * wb[0]; w[1] = rijy * dfx - dfb * wb[1]; w[2] = rijz * dfx - dfb * wb[2];
      w[3] = rijx * dfy - dfb * wb[3]; w[4] = rijy * dfy - dfb * wb[4]; w[5] = rijz * dfy - dfb * wb[5];
      w[6] = rijx * dfz - dfb * wb[6]; w[7] = rijy * dfz - dfb * wb[7]; w[8] = rijz * dfz - dfb * wb[8];
      if (own) {
#pragma unroll
        for (int q = 0; q < 9; q++) acc[1 + q] += w[q];
      }
      if (wpb) {
#pragma unroll
        for (int q = 0; q < 9; q++) wpb[9 * a + q] = w[q];
      }
      if (wpa) {
        // optional analysis output (not on the hot path): per-atom virial, half to i, half to j
#pragma unroll
        for (int q = 0; q < 9; q++) {
          atomicAdd(&wpa[9 * (size_t)s + q], 0.5 * w[q]);
          atomicAdd(&wpa[9 * (size_t)j + q], 0.5 * w[q]);
        }
      }
    }
  }
  f[3 * s] = fix; f[3 * s + 1] = fiy; f[3 * s + 2] = fiz;
  pe_own[s] = pei;
  for (int k = 0; k < nb; k++) {
    if constexpr (LEAN) G[b0 + (S.st[k][t] >> 4)] = make_double4(S.gx[k][t], S.gy[k][t], S.gz[k][t], 0.0);
    else G[b0 + S.slot[k][t]] = make_double4(S.gx[k][t], S.gy[k][t], S.gz[k][t], S.ge[k][t]);
  }
  return true;
}

// Main pass: one thread per atom with an NB-deep bond table; atoms with more bonds are appended to
// `queue` and handled by k_bop_center_queued with the deepest table, so NB can be sized for the
// typical atom instead of the worst one (shared memory per thread = 120 B x NB sets the occupancy).
template <int KIND, int NB, int MINB, bool VIRIAL, bool LEAN = false>
__global__ void __launch_bounds__(BOP_BLOCK, MINB)
k_bop_center(int nat, Mat3 A, BopDev P, const double4 *__restrict__ pos4,
             const long long *__restrict__ seed, const int2 *__restrict__ list,
             const int *__restrict__ mask, double4 *__restrict__ G, double *__restrict__ f,
             double *__restrict__ pe_own, double *__restrict__ wpa, double *__restrict__ epb,
             double *__restrict__ fpb, double *__restrict__ wpb, double *__restrict__ partials,
             int pstride, int *__restrict__ queue, int *__restrict__ qcount,
             const unsigned char *__restrict__ role, const int *__restrict__ stop, int boff) {
  if (stop && *stop) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BondSmem<NB, LEAN> &S = *reinterpret_cast<BondSmem<NB, LEAN> *>(smem_raw);
  const int t = threadIdx.x;
  const int blk = blockIdx.x + boff;     // the launch covers the blocks [boff, boff + gridDim.x)
  const int s = blk * BOP_BLOCK + t;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (s < nat && (!role || role[s] >= 1)) {
    if (!bop_center_atom<KIND, NB, VIRIAL, LEAN>(S, t, s, A, P, pos4, seed, list, mask, G, f, pe_own, wpa, epb, fpb,
                                                 wpb, acc, !role || role[s] >= 2))
      queue[atomicAdd(qcount, 1)] = s;
  }
  atx_block_sum<ATX_NSUM, BOP_BLOCK>(acc, S.red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * pstride + blk] = acc[k];
  }
}

// Deferred atoms, deepest bond table (1 block per SM).  Always launched; exits at once when the
// queue is empty.  Its block partials go behind those of the main pass.
template <int KIND>
__global__ void __launch_bounds__(BOP_BLOCK)
k_bop_center_queued(Mat3 A, BopDev P, const double4 *__restrict__ pos4,
                    const long long *__restrict__ seed, const int2 *__restrict__ list,
                    const int *__restrict__ mask, double4 *__restrict__ G, double *__restrict__ f,
                    double *__restrict__ pe_own, double *__restrict__ wpa, double *__restrict__ epb,
                    double *__restrict__ fpb, double *__restrict__ wpb, double *__restrict__ partials,
                    int pstride, int pofs, const int *__restrict__ queue,
                    const int *__restrict__ qcount, int *__restrict__ flag,
                    const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BondSmem<BOP_NB_MAX> &S = *reinterpret_cast<BondSmem<BOP_NB_MAX> *>(smem_raw);
  const int t = threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  const int nq = *qcount;
  for (int q = blockIdx.x * BOP_BLOCK + t; q < nq; q += gridDim.x * BOP_BLOCK) {
    const int sq = queue[q];
    if (!bop_center_atom<KIND, BOP_NB_MAX, true>(S, t, sq, A, P, pos4, seed, list, mask, G, f, pe_own, wpa, epb,
                                                 fpb, wpb, acc, !role || role[sq] >= 2))
      atomicOr(flag, 1);  // more than BOP_NB_MAX bonds on one atom: reported as an error by the host
  }
  atx_block_sum<ATX_NSUM, BOP_BLOCK>(acc, S.red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * pstride + pofs + blockIdx.x] = acc[k];
  }
}

// ===========================================================================================
// Screened variants (TersoffScr, KumagaiScr, BrennerScr): BOP_KERNEL compiled with SCREENING
// (bop_kernel.f90:563-1068 incl. the screening function 682-995; 1075-1529 incl. 1448-1520;
// 1531-1611), CUTOFF_T = exp_cutoff_t.  Two kernels, one thread per central atom:
//   k_bopscr_bonds   loop 1 of the reference: per list entry decide unscreened / screened /
//                    partially screened, evaluate S_ij = exp(-sum ((Cmax-C)/(C-Cmin))^2) over the
//                    screening neighbours k (taken from the SAME list of i), and write the bond
//                    table + screening-neighbour table of the atom to global memory.  COUNT mode
//                    only measures the table depths (called once per list build).
//   k_bopscr_center  loops 2 and 3 for the atom.  sfacbo of the bonds of i only receives
//                    contributions from i's own ij loop, so loop 3 can follow loop 2 in the same
//                    thread.  Every atom that receives force from centre i (j, k, screening
//                    neighbours) is an entry of i's list: forces are accumulated per list slot
//                    in G and collected by k_bop_gather as in the unscreened path.
// ===========================================================================================

__device__ __forceinline__ void bop_expcut(const ExpCut &t, double r, double &val, double &dval) {
  if (r <= t.r1) { val = 1.0; dval = 0.0; }
  else if (r >= t.r2) { val = 0.0; dval = 0.0; }
  else {
    const double x = t.fac1 * (r - t.r1);
    const double x2 = x * x;
    const double v = exp(-8 * x * x2);
    const double dv = -24 * x2 * v;
    dval = t.fac1 * t.fac2 * (dv + 3 * t.c * x2 + 4 * t.d * x * x2);
    val = t.fac2 * (v + t.c * x * x2 + t.d * x2 * x2 - t.off);
  }
}

// the three cutoffs of the screened kernels: exp_cutoff_t, or the cosine switch of JuslinScr
// (fCin / fCar / fCbo, juslin_func.f90:27-122: 0.5 (1 + cos(pi (r - r1) / (r2 - r1))))
__device__ __forceinline__ void bop_scrcut(int trig, const ExpCut &t, double r, double &val, double &dval) {
  if (!trig) {
    bop_expcut(t, r, val, dval);
  } else if (r > t.r2) {
    val = 0.0; dval = 0.0;
  } else if (r < t.r1) {
    val = 1.0; dval = 0.0;
  } else {
    const double fca = BOP_PI / (t.r2 - t.r1), fc = -0.5 * fca;
    const double arg = fca * (r - t.r1);
    val = 0.5 * (1.0 + cos(arg));
    dval = fc * sin(arg);
  }
}

__device__ __forceinline__ void bop_list_vec(const Mat3 &A, const double4 &pi, const double4 &pj, int packed,
                                             double &dx, double &dy, double &dz) {
  // r_j - r_i - Abox.dc
  dx = pj.x - pi.x; dy = pj.y - pi.y; dz = pj.z - pi.z;
  if (ATX_NONZERO_SHIFT(packed)) {
    int sx, sy, sz;
    atx_unpack_shift(packed, sx, sy, sz);
    double ax, ay, az;
    atx_image_vector(A, sx, sy, sz, ax, ay, az);
    dx -= ax; dy -= ay; dz -= az;
  }
}

#define TB(field, b) T.field[(size_t)(b) * T.nat + s]

template <bool COUNT>
__global__ void __launch_bounds__(128)
k_bopscr_bonds(int nat, Mat3 A, BopDev P, const BopScrDev *__restrict__ Sp, BopScrTab T,
               const double4 *__restrict__ pos4, const long long *__restrict__ seed,
               const int2 *__restrict__ list, int *__restrict__ flag,
               const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int nb = 0, ns = 0;
  if (s < nat && (!role || role[s] >= 1)) {
    const BopScrDev &S = *Sp;
    const double4 pi = pos4[s];
    const int eli = P.el2db[(int)pi.w];
    const long long b0 = seed[s], b1 = seed[s + 1];
    if (eli > 0)
      for (long long a = b0; a < b1; a++) {
        const int2 en = list[a];
        const int elj = P.el2db[ATX_ENTRY_EL(en.y)];
        if (elj <= 0) continue;
        const double4 pj = pos4[en.x];
        double rx, ry, rz;
        bop_list_vec(A, pi, pj, en.y, rx, ry, rz);
        double rlij = rx * rx + ry * ry + rz * rz;   // squared until the bond is accepted
        const int ij = bop_pidx(P, eli, elj);
        double fcar = 1.0, dfcar = 0.0, fcbo = 1.0, dfcbo = 0.0;
        bool bond = false;
        const int ineb = ns;

        if (rlij < S.cut_in_l[ij] * S.cut_in_l[ij]) {
          bond = true;
          rlij = sqrt(rlij);
        } else if (rlij < S.max_cut_sq[ij] && S.cut_out_l[ij] < S.cut_out_h[ij]) {
          bool screened = false, need_derivative = false;
          double sij = 0.0, dsijdrij = 0.0;
          for (long long kn = b0; kn < b1 && !(screened || sij < S.screening_threshold); kn++) {
            if (kn == a) continue;   // k /= j .or. dc(kn) /= dc(jn): a different list slot
            const int2 ek = list[kn];
            const double4 pk = pos4[ek.x];
            double kx, ky, kz;
            bop_list_vec(A, pi, pk, ek.y, kx, ky, kz);
            const double rlik = kx * kx + ky * ky + kz * kz;
            if (!(rlik < S.C_dr_cut[ij] * rlij)) continue;
            const double dot_ij_ik = rx * kx + ry * ky + rz * kz;
            const double jx = -rx + kx, jy = -ry + ky, jz = -rz + kz;
            const double dot_ij_jk = rx * jx + ry * jy + rz * jz;
            const double rljk = jx * jx + jy * jy + jz * jz;
            if (dot_ij_ik > S.dot_threshold && dot_ij_jk < -S.dot_threshold) {
              const double xik = rlik / rlij, xjk = rljk / rlij;
              const double xm = xik - xjk, xp = xik + xjk;
              double fac = 1.0 / (1 - xm * xm);
              const double C = (2 * xp - xm * xm - 1) * fac;
              if (C <= S.Cmin[ij]) {
                screened = true;
              } else if (C < S.Cmax[ij]) {
                need_derivative = true;
                const double Cmax_C = S.Cmax[ij] - C, C_Cmin = C - S.Cmin[ij];
                const double q = Cmax_C / C_Cmin;
                sij = sij - q * q;
                const double dCdrik = 4 * xik * fac * (1 + (C - 1) * xm);
                const double dCdrjk = 4 * xjk * fac * (1 - (C - 1) * xm);
                const double dCdrij = -(dCdrik + dCdrjk);
                fac = 2 * Cmax_C * S.dC[ij] / (C_Cmin * C_Cmin * C_Cmin);
                dsijdrij = dsijdrij + fac * dCdrij;
                if (!COUNT && ns < T.NS) {
                  TB(kslot, ns) = (int)(kn - b0);
                  TB(arik, ns) = fac * dCdrik / rlik;
                  TB(arjk, ns) = fac * dCdrjk / rljk;
                }
                ns++;
              }
            }
          }
          if ((screened || sij < S.screening_threshold) && rlij > S.cut_in_h2[ij]) {
            ns = ineb;   // fully screened: no bond, screening neighbours discarded
          } else {
            bond = true;
            rlij = sqrt(rlij);
            double fcin, dfcin, fa, dfa, fb, dfb;
            if (screened) {
              bop_scrcut(S.trig, S.cin[ij], rlij, fcin, dfcin);
              fcar = fcin; dfcar = dfcin; fcbo = fcin; dfcbo = dfcin;
              ns = ineb;
            } else if (need_derivative) {
              sij = exp(sij);
              bop_scrcut(S.trig, S.cin[ij], rlij, fcin, dfcin);
              bop_scrcut(S.trig, S.cout[ij], rlij, fa, dfa);
              bop_scrcut(S.trig, S.cbo[ij], rlij, fb, dfb);
              fcar = (1.0 - fcin) * sij * fa + fcin;
              dfcar = (1.0 - fcin) * sij * (dfa + fa * dsijdrij / rlij) - dfcin * sij * fa + dfcin;
              fcbo = (1.0 - fcin) * sij * fb + fcin;
              dfcbo = (1.0 - fcin) * sij * (dfb + fb * dsijdrij / rlij) - dfcin * sij * fb + dfcin;
              if (!COUNT) {
                const int hi = ns < T.NS ? ns : T.NS;
                for (int q = ineb; q < hi; q++) {
                  const double ar = TB(arik, q), aj = TB(arjk, q);
                  TB(boik, q) = ar * sij * fb * (1.0 - fcin);
                  TB(bojk, q) = aj * sij * fb * (1.0 - fcin);
                  TB(arik, q) = ar * sij * fa * (1.0 - fcin);
                  TB(arjk, q) = aj * sij * fa * (1.0 - fcin);
                  TB(sfac, q) = 0.0;
                }
              }
            } else {
              bop_scrcut(S.trig, S.cout[ij], rlij, fa, dfa);
              bop_scrcut(S.trig, S.cbo[ij], rlij, fb, dfb);
              if (rlij < S.cut_in_h[ij]) {
                bop_scrcut(S.trig, S.cin[ij], rlij, fcin, dfcin);
                fcar = (1.0 - fcin) * fa + fcin;
                dfcar = (1.0 - fcin) * dfa - dfcin * fa + dfcin;
                fcbo = (1.0 - fcin) * fb + fcin;
                dfcbo = (1.0 - fcin) * dfb - dfcin * fb + dfcin;
              } else {
                fcar = fa; dfcar = dfa; fcbo = fb; dfcbo = dfb;
              }
              ns = ineb;   // need_derivative false: nothing was stored
            }
          }
        } else if (rlij < S.cut_in_h2[ij]) {
          // pair without an outer cutoff: plain inner cutoff
          bond = true;
          rlij = sqrt(rlij);
          bop_scrcut(S.trig, S.cin[ij], rlij, fcar, dfcar);
          fcbo = fcar; dfcbo = dfcar;
        }

        if (bond) {
          if (!COUNT && nb < T.NB) {
            const double ri = 1.0 / rlij;
            TB(rnx, nb) = rx * ri; TB(rny, nb) = ry * ri; TB(rnz, nb) = rz * ri;
            TB(rl, nb) = rlij;
            TB(fcar, nb) = fcar; TB(dfcar, nb) = dfcar; TB(fcbo, nb) = fcbo; TB(dfcbo, nb) = dfcbo;
            TB(jidx, nb) = en.x; TB(slot, nb) = (int)(a - b0); TB(typ, nb) = ij;
            TB(sseed, nb) = ineb; TB(scnt, nb) = ns - ineb;
          }
          nb++;
        }
      }
    if (!COUNT) {
      T.nbond[s] = nb < T.NB ? nb : T.NB;
      if (nb > T.NB || ns > T.NS) atomicOr(flag, 1);
    }
  }
  if (COUNT) {
    nb = __reduce_max_sync(0xffffffffu, nb);
    ns = __reduce_max_sync(0xffffffffu, ns);
    if ((threadIdx.x & 31) == 0) {
      if (nb > 0) atomicMax(&flag[2], nb);
      if (ns > 0) atomicMax(&flag[3], ns);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(BOP_BLOCK)
k_bopscr_center(int nat, Mat3 A, BopDev P, const BopScrDev *__restrict__ Sp, BopScrTab T,
                const double4 *__restrict__ pos4, const long long *__restrict__ seed,
                const int2 *__restrict__ list, const int *__restrict__ mask, double4 *__restrict__ G,
                double *__restrict__ f, double *__restrict__ pe_own, double *__restrict__ wpa,
                double *__restrict__ epb, double *__restrict__ fpb, double *__restrict__ wpb,
                double *__restrict__ partials, const unsigned char *__restrict__ role,
                const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (BOP_BLOCK / 32)];
  const int s = blockIdx.x * BOP_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;

  if (s < nat && (!role || role[s] >= 1)) {
    const BopScrDev &S = *Sp;
    const double4 pi = pos4[s];
    const int eli = P.el2db[(int)pi.w];
    const long long b0 = seed[s];
    const int nb = eli > 0 ? T.nbond[s] : 0;
    const bool own = !role || role[s] >= 2;
    double fix = 0.0, fiy = 0.0, fiz = 0.0, pei = 0.0;
    const int mi = mask ? mask[s] : 1;

    // ---- loop 2 ----
    for (int ij = 0; ij < nb; ij++) {
      const int tij = TB(typ, ij);
      const int j = TB(jidx, ij);
      int maskfac = 2;
      if (mask) {
        int mj = mask[j];
        if (mi == 0 && mj == 0) maskfac = 0;
        else if (mi == 0 || mj == 0) maskfac = 1;
      }
      const double rlij = TB(rl, ij);
      if (!(maskfac > 0 && rlij < S.cut_out_h[tij])) continue;   // cut_ar_h = cut_out_h
      const double rlijr = 1.0 / rlij;
      const double nx = TB(rnx, ij), ny = TB(rny, ij), nz = TB(rnz, ij);
      const double rijx = rlij * nx, rijy = rlij * ny, rijz = rlij * nz;
      const double fcarij = TB(fcar, ij), dfcarijr = TB(dfcar, ij);
      double VAij, dVAij, VRij, dVRij;
      bop_VA<KIND>(P, tij, rlij, VAij, dVAij);
      bop_VR<KIND>(P, tij, rlij, VRij, dVRij);
      const double mf = 0.5 * maskfac;
      VAij *= mf; dVAij *= mf; VRij *= mf; dVRij *= mf;

      double zij = 0.0;
      double dix = 0, diy = 0, diz = 0, djx = 0, djy = 0, djz = 0;
      double wb[9], w[9];
#pragma unroll
      for (int q = 0; q < 9; q++) { wb[q] = 0.0; w[q] = 0.0; }

      for (int ik = 0; ik < nb; ik++) {
        if (ik == ij) continue;
        const int tik = TB(typ, ik);
        const double rlik = TB(rl, ik);
        if (!(rlik < S.cut_bo_h[tik])) {
          TB(kx, ik) = 0.0; TB(ky, ik) = 0.0; TB(kz, ik) = 0.0; TB(zf, ik) = 0.0;
          continue;
        }
        const double kx = TB(rnx, ik), ky = TB(rny, ik), kz = TB(rnz, ik);
        const double fcik = TB(fcbo, ik), dfcikr = TB(dfcbo, ik);
        double h_Dr, dh_dDr, g_costh, dg_dcosth;
        bop_h<KIND>(P, eli - 1, tij, tik, rlij - rlik, h_Dr, dh_dDr);
        const double costh = kx * nx + ky * ny + kz * nz;
        bop_g<KIND>(P, eli - 1, tik, costh, g_costh, dg_dcosth);
        double ex = kx * rlik - nx * rlij, ey = ky * rlik - ny * rlij, ez = kz * rlik - nz * rlij;
        const double disjk = sqrt(ex * ex + ey * ey + ez * ez);
        const double idis = 1.0 / disjk, rlikr = 1.0 / rlik;
        ex *= idis; ey *= idis; ez *= idis;
        const double dcsdij = rlikr - costh * rlijr;
        const double dcsdik = rlijr - costh * rlikr;
        const double dcsdjk = -disjk * rlijr * rlikr;
        const double dzfac = fcik * dg_dcosth * h_Dr;
        TB(zf, ik) = g_costh * h_Dr;
        zij += fcik * g_costh * h_Dr;
        const double dzdrij = g_costh * fcik * dh_dDr;
        const double dzdrik = g_costh * (dfcikr * h_Dr - fcik * dh_dDr);
        double dfx, dfy, dfz, dkx, dky, dkz;
        {
          double ci = -dcsdij * nx - dcsdik * kx, cj = dcsdij * nx - dcsdjk * ex, ck = dcsdik * kx + dcsdjk * ex;
          dix += -dzdrij * nx - dzdrik * kx + dzfac * ci;
          dfx = dzdrij * nx + dzfac * cj;
          dkx = dzdrik * kx + dzfac * ck;
        }
        {
          double ci = -dcsdij * ny - dcsdik * ky, cj = dcsdij * ny - dcsdjk * ey, ck = dcsdik * ky + dcsdjk * ey;
          diy += -dzdrij * ny - dzdrik * ky + dzfac * ci;
          dfy = dzdrij * ny + dzfac * cj;
          dky = dzdrik * ky + dzfac * ck;
        }
        {
          double ci = -dcsdij * nz - dcsdik * kz, cj = dcsdij * nz - dcsdjk * ez, ck = dcsdik * kz + dcsdjk * ez;
          diz += -dzdrij * nz - dzdrik * kz + dzfac * ci;
          dfz = dzdrij * nz + dzfac * cj;
          dkz = dzdrik * kz + dzfac * ck;
        }
        djx += dfx; djy += dfy; djz += dfz;
        TB(kx, ik) = dkx; TB(ky, ik) = dky; TB(kz, ik) = dkz;
        const double rikx = rlik * kx, riky = rlik * ky, rikz = rlik * kz;
        wb[0] -= rijx * dfx + rikx * dkx; wb[1] -= rijy * dfx + riky * dkx; wb[2] -= rijz * dfx + rikz * dkx;
        wb[3] -= rijx * dfy + rikx * dky; wb[4] -= rijy * dfy + riky * dky; wb[5] -= rijz * dfy + rikz * dky;
        wb[6] -= rijx * dfz + rikx * dkz; wb[7] -= rijy * dfz + riky * dkz; wb[8] -= rijz * dfz + rikz * dkz;
      }

      double bij, dfb;
      bop_bo<KIND>(P, eli - 1, tij, zij, fcarij, VAij, bij, dfb);
      const double e_bond = 0.5 * fcarij * (VRij + bij * VAij);
      pei += e_bond;
      const double dffac = 0.5 * (dVRij * fcarij + bij * dVAij * fcarij + VRij * dfcarijr + bij * VAij * dfcarijr);
      const double dfx = dffac * nx, dfy = dffac * ny, dfz = dffac * nz;
      fix += dfx - dfb * dix; fiy += dfy - dfb * diy; fiz += dfz - dfb * diz;
      double fjx = -dfx - dfb * djx, fjy = -dfy - dfb * djy, fjz = -dfz - dfb * djz;
      w[0] = rijx * dfx - dfb * wb[0]; w[1] = rijy * dfx - dfb * wb[1]; w[2] = rijz * dfx - dfb * wb[2];
      w[3] = rijx * dfy - dfb * wb[3]; w[4] = rijy * dfy - dfb * wb[4]; w[5] = rijz * dfy - dfb * wb[5];
      w[6] = rijx * dfz - dfb * wb[6]; w[7] = rijy * dfz - dfb * wb[7]; w[8] = rijz * dfz - dfb * wb[8];

      // forces on the other bond partners of i and the screening weight of the bonds i-k
      for (int ik = 0; ik < nb; ik++) {
        if (ik == ij) continue;
        const long long ak = b0 + TB(slot, ik);
        double4 g = G[ak];
        g.x -= dfb * TB(kx, ik); g.y -= dfb * TB(ky, ik); g.z -= dfb * TB(kz, ik);
        G[ak] = g;
        const int q0 = TB(sseed, ik), q1 = q0 + TB(scnt, ik);
        const double add = TB(zf, ik) * dfb;
        for (int q = q0; q < q1; q++) TB(sfac, q) += add;
      }

      // screening neighbours of bond i-j, attractive / repulsive part (bop_kernel.f90:1478-1512)
      {
        const double dff = 0.5 * (VRij + bij * VAij);
        const int q0 = TB(sseed, ij), q1 = q0 + TB(scnt, ij);
        for (int q = q0; q < q1; q++) {
          const long long ak = b0 + TB(kslot, q);
          const int2 ek = list[ak];
          const double4 pk = pos4[ek.x];
          double kx, ky, kz;
          bop_list_vec(A, pi, pk, ek.y, kx, ky, kz);
          const double jx = -rijx + kx, jy = -rijy + ky, jz = -rijz + kz;
          const double c1 = dff * TB(arik, q), c2 = dff * TB(arjk, q);
          const double d1x = c1 * kx, d1y = c1 * ky, d1z = c1 * kz;
          const double d2x = c2 * jx, d2y = c2 * jy, d2z = c2 * jz;
          fix += d1x; fiy += d1y; fiz += d1z;
          fjx += d2x; fjy += d2y; fjz += d2z;
          double4 g = G[ak];
          g.x -= d1x + d2x; g.y -= d1y + d2y; g.z -= d1z + d2z;
          G[ak] = g;
          w[0] += kx * d1x + jx * d2x; w[1] += ky * d1x + jy * d2x; w[2] += kz * d1x + jz * d2x;
          w[3] += kx * d1y + jx * d2y; w[4] += ky * d1y + jy * d2y; w[5] += kz * d1y + jz * d2y;
          w[6] += kx * d1z + jx * d2z; w[7] += ky * d1z + jy * d2z; w[8] += kz * d1z + jz * d2z;
        }
      }

      const long long aj = b0 + TB(slot, ij);
      {
        double4 g = G[aj];
        g.x += fjx; g.y += fjy; g.z += fjz; g.w += e_bond;
        G[aj] = g;
      }
      if (own) {
#pragma unroll
        for (int q = 0; q < 9; q++) acc[1 + q] += w[q];
        acc[0] += e_bond;
      }
      if (epb) epb[aj] += e_bond;
      if (fpb) { fpb[3 * aj] += dfx; fpb[3 * aj + 1] += dfy; fpb[3 * aj + 2] += dfz; }
      if (wpb) {
#pragma unroll
        for (int q = 0; q < 9; q++) wpb[9 * aj + q] += w[q];
      }
      if (wpa) {
#pragma unroll
        for (int q = 0; q < 9; q++) {
          atomicAdd(&wpa[9 * (size_t)s + q], 0.5 * w[q]);
          atomicAdd(&wpa[9 * (size_t)j + q], 0.5 * w[q]);
        }
      }
    }

    // ---- loop 3: screening forces of the bond-order cutoff (bop_kernel.f90:1531-1611) ----
    for (int ij = 0; ij < nb; ij++) {
      const int q0 = TB(sseed, ij), q1 = q0 + TB(scnt, ij);
      if (q0 >= q1) continue;
      const int j = TB(jidx, ij);
      const double rlij = TB(rl, ij);
      const double rijx = rlij * TB(rnx, ij), rijy = rlij * TB(rny, ij), rijz = rlij * TB(rnz, ij);
      double fjx = 0, fjy = 0, fjz = 0;
      double w[9];
#pragma unroll
      for (int q = 0; q < 9; q++) w[q] = 0.0;
      for (int q = q0; q < q1; q++) {
        const double sf = TB(sfac, q);
        const double c1 = sf * TB(boik, q), c2 = sf * TB(bojk, q);
        const long long ak = b0 + TB(kslot, q);
        const int2 ek = list[ak];
        const double4 pk = pos4[ek.x];
        double kx, ky, kz;
        bop_list_vec(A, pi, pk, ek.y, kx, ky, kz);
        const double jx = -rijx + kx, jy = -rijy + ky, jz = -rijz + kz;
        const double d1x = c1 * kx, d1y = c1 * ky, d1z = c1 * kz;
        const double d2x = c2 * jx, d2y = c2 * jy, d2z = c2 * jz;
        fix += d1x; fiy += d1y; fiz += d1z;
        fjx += d2x; fjy += d2y; fjz += d2z;
        double4 g = G[ak];
        g.x -= d1x + d2x; g.y -= d1y + d2y; g.z -= d1z + d2z;
        G[ak] = g;
        w[0] += kx * d1x + jx * d2x; w[1] += ky * d1x + jy * d2x; w[2] += kz * d1x + jz * d2x;
        w[3] += kx * d1y + jx * d2y; w[4] += ky * d1y + jy * d2y; w[5] += kz * d1y + jz * d2y;
        w[6] += kx * d1z + jx * d2z; w[7] += ky * d1z + jy * d2z; w[8] += kz * d1z + jz * d2z;
      }
      const long long aj = b0 + TB(slot, ij);
      {
        double4 g = G[aj];
        g.x += fjx; g.y += fjy; g.z += fjz;
        G[aj] = g;
      }
      if (own) {
#pragma unroll
        for (int q = 0; q < 9; q++) acc[1 + q] += w[q];
      }
      if (wpb) {
#pragma unroll
        for (int q = 0; q < 9; q++) wpb[9 * aj + q] += w[q];
      }
      if (wpa) {
#pragma unroll
        for (int q = 0; q < 9; q++) {
          atomicAdd(&wpa[9 * (size_t)s + q], 0.5 * w[q]);
          atomicAdd(&wpa[9 * (size_t)j + q], 0.5 * w[q]);
        }
      }
    }
    f[3 * s] = fix; f[3 * s + 1] = fiy; f[3 * s + 2] = fiz;
    pe_own[s] = pei;
  }
  atx_block_sum<ATX_NSUM, BOP_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}
#undef TB

__global__ void k_bop_gather(int nat, const long long *__restrict__ seed, const int *__restrict__ rev,
                             const double4 *__restrict__ G, const double *__restrict__ pe_own,
                             double *__restrict__ f, double *__restrict__ epa,
                             const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  if (role && role[s] < 2) {
    // ghosts: their owner computes (and, under LAMMPS, reverse-communicates nothing from here)
    f[3 * s] = 0.0; f[3 * s + 1] = 0.0; f[3 * s + 2] = 0.0;
    if (epa) epa[s] = 0.0;
    return;
  }
  double fx = f[3 * s], fy = f[3 * s + 1], fz = f[3 * s + 2], pe = pe_own[s];
  for (long long a = seed[s]; a < seed[s + 1]; a++) {
    int b = rev[a];
    if (b >= 0) {
      double4 g = G[b];
      fx += g.x; fy += g.y; fz += g.z; pe += g.w;
    }
  }
  f[3 * s] = fx; f[3 * s + 1] = fy; f[3 * s + 2] = fz;
  if (epa) epa[s] = 0.5 * pe;
}

// ---------------------------------------------------------------------------

extern "C" int atx_bop_create(atx_ctx *ctx, const atx_bop_params *par, atx_bop **out) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (!ctx || !par || !out) return ATX_ERROR_UNSPECIFIED;
  if (par->kind < 1 || par->kind > 3 || par->nel < 1 || par->nel > ATX_BOP_MAX_EL) {
    atx_set_error("atx_bop_create: invalid kind or number of elements.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_bop *pot = new atx_bop();
  pot->ctx = ctx;
  pot->par = *par;
  BopDev &D = pot->dev;
  D.kind = par->kind;
  D.nel = par->nel;
  int npairs = par->nel * (par->nel + 1) / 2;
  for (int i = 0; i < npairs; i++) {
    D.r1[i] = par->r1[i]; D.r2[i] = par->r2[i];
    D.r1sq[i] = par->r1[i] * par->r1[i]; D.r2sq[i] = par->r2[i] * par->r2[i];
    D.cfac[i] = BOP_PI / (par->r2[i] - par->r1[i]);
    D.A[i] = par->A[i]; D.B[i] = par->B[i]; D.xi[i] = par->xi[i]; D.lambda[i] = par->lambda[i];
    D.mu[i] = par->mu[i]; D.omega[i] = par->omega[i]; D.mubo[i] = par->mubo[i]; D.m[i] = par->m[i];
    if (par->kind == ATX_BOP_BRENNER) {
      // brenner_module.f90:269-288
      if (par->pd[i] * par->pd[i] == 0.0) {
        atx_set_error("d = 0! This leads to problems computing c**2/d**2. Please specify d != 0.");
        delete pot;
        return ATX_ERROR_UNSPECIFIED;
      }
      if (par->S[i] <= 1.0) {
        atx_set_error("S <= 1! This leads to problems computing (S-1)**(-1). Please specify S > 1.");
        delete pot;
        return ATX_ERROR_UNSPECIFIED;
      }
      D.bo_exp[i] = -0.5 / par->pn[i];
      D.bo_fac[i] = 0.5 * D.bo_exp[i] * par->pn[i];
      D.bo_exp1[i] = D.bo_exp[i] - 1.0;
      D.expR[i] = par->pbeta[i] * sqrt(2 * par->S[i]);
      D.expA[i] = par->pbeta[i] * sqrt(2 / par->S[i]);
      D.c_sq[i] = par->pc[i] * par->pc[i];
      D.d_sq[i] = par->pd[i] * par->pd[i];
      D.c_d[i] = D.c_sq[i] / D.d_sq[i];
      D.VR_f[i] = par->D0[i] / (par->S[i] - 1);
      D.VA_f[i] = par->S[i] * par->D0[i] / (par->S[i] - 1);
      D.r0[i] = par->r0[i]; D.gamma[i] = par->gamma[i]; D.ph[i] = par->ph[i]; D.pn[i] = par->pn[i];
    }
  }
  for (int i = 0; i < par->nel; i++) {
    D.beta[i] = par->beta[i]; D.n[i] = par->n[i]; D.c[i] = par->c[i]; D.d[i] = par->d[i];
    D.h[i] = par->h[i]; D.eta[i] = par->eta[i]; D.delta[i] = par->delta[i];
    D.c1[i] = par->c1[i]; D.c2[i] = par->c2[i]; D.c3[i] = par->c3[i]; D.c4[i] = par->c4[i];
    D.c5[i] = par->c5[i];
    D.tb[i] = par->kind == ATX_BOP_TERSOFF ? pow(par->beta[i], par->n[i]) : 0.0;
    D.te[i] = par->kind == ATX_BOP_TERSOFF ? -0.5 / par->n[i] : 0.0;
    D.c_sq_e[i] = par->c[i] * par->c[i];
    D.d_sq_e[i] = par->d[i] * par->d[i];
    D.one_p_c2d2[i] = par->kind == ATX_BOP_TERSOFF ? 1.0 + D.c_sq_e[i] / D.d_sq_e[i] : 0.0;
  }
  for (int k = 0; k < 32; k++) D.el2db[k] = -1;
  ATX_PASS(pot->flag.reserve(64));
  ATX_CUDA(cudaMemset(pot->flag.ptr, 0, 64 * sizeof(int)));
  *out = pot;
  return 0;
}

extern "C" int atx_bop_create_juslin(atx_ctx *ctx, const atx_juslin_params *par, atx_bop **out) {
  if (ctx) cudaSetDevice(ctx->device);
  if (!ctx || !par || !out) return ATX_ERROR_UNSPECIFIED;
  if (par->nel < 1 || par->nel > ATX_BOP_MAX_EL) {
    atx_set_error("atx_bop_create_juslin: invalid number of elements.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_bop *pot = new atx_bop();
  pot->ctx = ctx;
  pot->par.kind = ATX_BOP_JUSLIN;
  pot->par.nel = par->nel;
  for (int i = 0; i < par->nel; i++) pot->par.Z[i] = par->Z[i];
  BopDev &D = pot->dev;
  D.kind = ATX_BOP_JUSLIN;
  D.nel = par->nel;
  const int npairs = par->nel * par->nel;
  for (int i = 0; i < npairs; i++) {
    // juslin_module.f90:322-357
    D.r1[i] = par->r1[i]; D.r2[i] = par->r2[i];
    D.r1sq[i] = par->r1[i] * par->r1[i]; D.r2sq[i] = par->r2[i] * par->r2[i];
    D.cfac[i] = BOP_PI / (par->r2[i] - par->r1[i]);
    if (par->d[i] * par->d[i] == 0.0) {
      atx_set_error("d = 0! This leads to problems computing c**2/d**2. Please specify d != 0.");
      delete pot;
      return ATX_ERROR_UNSPECIFIED;
    }
    if (par->S[i] <= 1.0) {
      atx_set_error("S <= 1! This leads to problems computing (S-1)**(-1). Please specify S > 1.");
      delete pot;
      return ATX_ERROR_UNSPECIFIED;
    }
    D.bo_exp[i] = -0.5 / par->n[i];
    D.bo_fac[i] = 0.5 * D.bo_exp[i] * par->n[i];
    D.bo_exp1[i] = D.bo_exp[i] - 1.0;
    D.expR[i] = par->beta[i] * sqrt(2 * par->S[i]);
    D.expA[i] = par->beta[i] * sqrt(2 / par->S[i]);
    D.c_sq[i] = par->c[i] * par->c[i];
    D.d_sq[i] = par->d[i] * par->d[i];
    D.c_d[i] = D.c_sq[i] / D.d_sq[i];
    D.VR_f[i] = par->D0[i] / (par->S[i] - 1);
    D.VA_f[i] = par->S[i] * par->D0[i] / (par->S[i] - 1);
    D.r0[i] = par->r0[i]; D.gamma[i] = par->gamma[i]; D.ph[i] = par->h[i]; D.pn[i] = par->n[i];
  }
  const int ntrip = par->nel * par->nel * par->nel;
  for (int i = 0; i < ntrip; i++) {
    D.talpha[i] = par->alpha[i]; D.tomega[i] = par->omega[i]; D.tm[i] = par->m[i];
  }
  for (int k = 0; k < 32; k++) D.el2db[k] = -1;
  ATX_PASS(pot->flag.reserve(64));
  ATX_CUDA(cudaMemset(pot->flag.ptr, 0, 64 * sizeof(int)));
  *out = pot;
  return 0;
}

static void bop_expcut_init(ExpCut &t, double r1, double r2) {
  // exp_cutoff_init, src/support/cutoff.f90:232-255
  t.r1 = r1;
  t.r2 = r2;
  t.fac1 = 1.0 / (r2 - r1);
  const double val1 = exp(-8.0);
  const double dval1 = -24 * val1;
  const double ddval1 = -48 * val1 - 24 * dval1;
  t.c = (-3 * dval1 + ddval1) / 3;
  t.d = (2 * dval1 - ddval1) / 4;
  t.fac2 = 1.0 / (1 - val1 - t.c - t.d);
  t.off = val1 + t.c + t.d;
}

extern "C" int atx_bop_create_screened(atx_ctx *ctx, const atx_bop_params *par,
                                       const atx_bop_screening *scr, atx_bop **out) {
  if (ctx) cudaSetDevice(ctx->device);
  if (!scr) return ATX_ERROR_UNSPECIFIED;
  ATX_PASS(atx_bop_create(ctx, par, out));
  atx_bop *pot = *out;
  pot->screened = true;
  pot->scr_par = atx_juslin_screening{};
  for (int i = 0; i < ATX_BOP_MAX_PAIRS; i++) {
    pot->scr_par.or1[i] = scr->or1[i]; pot->scr_par.or2[i] = scr->or2[i];
    pot->scr_par.bor1[i] = scr->bor1[i]; pot->scr_par.bor2[i] = scr->bor2[i];
    pot->scr_par.Cmin[i] = scr->Cmin[i]; pot->scr_par.Cmax[i] = scr->Cmax[i];
  }
  // default_bind_to_func.f90:44-104
  BopScrDev S{};
  const int npairs = par->nel * (par->nel + 1) / 2;
  for (int i = 0; i < npairs; i++) {
    S.Cmin[i] = scr->Cmin[i];
    S.Cmax[i] = scr->Cmax[i];
    S.dC[i] = S.Cmax[i] - S.Cmin[i];
    S.C_dr_cut[i] = S.Cmax[i] > 2.0 ? S.Cmax[i] * S.Cmax[i] / (4 * (S.Cmax[i] - 1)) : 1.0;
    bop_expcut_init(S.cin[i], par->r1[i], par->r2[i]);
    S.cut_in_l[i] = par->r1[i];
    S.cut_in_h[i] = par->r2[i];
    S.cut_in_h2[i] = par->r2[i] * par->r2[i];
    bop_expcut_init(S.cout[i], scr->or1[i], scr->or2[i]);
    S.cut_out_l[i] = scr->or1[i];
    S.cut_out_h[i] = scr->or2[i];
    bop_expcut_init(S.cbo[i], scr->bor1[i], scr->bor2[i]);
    S.cut_bo_h[i] = scr->bor2[i];
    double m = S.cut_in_h[i];
    if (S.cut_out_h[i] > m) m = S.cut_out_h[i];
    if (S.cut_bo_h[i] > m) m = S.cut_bo_h[i];
    S.max_cut_sq[i] = m * m;
  }
  S.screening_threshold = log(1e-6);   // tersoff_type.f90:86
  S.dot_threshold = 1e-10;
  ATX_PASS(pot->scr_dev.reserve(1));
  ATX_CUDA(cudaMemcpy(pot->scr_dev.ptr, &S, sizeof(S), cudaMemcpyHostToDevice));
  return 0;
}

// JuslinScr: juslin_module.f90 compiled with SCREENING (juslin_scr.f90): nel**2 pair rows, cosine
// switches for all three cutoffs, C_dr_cut = Cmax**2 / (4 (Cmax - 1)) for every pair (:301-307)
extern "C" int atx_bop_create_juslin_screened(atx_ctx *ctx, const atx_juslin_params *par,
                                              const atx_juslin_screening *scr, atx_bop **out) {
  if (ctx) cudaSetDevice(ctx->device);
  if (!scr) return ATX_ERROR_UNSPECIFIED;
  ATX_PASS(atx_bop_create_juslin(ctx, par, out));
  atx_bop *pot = *out;
  pot->screened = true;
  pot->scr_par = *scr;
  BopScrDev S{};
  S.trig = 1;
  const int npairs = par->nel * par->nel;
  for (int i = 0; i < npairs; i++) {
    S.Cmin[i] = scr->Cmin[i];
    S.Cmax[i] = scr->Cmax[i];
    S.dC[i] = S.Cmax[i] - S.Cmin[i];
    S.C_dr_cut[i] = S.Cmax[i] * S.Cmax[i] / (4 * (S.Cmax[i] - 1));
    S.cin[i].r1 = par->r1[i]; S.cin[i].r2 = par->r2[i];       // only r1 / r2 enter the cosine switch
    S.cout[i].r1 = scr->or1[i]; S.cout[i].r2 = scr->or2[i];
    S.cbo[i].r1 = scr->bor1[i]; S.cbo[i].r2 = scr->bor2[i];
    S.cut_in_l[i] = par->r1[i];
    S.cut_in_h[i] = par->r2[i];
    S.cut_in_h2[i] = par->r2[i] * par->r2[i];
    S.cut_out_l[i] = scr->or1[i];
    S.cut_out_h[i] = scr->or2[i];
    S.cut_bo_h[i] = scr->bor2[i];
    double m = S.cut_in_h[i];
    if (S.cut_out_h[i] > m) m = S.cut_out_h[i];
    if (S.cut_bo_h[i] > m) m = S.cut_bo_h[i];
    S.max_cut_sq[i] = m * m;
  }
  S.screening_threshold = log(1e-6);
  S.dot_threshold = 1e-10;
  ATX_PASS(pot->scr_dev.reserve(1));
  ATX_CUDA(cudaMemcpy(pot->scr_dev.ptr, &S, sizeof(S), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int atx_bop_destroy(atx_bop *pot) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  delete pot;
  return 0;
}

extern "C" int atx_bop_bind_to(atx_bop *pot, atx_particles *p, atx_neighbors *nl, int nel,
                               const int *el2Z) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (nel > 31) {
    atx_set_error("Too many particle element ids.");
    return ATX_ERROR_UNSPECIFIED;
  }
  BopDev &D = pot->dev;
  for (int k = 0; k < 32; k++) D.el2db[k] = -1;
  for (int k = 0; k < nel; k++)
    for (int e = 0; e < pot->par.nel; e++)
      if (el2Z[k] > 0 && el2Z[k] == pot->par.Z[e]) D.el2db[k + 1] = e + 1;
  // default_bind_to_func.f90:118-141: request r2 of every present pair
  if (nl)
    for (int i = 1; i <= nel; i++)
      for (int j = 1; j <= nel; j++)
        if (D.el2db[i] > 0 && D.el2db[j] > 0) {
          int a = D.el2db[i], b = D.el2db[j];
          int x = (a - 1) + (b - 1) * D.nel, y = (b - 1) + (a - 1) * D.nel;
          int c = (a - 1) * a / 2, d = (b - 1) * b / 2;
          int ij = (x < y ? x : y) - (c < d ? c : d);
          if (D.kind == ATX_BOP_JUSLIN) ij = (b - 1) + (a - 1) * D.nel;   // PAIR_INDEX_NS
          double cutoff = D.r2[ij];
          if (pot->screened) {
            // default_bind_to_func.f90:106-130: sqrt(C_dr_cut(pair)) * the largest cutoff of ANY pair
            const atx_juslin_screening &sp = pot->scr_par;
            const bool jus = D.kind == ATX_BOP_JUSLIN;
            const int npairs = jus ? D.nel * D.nel : D.nel * (D.nel + 1) / 2;
            double mx = 0.0;
            for (int q = 0; q < npairs; q++) {
              mx = std::max(mx, D.r2[q]);
              mx = std::max(mx, sp.or2[q]);
              mx = std::max(mx, sp.bor2[q]);
            }
            const double cm = sp.Cmax[ij];
            // juslin_module.f90:301-307, 379-395: the unconditional form
            if (jus) cutoff = cm > 1.0 ? std::sqrt(cm * cm / (4 * (cm - 1))) * mx : 0.0;
            else cutoff = std::sqrt(cm > 2.0 ? cm * cm / (4 * (cm - 1)) : 1.0) * mx;
          }
          ATX_PASS(atx_neighbors_request_interaction_range_pair(nl, cutoff, i, j));
        }
  pot->bound = true;
  return 0;
}

template <int KIND, int NB, int MINB, bool VIRIAL, bool LEAN = false>
static int launch_center_v(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                         const PotOut &o, double *pe_own, double *epb, double *fpb, double *wpb,
                         int nblocks, int pstride, int b0 = 0) {
  if (nblocks <= 0) return 0;
  size_t smem = sizeof(BondSmem<NB, LEAN>);
  static int attr_dev = -1;   // the attribute belongs to the device: set it again when the device changes
  if (attr_dev != pot->ctx->device) {
    ATX_CUDA(cudaFuncSetAttribute(k_bop_center<KIND, NB, MINB, VIRIAL, LEAN>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (LEAN)   // eight blocks of 27.8 KB need (nearly) the whole shared-memory carve-out
      ATX_CUDA(cudaFuncSetAttribute(k_bop_center<KIND, NB, MINB, VIRIAL, LEAN>,
                                    cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_dev = pot->ctx->device;
  }
  k_bop_center<KIND, NB, MINB, VIRIAL, LEAN><<<nblocks, BOP_BLOCK, smem, pot->ctx->stream>>>(
      nl->nat, p->Abox, pot->dev, nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, mask, pot->G.ptr, o.f,
      pe_own, o.wpa, epb, fpb, wpb, pot->sc.partials.ptr, pstride, pot->queue.ptr, pot->flag.ptr + 1,
      o.role, o.stop, b0);
  ATX_LAUNCHED();
  return 0;
}

// NVE stepping does not need the virial (PotOut::want_virial = false): that instantiation drops the
// nine virial accumulators and their updates in the k-loop (fewer registers, ~25 % fewer FP64 ops)
template <int KIND, int NB, int MINB = 1>
static int launch_center(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                         const PotOut &o, double *pe_own, double *epb, double *fpb, double *wpb,
                         int nblocks, int pstride, int b0 = 0) {
  if (o.want_virial || o.wpa || wpb)
    return launch_center_v<KIND, NB, MINB, true>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nblocks, pstride, b0);
  if constexpr (NB == 4) {
    // an MD step whose energy nobody reads: the lean table, eight blocks per SM
    static const bool lean_env = !(getenv("ATX_BOP_LEAN") && atoi(getenv("ATX_BOP_LEAN")) == 0);
    if (lean_env && !mask && !epb && !fpb && !o.epa && !o.want_sums)
      return launch_center_v<KIND, 4, 8, false, true>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nblocks, pstride, b0);
  }
  return launch_center_v<KIND, NB, MINB, false>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nblocks, pstride, b0);
}

// bond-table depths that are compiled; the main pass uses the smallest one that holds (nearly)
// every atom, the queued pass BOP_NB_MAX
static const int kBopDepths[] = {4, 6, 8, 12, BOP_NB_MAX};

template <int KIND>
static int launch_center_range(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                               const PotOut &o, double *pe_own, double *epb, double *fpb, double *wpb,
                               int b0, int b1, int pstride) {
  const int nb = b1 - b0;
  switch (pot->nb_cap) {
    // depth 4: 30.9 KB of shared memory per block -> 7 blocks/SM if the kernel stays within 146 registers
    case 4: return launch_center<KIND, 4, 7>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nb, pstride, b0);
    case 6: return launch_center<KIND, 6>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nb, pstride, b0);
    case 8: return launch_center<KIND, 8>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nb, pstride, b0);
    case 12: return launch_center<KIND, 12>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nb, pstride, b0);
    default: return launch_center<KIND, BOP_NB_MAX>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, nb, pstride, b0);
  }
}

template <int KIND>
static int launch_center_nb(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask,
                            const PotOut &o, double *pe_own, double *epb, double *fpb, double *wpb,
                            int nblocks) {
  cudaStream_t st = pot->ctx->stream;
  const int nq = pot->ctx->sm_count;            // blocks of the queued pass
  const int pstride = nblocks + nq;
  // PotOut::phase: 1 = the centre blocks inside [split_lo, split_hi) only, 2 = everything else
  int bl = (o.split_lo + BOP_BLOCK - 1) / BOP_BLOCK, bh = o.split_hi / BOP_BLOCK;
  if (bh > nblocks) bh = nblocks;
  if (bh < bl || o.phase == 0) bl = bh = 0;
  if (o.phase != 2) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr + 1, 0, sizeof(int), st));   // queue length
  ProfScope ps_(pot->ctx, "bop_force");
  if (o.phase == 1)
    return launch_center_range<KIND>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, bl, bh, pstride);
  ATX_PASS(launch_center_range<KIND>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, 0, bl, pstride));
  ATX_PASS(launch_center_range<KIND>(pot, p, nl, mask, o, pe_own, epb, fpb, wpb, bh, nblocks, pstride));
  size_t smem = sizeof(BondSmem<BOP_NB_MAX>);
  static int attr_dev = -1;
  if (attr_dev != pot->ctx->device) {
    ATX_CUDA(cudaFuncSetAttribute(k_bop_center_queued<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr_dev = pot->ctx->device;
  }
  k_bop_center_queued<KIND><<<nq, BOP_BLOCK, smem, st>>>(
      p->Abox, pot->dev, nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, mask, pot->G.ptr, o.f, pe_own, o.wpa,
      epb, fpb, wpb, pot->sc.partials.ptr, pstride, nblocks, pot->queue.ptr, pot->flag.ptr + 1,
      pot->flag.ptr, o.role, o.stop);
  ATX_LAUNCHED();
  return 0;
}

// lays the BopScrTab pointers over the two backing buffers
static int bopscr_layout(atx_bop *pot, int nat, int NB, int NS) {
  BopScrTab &T = pot->tab;
  T.nat = nat; T.NB = NB; T.NS = NS;
  const size_t nb = (size_t)NB * nat, ns = (size_t)NS * nat;
  ATX_PASS(pot->scr_d.reserve(12 * nb + 5 * ns + 16));
  ATX_PASS(pot->scr_i.reserve(5 * nb + ns + (size_t)nat + 16));
  double *d = pot->scr_d.ptr;
  double **df[] = {&T.rnx, &T.rny, &T.rnz, &T.rl, &T.fcar, &T.dfcar, &T.fcbo, &T.dfcbo, &T.kx, &T.ky, &T.kz, &T.zf};
  for (double **q : df) { *q = d; d += nb; }
  double **sf[] = {&T.arik, &T.arjk, &T.boik, &T.bojk, &T.sfac};
  for (double **q : sf) { *q = d; d += ns; }
  int *i = pot->scr_i.ptr;
  int **bi[] = {&T.jidx, &T.slot, &T.typ, &T.sseed, &T.scnt};
  for (int **q : bi) { *q = i; i += nb; }
  T.kslot = i; i += ns;
  T.nbond = i;
  return 0;
}

static int bopscr_compute(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask_sorted,
                          const PotOut &o, double *pe_own, double *epb, double *fpb, double *wpb) {
  atx_ctx *ctx = pot->ctx;
  cudaStream_t st = ctx->stream;
  const int nat = nl->nat;
  const int nblocks = nat > 0 ? (nat + BOP_BLOCK - 1) / BOP_BLOCK : 1;
  ATX_PASS(pot->sc.partials.reserve((size_t)nblocks * ATX_NSUM));
  if (!o.stop) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr, 0, sizeof(int), st));
  if (!o.stop && (pot->nb_cap == 0 || pot->sized_nl != nl || pot->sized_build != nl->nbuilds)) {
    // table depths from a counting pass over the current configuration (+ margin for the motion
    // inside a Verlet shell; exceeding it during batched MD raises the overflow flag)
    ATX_CUDA(cudaMemsetAsync(pot->flag.ptr + 2, 0, 2 * sizeof(int), st));
    if (nat > 0) {
      k_bopscr_bonds<true><<<(nat + 127) / 128, 128, 0, st>>>(nat, p->Abox, pot->dev, pot->scr_dev.ptr, pot->tab,
                                                              nl->pos4.ptr, nl->seed.ptr, nl->list.ptr,
                                                              pot->flag.ptr, o.role, nullptr);
      ATX_LAUNCHED();
    }
    ATX_PASS(pot->hflag.reserve(64));
    ATX_CUDA(cudaMemcpyAsync(pot->hflag.ptr, pot->flag.ptr + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    const int NB = pot->hflag.ptr[0] + 2, NS = pot->hflag.ptr[1] + pot->hflag.ptr[1] / 4 + 8;
    ATX_PASS(bopscr_layout(pot, nat, NB, NS));
    pot->nb_cap = NB;
    pot->sized_nl = nl;
    pot->sized_build = nl->nbuilds;
  } else if (pot->tab.nat != nat) {
    ATX_PASS(bopscr_layout(pot, nat, pot->tab.NB, pot->tab.NS));
  }
  ATX_CUDA(cudaMemsetAsync(pot->G.ptr, 0, sizeof(double4) * ((size_t)nl->npairs + 1), st));
  if (o.wpa) ATX_CUDA(cudaMemsetAsync(o.wpa, 0, sizeof(double) * 9 * (size_t)nat, st));
  if (nat > 0) {
    ProfScope ps_(ctx, "bop_force");
    k_bopscr_bonds<false><<<(nat + 127) / 128, 128, 0, st>>>(nat, p->Abox, pot->dev, pot->scr_dev.ptr, pot->tab,
                                                             nl->pos4.ptr, nl->seed.ptr, nl->list.ptr,
                                                             pot->flag.ptr, o.role, o.stop);
    ATX_LAUNCHED();
#define BOPSCR_CENTER(KIND)                                                                              \
    k_bopscr_center<KIND><<<nblocks, BOP_BLOCK, 0, st>>>(                                                \
        nat, p->Abox, pot->dev, pot->scr_dev.ptr, pot->tab, nl->pos4.ptr, nl->seed.ptr, nl->list.ptr,   \
        mask_sorted, pot->G.ptr, o.f, pe_own, o.wpa, epb, fpb, wpb, pot->sc.partials.ptr, o.role, o.stop)
    switch (pot->dev.kind) {
      case ATX_BOP_TERSOFF: BOPSCR_CENTER(ATX_BOP_TERSOFF); break;
      case ATX_BOP_KUMAGAI: BOPSCR_CENTER(ATX_BOP_KUMAGAI); break;
      case ATX_BOP_JUSLIN: BOPSCR_CENTER(ATX_BOP_JUSLIN); break;
      default: BOPSCR_CENTER(ATX_BOP_BRENNER);
    }
#undef BOPSCR_CENTER
    ATX_LAUNCHED();
    k_bop_gather<<<(nat + 127) / 128, 128, 0, st>>>(nat, nl->seed.ptr, nl->rev.ptr, pot->G.ptr, pe_own,
                                                    o.f, o.epa, o.role, o.stop);
    ATX_LAUNCHED();
  }
  if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nblocks, o.sums, o.stop));
  return 0;
}

static int bop_compute(atx_bop *pot, atx_particles *p, atx_neighbors *nl, const int *mask_sorted,
                       const PotOut &o, double *epb, double *fpb, double *wpb) {
  atx_ctx *ctx = pot->ctx;
  cudaStream_t st = ctx->stream;
  int nat = nl->nat;
  ATX_PASS(atx_neighbors_ensure_rev(nl));
  ATX_PASS(pot->G.reserve((size_t)nl->npairs + 1));
  ATX_PASS(pot->sc.epa.reserve((size_t)nat + 1));  // pe_own scratch
  double *pe_own = pot->sc.epa.ptr;
  if (pot->screened) return bopscr_compute(pot, p, nl, mask_sorted, o, pe_own, epb, fpb, wpb);
  int nblocks = (nat + BOP_BLOCK - 1) / BOP_BLOCK;
  if (nblocks < 1) nblocks = 1;
  const int ntot = nblocks + ctx->sm_count;   // main-pass blocks + queued-pass blocks
  ATX_PASS(pot->sc.partials.reserve((size_t)ntot * ATX_NSUM));
  ATX_PASS(pot->queue.reserve((size_t)nat + 1));
  if (!o.stop) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr, 0, sizeof(int), st));
  // (re-evaluated at the first call and then at every 16th list build: atoms that outgrow the chosen
  // depth in between are handled by the queued pass, so a stale depth costs time, never correctness)
  if (!o.stop && (pot->nb_cap == 0 || pot->sized_nl != nl || nl->nbuilds - pot->sized_build >= 16)) {
    // host-synchronous call (library mode, MD start, after every list rebuild): pick the bond-table
    // depth from the histogram of bonds per atom.  Atoms beyond it (or that gain bonds while the
    // list is reused) go through the queued pass, so the depth only has to fit the typical atom.
    ATX_CUDA(cudaMemsetAsync(pot->flag.ptr + 8, 0, 32 * sizeof(int), st));
    if (nat > 0) {
      k_bop_count_bonds<<<(nat + 127) / 128, 128, 0, st>>>(nat, p->Abox, pot->dev, nl->pos4.ptr, nl->seed.ptr,
                                                           nl->list.ptr, o.role, pot->flag.ptr + 8);
      ATX_LAUNCHED();
    }
    ATX_PASS(pot->hflag.reserve(64));
    ATX_CUDA(cudaMemcpyAsync(pot->hflag.ptr, pot->flag.ptr + 8, 32 * sizeof(int), cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    const int *hist = pot->hflag.ptr;
    long long total = 0;
    for (int b = 0; b < 32; b++) total += hist[b];
    if (hist[31] > 0 || [&] { for (int b = BOP_NB_MAX + 1; b < 31; b++) if (hist[b]) return true; return false; }()) {
      atx_set_error("Internal neighbor list exhausted, *nebmax* too small: an atom has more than " +
                    std::to_string(BOP_NB_MAX) + " bonds.");
      return ATX_ERROR_UNSPECIFIED;
    }
    // smallest compiled depth that leaves at most 0.1 % of the atoms to the queued pass
    int cap = BOP_NB_MAX;
    for (int d : kBopDepths) {
      long long over = 0;
      for (int b = d + 1; b < 32; b++) over += hist[b];
      if (over * 1000 <= total) { cap = d; break; }
    }
    pot->nb_cap = cap;
    pot->sized_nl = nl;
    pot->sized_build = nl->nbuilds;
  }
  if (o.wpa) ATX_CUDA(cudaMemsetAsync(o.wpa, 0, sizeof(double) * 9 * (size_t)nat, st));
  switch (pot->dev.kind) {
    case ATX_BOP_TERSOFF:
      ATX_PASS(launch_center_nb<ATX_BOP_TERSOFF>(pot, p, nl, mask_sorted, o, pe_own, epb, fpb, wpb, nblocks));
      break;
    case ATX_BOP_KUMAGAI:
      ATX_PASS(launch_center_nb<ATX_BOP_KUMAGAI>(pot, p, nl, mask_sorted, o, pe_own, epb, fpb, wpb, nblocks));
      break;
    case ATX_BOP_JUSLIN:
      ATX_PASS(launch_center_nb<ATX_BOP_JUSLIN>(pot, p, nl, mask_sorted, o, pe_own, epb, fpb, wpb, nblocks));
      break;
    default:
      ATX_PASS(launch_center_nb<ATX_BOP_BRENNER>(pot, p, nl, mask_sorted, o, pe_own, epb, fpb, wpb, nblocks));
  }
  if (o.phase == 1) return 0;   // interior centres only; phase 2 finishes the evaluation
  if (nat > 0) {
    ProfScope ps_(ctx, "bop_gather");
    k_bop_gather<<<(nat + 127) / 128, 128, 0, st>>>(nat, nl->seed.ptr, nl->rev.ptr, pot->G.ptr, pe_own,
                                                    o.f, o.epa, o.role, o.stop);
    ATX_LAUNCHED();
  }
  if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, ntot, o.sums, o.stop));
  return 0;
}

// overflow of the bond table during guarded (batched MD) steps; called by the MD drivers after a sync
// the split evaluation of PotOut::phase exists for the unscreened kernels only
bool atx_bop_supports_split(atx_bop *pot) { return pot && !pot->screened; }

int atx_bop_check_overflow(atx_bop *pot) {
  ATX_PASS(pot->hflag.reserve(4));
  ATX_CUDA(cudaMemcpyAsync(pot->hflag.ptr, pot->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, pot->ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(pot->ctx->stream));
  if (pot->hflag.ptr[0]) {
    atx_set_error("Internal neighbor list exhausted, *nebmax* too small (bond table overflow during MD).");
    return ATX_ERROR_UNSPECIFIED;
  }
  return 0;
}

int atx_bop_compute_device(atx_bop *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o) {
  return bop_compute(pot, p, nl, mask_sorted, o, nullptr, nullptr, nullptr);
}

// device-slot -> host-slot remap of per-bond outputs
__global__ void k_bop_perbond_to_host(int nat, int ncomp, const int *__restrict__ order,
                                      const long long *__restrict__ seed,
                                      const long long *__restrict__ hseed,
                                      const double *__restrict__ in, double *__restrict__ out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  long long w = hseed[order[s]];
  for (long long a = seed[s]; a < seed[s + 1]; a++, w++)
    for (int c = 0; c < ncomp; c++) out[(size_t)w * ncomp + c] = in[(size_t)a * ncomp + c];
}

int atx_neighbors_host_seed(atx_neighbors *nl, DevBuf<long long> &hseed);

int atx_perbond_to_host(atx_ctx *ctx, atx_neighbors *nl, int ncomp, const double *dev_slots,
                        double *host, PinBuf<double> &stage) {
  DevBuf<long long> hseed;
  ATX_PASS(atx_neighbors_host_seed(nl, hseed));
  size_t need = (size_t)(nl->npairs + nl->nat + 1) * ncomp;
  DevBuf<double> out;
  ATX_PASS(out.reserve(need));
  ATX_CUDA(cudaMemsetAsync(out.ptr, 0, sizeof(double) * need, ctx->stream));
  if (nl->nat > 0) {
    k_bop_perbond_to_host<<<(nl->nat + 127) / 128, 128, 0, ctx->stream>>>(
        nl->nat, ncomp, nl->order.ptr, nl->seed.ptr, hseed.ptr, dev_slots, out.ptr);
    ATX_LAUNCHED();
  }
  return atx_accumulate_to_host(ctx, out.ptr, host, need, stage);
}

extern "C" int atx_bop_energy_and_forces(atx_bop *pot, atx_particles *p, atx_neighbors *nl,
                                         const int *mask, double *epot, double *f, double *wpot,
                                         double *epot_per_at, double *epot_per_bond,
                                         double *f_per_bond, double *wpot_per_at,
                                         double *wpot_per_bond) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (!pot->bound) {
    atx_set_error("bind_to has not been called on this potential.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_ctx *ctx = pot->ctx;
  ATX_PASS(atx_neighbors_update(nl, p));
  PotOut o;
  ATX_PASS(pot->sc.f.reserve(3 * (size_t)nl->nat + 3));
  ATX_PASS(pot->sc.sums.reserve(ATX_NSUM));
  ATX_PASS(pot->epa_out.reserve((size_t)nl->nat + 1));
  o.f = pot->sc.f.ptr;
  o.sums = pot->sc.sums.ptr;
  if (epot_per_at) o.epa = pot->epa_out.ptr;
  if (wpot_per_at) {
    ATX_PASS(pot->sc.wpa.reserve(9 * (size_t)nl->nat + 9));
    o.wpa = pot->sc.wpa.ptr;
  }
  const int *mask_sorted = nullptr;
  ATX_PASS(atx_prepare_mask(ctx, nl, pot->sc, mask, &mask_sorted));
  size_t nslots = (size_t)nl->npairs + 1;
  double *epb = nullptr, *fpb = nullptr, *wpb = nullptr;
  if (nl->external) {
    if (epot_per_bond || f_per_bond || wpot_per_bond) {
      // the host-layout slot map needs nl->count, which a caller-supplied list never fills
      atx_set_error("Per-bond outputs are not available with an external neighbour list.");
      return ATX_ERROR_UNSPECIFIED;
    }
    o.role = nl->role_ext.ptr;
  }
  // The screened tables are sized when the list is built; with a library-mode Verlet shell the
  // list can be reused while atoms gain bonds or screening neighbours.  If a table overflows, size
  // again from the current configuration and repeat once.
  for (int attempt = 0;; attempt++) {
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
    ATX_PASS(bop_compute(pot, p, nl, mask_sorted, o, epb, fpb, wpb));
    int h = 0;
    ATX_CUDA(cudaMemcpyAsync(&h, pot->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!h) break;
    if (pot->screened && attempt == 0) {
      pot->nb_cap = 0;   // forces the counting pass
      continue;
    }
    atx_set_error("Internal neighbor list exhausted, *nebmax* too small.");
    return ATX_ERROR_UNSPECIFIED;
  }
  if (epb) ATX_PASS(atx_perbond_to_host(ctx, nl, 1, epb, epot_per_bond, pot->sc.stage));
  if (fpb) ATX_PASS(atx_perbond_to_host(ctx, nl, 3, fpb, f_per_bond, pot->sc.stage));
  if (wpb) ATX_PASS(atx_perbond_to_host(ctx, nl, 9, wpb, wpot_per_bond, pot->sc.stage));
  // epot = 0.5*sum(pe) = sum over directed bonds of e_bond (bop_kernel.f90:1613)
  return atx_finish_to_host(ctx, nl, pot->sc, o, epot, f, wpot, epot_per_at, wpot_per_at);
}

extern "C" int atx_bop_set_store_outputs(atx_bop *pot, int on) {
  if (!pot) return ATX_ERROR_UNSPECIFIED;
  pot->sc.store_outputs = on != 0;
  return 0;
}
