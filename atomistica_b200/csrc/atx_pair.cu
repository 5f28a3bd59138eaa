// Pair potentials on the device: LJCut, Harmonic, DoubleHarmonic.
//
// Replaces the loops of src/potentials/pair_potentials/lj_cut.f90:190-330,
// harmonic.f90:150-225 and double_harmonic.f90:140-230.  The reference walks half of the list (or
// all of it with a factor 1/2) and scatters +df / -df to both partners; here every atom gathers
// over its own full list, which gives the same sums without atomics:
//
//   LJCut           weight of an undirected pair = w_i + w_j (mask weights 0/1), pair energy
//                   0.5*weight*(E - offset), each partner receives half of it; gathered per
//                   directed entry: f_i += omega F r^, e_i += 0.5 omega (E - offset),
//                   omega = (w_i + w_j)/2.  Self images (i == j) carry weight w_i in the
//                   reference -- the same omega.
//   Harmonic        pairs with i > j only (no self images); gathered: entries with i != j,
//                   f_i += F r^, e_i += 0.5 E.
//   DoubleHarmonic  every directed entry with a factor 1/2 (self images included); the per-atom
//                   energies receive en/2 per entry AND partner, i.e. they sum to 2 epot (kept).
//   r6              as Harmonic (i > j).
//   BornMayer       pairs with i <= j (ORIGINAL numbering) including the i == j image entries,
//                   which therefore count with full weight per directed entry; the element test
//                   is asymmetric (born_mayer.f90:222-262: if the lower-numbered atom matches el1
//                   its partner must match el2, otherwise it must match el2 and the partner el1);
//                   only epot and f are produced.
// Virial per directed entry: -0.5 omega F/r dr (x) dr, the same amount goes to wpot_per_at(i).
// Element filters are the bit masks of src/core/filter.f90 (bit k = particle element id k).
#include "atx_potential_common.cuh"

struct PairDev {
  int kind;
  double p[8];      // LJCut: epsilon sigma cutoff offset; Harmonic: k r0 cutoff offset; DoubleHarmonic: k1 r1 k2 r2 cutoff rm
  double cut_sq;
  int el1, el2;
};

struct atx_pair {
  atx_ctx *ctx = nullptr;
  PairDev dev{};
  double cutoff = 0.0;
  bool bound = false;
  PotScratch sc;
};

#define PAIR_BLOCK 128

__global__ void __launch_bounds__(PAIR_BLOCK)
k_pair(int nat, Mat3 A, PairDev P, const double4 *__restrict__ pos4, const int *__restrict__ order,
       const long long *__restrict__ seed, const int2 *__restrict__ list, const int *__restrict__ mask,
       double *__restrict__ f, double *__restrict__ epa, double *__restrict__ wpa,
       double *__restrict__ partials, const unsigned char *__restrict__ role,
       const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * (PAIR_BLOCK / 32)];
  const int s = blockIdx.x * PAIR_BLOCK + threadIdx.x;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (s < nat && (!role || role[s] >= 2)) {
    const double4 pi = pos4[s];
    const int eli = (int)pi.w;
    const bool i1 = (P.el1 >> eli) & 1, i2 = (P.el2 >> eli) & 1;
    const int wi = mask ? (mask[s] != 0) : 1;
    double fx = 0, fy = 0, fz = 0, ei = 0, eat = 0;
    double w[9];
#pragma unroll
    for (int q = 0; q < 9; q++) w[q] = 0.0;
    if (i1 || i2)
      for (long long a = seed[s]; a < seed[s + 1]; a++) {
        const int2 en = list[a];
        const int elj = ATX_ENTRY_EL(en.y);
        const bool j1 = (P.el1 >> elj) & 1, j2 = (P.el2 >> elj) & 1;
        double omega = 1.0;
        if (P.kind == ATX_PAIR_BORN_MAYER) {
          // the reference decides from the LOWER original index of the pair
          const bool self = en.x == s;
          const bool ilow = self || order[s] < order[en.x];
          const bool l1 = ilow ? i1 : j1, l2 = ilow ? i2 : j2, h1 = ilow ? j1 : i1, h2 = ilow ? j2 : i2;
          if (!(l1 ? h2 : (l2 && h1))) continue;
          omega = 1.0;   // force: f_i += F r^ (self images cancel between the +s and -s entries)
        } else {
          if (!((i1 && j2) || (i2 && j1))) continue;
          if ((P.kind == ATX_PAIR_HARMONIC || P.kind == ATX_PAIR_R6) && en.x == s) continue;   // i > j: never a self image
        }
        if (P.kind == ATX_PAIR_LJCUT) {
          const int wj = (en.x == s) ? wi : (mask ? (mask[en.x] != 0) : 1);
          omega = 0.5 * (wi + wj);
          if (omega == 0.0) continue;
        }
        const double4 pj = pos4[en.x];
        double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;   // r_i - r_j + Abox.dc
        if (ATX_NONZERO_SHIFT(en.y)) {
          int sx, sy, sz;
          atx_unpack_shift(en.y, sx, sy, sz);
          double ax, ay, az;
          atx_image_vector(A, sx, sy, sz, ax, ay, az);
          dx += ax; dy += ay; dz += az;
        }
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (!(r2 < P.cut_sq)) continue;
        const double r = sqrt(r2);
        double E, F;   // pair energy and -dE/dr
        if (P.kind == ATX_PAIR_LJCUT) {
          const double sr = P.p[1] / r;
          const double s2 = sr * sr, s6 = s2 * s2 * s2, s12 = s6 * s6;
          E = 4 * P.p[0] * (s12 - s6) - P.p[3];
          F = 24 * P.p[0] * (2 * s12 - s6) / r;
        } else if (P.kind == ATX_PAIR_HARMONIC) {
          F = P.p[0] * (P.p[1] - r);
          E = 0.5 * F * (P.p[1] - r) - P.p[3];
        } else if (P.kind == ATX_PAIR_R6) {
          const double q = P.p[1] + r, q2 = q * q;
          E = P.p[0] / (q2 * q2 * q2);
          F = 6 * E / q;
        } else if (P.kind == ATX_PAIR_BORN_MAYER) {
          const double ex = exp(-r / P.p[1]);
          E = P.p[0] * ex - P.p[3];
          F = (P.p[0] / P.p[1]) * ex;
        } else {
          if (r < P.p[5]) { F = P.p[0] * (P.p[1] - r); E = 0.5 * F * (P.p[1] - r); }
          else { F = P.p[2] * (P.p[3] - r); E = 0.5 * F * (P.p[3] - r); }
        }
        const double c = omega * F / r;
        fx += c * dx; fy += c * dy; fz += c * dz;
        if (P.kind == ATX_PAIR_BORN_MAYER) {
          ei += (en.x == s) ? E : 0.5 * E;   // image entries of an atom with itself count fully, twice
          continue;                          // no virial, no per-atom energy in the reference
        }
        ei += 0.5 * omega * E;
        eat += (P.kind == ATX_PAIR_DOUBLE_HARMONIC) ? E : 0.5 * omega * E;
        const double h = -0.5 * c;
        w[0] += h * dx * dx; w[1] += h * dy * dx; w[2] += h * dz * dx;
        w[3] += h * dx * dy; w[4] += h * dy * dy; w[5] += h * dz * dy;
        w[6] += h * dx * dz; w[7] += h * dy * dz; w[8] += h * dz * dz;
      }
    f[3 * s] = fx; f[3 * s + 1] = fy; f[3 * s + 2] = fz;
    if (epa) epa[s] = eat;
    if (wpa) {
#pragma unroll
      for (int q = 0; q < 9; q++) wpa[9 * (size_t)s + q] = w[q];
    }
    acc[0] = ei;
#pragma unroll
    for (int q = 0; q < 9; q++) acc[1 + q] = w[q];
  } else if (s < nat) {
    f[3 * s] = 0.0; f[3 * s + 1] = 0.0; f[3 * s + 2] = 0.0;
    if (epa) epa[s] = 0.0;
  }
  atx_block_sum<ATX_NSUM, PAIR_BLOCK>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

extern "C" int atx_pair_create(atx_ctx *ctx, const atx_pair_params *par, atx_pair **out) {
  if (ctx) cudaSetDevice(ctx->device);
  if (!ctx || !par || !out) return ATX_ERROR_UNSPECIFIED;
  if (par->kind < ATX_PAIR_LJCUT || par->kind > ATX_PAIR_R6) {
    atx_set_error("atx_pair_create: unknown pair potential.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_pair *pot = new atx_pair();
  pot->ctx = ctx;
  PairDev &D = pot->dev;
  D.kind = par->kind;
  const double *q = par->p;
  if (par->kind == ATX_PAIR_LJCUT) {
    // lj_cut.f90:176-181: offset = 4 eps ((sigma/rc)**12 - (sigma/rc)**6) when shift
    D.p[0] = q[0]; D.p[1] = q[1]; D.p[2] = q[2];
    D.p[3] = par->shift ? 4 * q[0] * (pow(q[1] / q[2], 12) - pow(q[1] / q[2], 6)) : 0.0;
    pot->cutoff = q[2];
  } else if (par->kind == ATX_PAIR_HARMONIC) {
    // harmonic.f90:134-137
    D.p[0] = q[0]; D.p[1] = q[1]; D.p[2] = q[2];
    D.p[3] = par->shift ? 0.5 * q[0] * (q[2] - q[1]) * (q[2] - q[1]) : 0.0;
    pot->cutoff = q[2];
  } else if (par->kind == ATX_PAIR_BORN_MAYER) {
    // born_mayer.f90:169: always shifted to zero at the cutoff
    D.p[0] = q[0]; D.p[1] = q[1]; D.p[2] = q[2];
    D.p[3] = q[0] * exp(-q[2] / q[1]);
    pot->cutoff = q[2];
  } else if (par->kind == ATX_PAIR_R6) {
    D.p[0] = q[0]; D.p[1] = q[1]; D.p[2] = q[2];
    pot->cutoff = q[2];
  } else {
    // double_harmonic.f90:136: rm = (r1 + r2)/2
    D.p[0] = q[0]; D.p[1] = q[1]; D.p[2] = q[2]; D.p[3] = q[3]; D.p[4] = q[4];
    D.p[5] = (q[1] + q[3]) / 2;
    pot->cutoff = q[4];
  }
  D.cut_sq = pot->cutoff * pot->cutoff;
  *out = pot;
  return 0;
}

extern "C" int atx_pair_destroy(atx_pair *pot) {
  delete pot;
  return 0;
}

extern "C" int atx_pair_bind_to(atx_pair *pot, atx_particles *p, atx_neighbors *nl, int el1, int el2) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (!pot) return ATX_ERROR_UNSPECIFIED;
  pot->dev.el1 = el1;
  pot->dev.el2 = el2;
  // lj_cut.f90:168-174: request the cutoff for every element pair the filters select
  if (nl && el1 != 0 && el2 != 0)
    for (int i = 1; i < 32; i++)
      for (int j = 1; j < 32; j++)
        if (((el1 >> i) & 1) && ((el2 >> j) & 1))
          ATX_PASS(atx_neighbors_request_interaction_range_pair(nl, pot->cutoff, i, j));
  pot->bound = true;
  return 0;
}

int atx_pair_compute_device(atx_pair *pot, atx_particles *p, atx_neighbors *nl, const int *mask_sorted,
                            const PotOut &o) {
  atx_ctx *ctx = pot->ctx;
  const int nat = nl->nat;
  const int nblocks = nat > 0 ? (nat + PAIR_BLOCK - 1) / PAIR_BLOCK : 1;
  ATX_PASS(pot->sc.partials.reserve((size_t)nblocks * ATX_NSUM));
  ProfScope ps_(ctx, "pair_force");
  k_pair<<<nblocks, PAIR_BLOCK, 0, ctx->stream>>>(nat, p->Abox, pot->dev, nl->pos4.ptr, nl->order.ptr,
                                                  nl->seed.ptr, nl->list.ptr, mask_sorted, o.f, o.epa, o.wpa,
                                                  pot->sc.partials.ptr, o.role, o.stop);
  ATX_LAUNCHED();
  return o.want_sums ? atx_reduce_partials(ctx, pot->sc.partials.ptr, nblocks, o.sums, o.stop) : 0;
}

extern "C" int atx_pair_energy_and_forces(atx_pair *pot, atx_particles *p, atx_neighbors *nl,
                                          const int *mask, double *epot, double *f, double *wpot,
                                          double *epot_per_at, double *wpot_per_at) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (!pot->bound) {
    atx_set_error("bind_to has not been called on this potential.");
    return ATX_ERROR_UNSPECIFIED;
  }
  if (mask && pot->dev.kind != ATX_PAIR_LJCUT) {
    atx_set_error("This potential does not support masks.");
    return ATX_ERROR_UNSPECIFIED;
  }
  ATX_PASS(atx_neighbors_update(nl, p));
  PotOut o;
  ATX_PASS(atx_prepare_out(pot->ctx, nl, pot->sc, epot_per_at != nullptr, wpot_per_at != nullptr, o));
  const int *mask_sorted = nullptr;
  ATX_PASS(atx_prepare_mask(pot->ctx, nl, pot->sc, mask, &mask_sorted));
  if (nl->external) o.role = nl->role_ext.ptr;
  ATX_PASS(atx_pair_compute_device(pot, p, nl, mask_sorted, o));
  return atx_finish_to_host(pot->ctx, nl, pot->sc, o, epot, f, wpot, epot_per_at, wpot_per_at);
}

extern "C" int atx_pair_set_store_outputs(atx_pair *pot, int on) {
  if (!pot) return ATX_ERROR_UNSPECIFIED;
  pot->sc.store_outputs = on != 0;
  return 0;
}
