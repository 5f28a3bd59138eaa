// Pieces shared by the potential kernels: block reductions of (energy, virial), output handling.
#pragma once

#include "atx_internal.cuh"

#define ATX_NSUM 10  // epot + 9 virial components (column-major 3x3)

// Per-call device outputs of a potential, all in SORTED atom order.
struct PotOut {
  double *f = nullptr;      // (3,nat) overwritten
  double *epa = nullptr;    // (nat) overwritten, optional
  double *wpa = nullptr;    // (9,nat) overwritten, optional
  double *sums = nullptr;   // ATX_NSUM doubles: epot, wpot(3,3)
  const int *stop = nullptr;  // MD driver: kernels return immediately when *stop != 0
  bool want_virial = true;    // NVE stepping does not need wpot
  bool want_sums = true;      // false: the per-block partial sums are not reduced (MD steps whose energy nobody reads)
  // domain decomposition: per sorted atom 2 = owned (everything), 1 = inner ghost (densities /
  // bond-order terms only, no forces), 0 = outer ghost (position only).  nullptr: all owned.
  const unsigned char *role = nullptr;
  // split evaluation (domain decomposition, overlap of the halo exchange with interior work):
  // phase 0 = everything; phase 1 = only the centre atoms of the sorted range [split_lo, split_hi)
  // (rounded inwards to the kernel's block size; no gather, no sums); phase 2 = the remaining centres,
  // the deferred atoms, the gather and the sums.  Phases 1 and 2 of one evaluation must see the same
  // list; positions outside the range may change between them.
  int phase = 0;
  int split_lo = 0, split_hi = 0;
};

// Scratch owned by every potential object for library-mode calls.
struct PotScratch {
  DevBuf<double> f, epa, wpa, sums, partials, out;
  DevBuf<int> mask_sorted, mask_in;
  PinBuf<double> stage;
  PinBuf<double> stage_small;
  // per-atom outputs (f, epot_per_at, wpot_per_at) are STORED into the caller's arrays instead of
  // added (atx_*_set_store_outputs); saves zero-initialising and a host pass for callers that
  // hand in a fresh buffer every call
  bool store_outputs = false;
};

// Block-level sum of v[0..N) over all threads of the block; result valid in thread 0.
template <int N, int BLOCK>
__device__ __forceinline__ void atx_block_sum(double (&v)[N], double *smem /* N*BLOCK/32 */) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) smem[k * (BLOCK / 32) + wid] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      double x = 0.0;
      for (int w = 0; w < BLOCK / 32; w++) x += smem[k * (BLOCK / 32) + w];
      v[k] = x;
    }
  }
}

// deterministic final reduction of per-block partials [nblocks][ATX_NSUM] -> sums[ATX_NSUM]
int atx_reduce_partials(atx_ctx *ctx, const double *partials, int nblocks, double *sums,
                        const int *stop = nullptr);

// sorted -> original order: out_orig[order[s]] (+)= in_sorted[s], ncomp doubles per atom
int atx_unsort(atx_ctx *ctx, int nat, int ncomp, const int *order, const double *in_sorted,
               double *out_orig);
// original -> sorted for int arrays (mask)
int atx_sort_int(atx_ctx *ctx, int nat, const int *order, const int *in_orig, int *out_sorted);

// library-mode epilogue: bring f / per-atom outputs / sums back in original order and ADD them
// into the caller's host arrays.
int atx_finish_to_host(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, const PotOut &o,
                       double *epot, double *f, double *wpot, double *epot_per_at,
                       double *wpot_per_at);
int atx_prepare_out(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, bool want_epa, bool want_wpa,
                    PotOut &o);
int atx_prepare_mask(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, const int *mask_host,
                     const int **mask_sorted);
