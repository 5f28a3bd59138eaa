// Cell-list neighbour build on the device.
//
// Replaces the bodies of
//   neighbors_binning_init    src/python/f90/python_neighbors.f90:765-871  (host, neighbors_geometry)
//   neighbors_binning_update  src/python/f90/python_neighbors.f90:904-959  (k_cell_assign .. k_gather)
//   fill_neighbor_list        src/python/f90/python_neighbors.f90:570-754  (k_pairs<false/true>)
//
// Design: atoms are counting-sorted by cell (stable in the original atom index, which reproduces
// the reference's ascending linked-list order inside a cell); positions are gathered into a
// cell-ordered array of 32-byte records (x,y,z,element) so that one neighbour gather is one DRAM
// sector; the pair list is a CSR in *sorted* numbering with an 8-byte entry {j, packed shift}.
// All potentials work in sorted numbering; the host-layout list (original numbering, 1-based,
// terminator slots) is materialised only on request.
//
// The acceptance predicate is evaluated with explicit round-to-nearest multiplies and adds
// (__dmul_rn/__dadd_rn, never contracted to FMA) in the reference's association order, so the
// pair set is bit-identical to the CPU build.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "atx_internal.cuh"

// ---------------------------------------------------------------------------
// geometry (host): neighbors_binning_init
// ---------------------------------------------------------------------------

static double dot3h(const double *a, const double *b) {
  double s = 0.0;
  for (int i = 0; i < 3; i++) s += a[i] * b[i];
  return s;
}
static void cross3h(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

static void neighbors_geometry(atx_neighbors *nl, const atx_particles *p) {
  const double *A = p->Abox.m, *B = p->Bbox.m;
  double bin_size = nl->cutoff, cell_size[9];
  for (int x = 0; x < 3; x++) {
    double box = std::sqrt(dot3h(&A[3 * x], &A[3 * x]));
    int n = (int)(box / bin_size);
    nl->n_cells[x] = n < 3 ? 3 : n;
  }
  for (int x = 0; x < 3; x++)
    for (int i = 0; i < 3; i++) {
      cell_size[3 * x + i] = A[3 * x + i] / nl->n_cells[x];
      nl->rec_cell_size.m[3 * i + x] = B[3 * i + x] * nl->n_cells[x];  // rec(x,i) = Bbox(x,i)*n(x)
    }
  double nx[3], ny[3], nz[3];
  cross3h(&cell_size[3], &cell_size[6], nx);
  cross3h(&cell_size[6], &cell_size[0], ny);
  cross3h(&cell_size[0], &cell_size[3], nz);
  double cv = dot3h(&cell_size[0], nx);
  double nxx = dot3h(nx, nx), nyy = dot3h(ny, ny), nzz = dot3h(nz, nz);
  for (int i = 0; i < 3; i++) {
    nx[i] = cv * nx[i] / nxx;
    ny[i] = cv * ny[i] / nyy;
    nz[i] = cv * nz[i] / nzz;
  }
  nl->sten[0] = (int)(nl->cutoff / std::sqrt(dot3h(nx, nx))) + 1;
  nl->sten[1] = (int)(nl->cutoff / std::sqrt(dot3h(ny, ny))) + 1;
  nl->sten[2] = (int)(nl->cutoff / std::sqrt(dot3h(nz, nz))) + 1;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

struct Geo {
  Mat3 rec;
  Mat3 A;
  Mat3 cvec;   // cell vectors of the binning grid: column k = Abox(:,k) / n_cells(k)
  int n[3];
  int pbc[3];
  int sten[3];
  double cutoff_sq;
};

// floor(matmul(rec_cell_size, r)), wrapped into the box; shift counts the wraps (+1 per +n)
__device__ __forceinline__ void wrap_cell(const Geo &g, double x, double y, double z, int c[3],
                                          int s[3]) {
  double rr[3] = {x, y, z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double v = __dadd_rn(__dadd_rn(__dmul_rn(g.rec.m[k], rr[0]), __dmul_rn(g.rec.m[3 + k], rr[1])),
                         __dmul_rn(g.rec.m[6 + k], rr[2]));
    long long ck = (long long)floor(v);
    int n = g.n[k];
    int sh = 0;
    if (g.pbc[k]) {
      if (ck < 0) {
        long long q = (-ck + n - 1) / n;
        ck += q * n;
        sh += (int)q;
      } else if (ck >= n) {
        long long q = ck / n;
        ck -= q * n;
        sh -= (int)q;
      }
    } else {
      if (ck < 0) ck = 0;
      if (ck >= n) ck = n - 1;
    }
    c[k] = (int)ck;
    s[k] = sh;
  }
}

__global__ void k_cell_assign(int nat, const double *__restrict__ r, Geo g,
                              int4 *__restrict__ cellshift, int *__restrict__ cell_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nat) return;
  int c[3], s[3];
  wrap_cell(g, r[3 * i], r[3 * i + 1], r[3 * i + 2], c, s);
  int cid = (c[0] * g.n[1] + c[1]) * g.n[2] + c[2];
  cellshift[i] = make_int4(cid, s[0], s[1], s[2]);
  atomicAdd(&cell_count[cid], 1);
}

__global__ void k_cell_scatter(int nat, const int4 *__restrict__ cellshift,
                               const int *__restrict__ cell_start, int *__restrict__ cell_fill,
                               int *__restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nat) return;
  int cid = cellshift[i].x;
  int slot = cell_start[cid] + atomicAdd(&cell_fill[cid], 1);
  order[slot] = i;
}

// ascending original index inside every cell == the reference's linked-list order
__global__ void k_cell_sort(int ncell, const int *__restrict__ cell_start, int *__restrict__ order) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int b = cell_start[c], e = cell_start[c + 1];
  for (int a = b + 1; a < e; a++) {
    int v = order[a];
    int q = a - 1;
    while (q >= b && order[q] > v) {
      order[q + 1] = order[q];
      q--;
    }
    order[q + 1] = v;
  }
}

// scal[5]: bit pattern of the largest |component| of a cell-relative position (non-negative doubles
// order like unsigned integers); scal[6]: set when a wrap shift does not fit the 8-bit packing
// the same for well-filled cells (metals: ~20 atoms per cell): one warp per cell ranks every atom by
// counting the smaller indices of the cell (indices are distinct), no serial insertion
__global__ void __launch_bounds__(256)
k_cell_sort_warp(int ncell, const int *__restrict__ cell_start, int *__restrict__ order) {
  const int c = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= ncell) return;
  const int b = cell_start[c], n = cell_start[c + 1] - b;
  if (n <= 1) return;
  if (n <= 32) {
    const int v = lane < n ? order[b + lane] : 0x7fffffff;
    int rank = 0;
    for (int k = 0; k < n; k++) rank += __shfl_sync(0xffffffffu, v, k) < v;
    __syncwarp();
    if (lane < n) order[b + rank] = v;
  } else if (n <= 256) {
    int v[8], rk[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int t = lane + 32 * q;
      v[q] = t < n ? order[b + t] : 0x7fffffff;
      rk[q] = 0;
    }
    for (int k = 0; k < n; k++) {
      const int u = order[b + k];
#pragma unroll
      for (int q = 0; q < 8; q++) rk[q] += u < v[q];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (lane + 32 * q < n) order[b + rk[q]] = v[q];
  } else if (lane == 0) {
    for (int a = b + 1; a < b + n; a++) {
      int v = order[a];
      int q = a - 1;
      while (q >= b && order[q] > v) {
        order[q + 1] = order[q];
        q--;
      }
      order[q + 1] = v;
    }
  }
}

__global__ void k_gather_sorted(int nat, const double *__restrict__ r, const int *__restrict__ el,
                                const int4 *__restrict__ cellshift, const int *__restrict__ order,
                                double4 *__restrict__ pos4, int4 *__restrict__ sshift,
                                int *__restrict__ inv, Geo g, float4 *__restrict__ posf,
                                long long *__restrict__ scal, double ext_max) {
  // scal[6] is raised when the single-precision records cannot be used: a cell-relative coordinate
  // beyond ext_max (the bound the host derived the band width from; only atoms outside a non-periodic
  // cell exceed it), a wrap count beyond +-59 cells, or a NaN position
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (s < nat) {
    int i = order[s];
    double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
    int e = el ? el[i] : 1;
    pos4[s] = make_double4(x, y, z, (double)e);
    int4 cs = cellshift[i];
    if (sshift) sshift[s] = cs;
    if (inv) inv[i] = s;
    if (posf) {
      // wrapped position minus the origin of the atom's cell: r + A.(shift - c/n)
      int c2 = cs.x % g.n[2], c1 = (cs.x / g.n[2]) % g.n[1], c0 = cs.x / (g.n[2] * g.n[1]);
      double t0 = (double)cs.y - (double)c0 / g.n[0], t1 = (double)cs.z - (double)c1 / g.n[1],
             t2 = (double)cs.w - (double)c2 / g.n[2];
      double px = x + g.A.m[0] * t0 + g.A.m[3] * t1 + g.A.m[6] * t2;
      double py = y + g.A.m[1] * t0 + g.A.m[4] * t1 + g.A.m[7] * t2;
      double pz = z + g.A.m[2] * t0 + g.A.m[5] * t1 + g.A.m[8] * t2;
      const double ext = fmax(fabs(px), fmax(fabs(py), fabs(pz)));
      bad = (abs(cs.y) >= 60) | (abs(cs.z) >= 60) | (abs(cs.w) >= 60) | !(ext <= ext_max);
      int w = bad ? 0 : (atx_pack_shift(cs.y, cs.z, cs.w) | (e << 24));
      posf[s] = make_float4((float)px, (float)py, (float)pz, __int_as_float(w));
    }
  }
  if (posf) {
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicMax((unsigned long long *)&scal[6], 1ull);
  }
}

// Pair search: one thread per atom (sorted order), reference stencil order.
// FILL=false: count; FILL=true: write entries at seed[s].
template <bool FILL>
__global__ void __launch_bounds__(128)
k_pairs(int nat, Geo g, const double4 *__restrict__ pos4, const int4 *__restrict__ sshift,
        const int *__restrict__ cell_start, const int *__restrict__ order,
        int *__restrict__ count, const long long *__restrict__ seed, int2 *__restrict__ list,
        long long *__restrict__ scal, int2 *__restrict__ rows, int rows_cap) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  double4 pi = pos4[s];
  int4 cs = sshift[s];
  int ci[3];
  ci[2] = cs.x % g.n[2];
  ci[1] = (cs.x / g.n[2]) % g.n[1];
  ci[0] = cs.x / (g.n[2] * g.n[1]);
  long long w = FILL ? seed[s] : 0;
  int cnt = 0;
  for (int x = -g.sten[0]; x <= g.sten[0]; x++) {
    int cx = ci[0] + x, sx = cs.y;
    if (g.pbc[0]) {
      while (cx < 0) { cx += g.n[0]; sx += 1; }
      while (cx >= g.n[0]) { cx -= g.n[0]; sx -= 1; }
    } else if (cx < 0 || cx >= g.n[0]) continue;
    for (int y = -g.sten[1]; y <= g.sten[1]; y++) {
      int cy = ci[1] + y, sy = cs.z;
      if (g.pbc[1]) {
        while (cy < 0) { cy += g.n[1]; sy += 1; }
        while (cy >= g.n[1]) { cy -= g.n[1]; sy -= 1; }
      } else if (cy < 0 || cy >= g.n[1]) continue;
      for (int z = -g.sten[2]; z <= g.sten[2]; z++) {
        int cz = ci[2] + z, sz = cs.w;
        if (g.pbc[2]) {
          while (cz < 0) { cz += g.n[2]; sz += 1; }
          while (cz >= g.n[2]) { cz -= g.n[2]; sz -= 1; }
        } else if (cz < 0 || cz >= g.n[2]) continue;
        int cid = (cx * g.n[1] + cy) * g.n[2] + cz;
        int b = cell_start[cid], e = cell_start[cid + 1];
        for (int t = b; t < e; t++) {
          int4 cj = sshift[t];
          int s2x = sx - cj.y, s2y = sy - cj.z, s2z = sz - cj.w;
          const bool zero = (s2x | s2y | s2z) == 0;
          if (t == s && zero) continue;
          double4 pj = pos4[t];
          double dx = __dsub_rn(pi.x, pj.x), dy = __dsub_rn(pi.y, pj.y), dz = __dsub_rn(pi.z, pj.z);
          if (!zero) {
            // + matmul(Abox, shift2); adding the exact zero vector is skipped (bit-identical)
            double ax, ay, az;
            atx_image_vector(g.A, s2x, s2y, s2z, ax, ay, az);
            dx = __dadd_rn(dx, ax); dy = __dadd_rn(dy, ay); dz = __dadd_rn(dz, az);
          }
          double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          if (d2 < g.cutoff_sq) {
            if (FILL) {
              int bad = (abs(s2x) >= ATX_SHIFT_BIAS) | (abs(s2y) >= ATX_SHIFT_BIAS) |
                        (abs(s2z) >= ATX_SHIFT_BIAS);
              if (bad) atomicMax((unsigned long long *)&scal[3], 1ull);
              list[w++] = make_int2(t, atx_pack_shift(s2x, s2y, s2z) | ((int)pj.w << 24));
            } else if (rows && cnt < rows_cap) {
              // single-pass build: park the hit in this atom's fixed-width row; k_rows_to_csr
              // compacts the rows once the offsets are known (no second distance search)
              int bad = (abs(s2x) >= ATX_SHIFT_BIAS) | (abs(s2y) >= ATX_SHIFT_BIAS) |
                        (abs(s2z) >= ATX_SHIFT_BIAS);
              if (bad) atomicMax((unsigned long long *)&scal[3], 1ull);
              rows[(size_t)s * rows_cap + cnt] = make_int2(t, atx_pack_shift(s2x, s2y, s2z) | ((int)pj.w << 24));
            }
            cnt++;
          }
        }
      }
    }
  }
  if (!FILL) count[s] = cnt;
}

// Pair search with a single-precision pre-filter: the same walk as k_pairs, but a candidate costs one
// 16-byte record (cell-relative float position + packed element / wrap shift) and a handful of FP32
// operations.  d2 is first evaluated in float from cell-relative coordinates: |d2_f - d2| < delta
// for every candidate whose classification matters (bound in atx_neighbors_update), so
//   d2_f >= rc^2 + delta -> outside,   d2_f < rc^2 - delta -> inside,
// and only inside the band of half-width delta (a fraction ~1e-5 of the candidates) the reference's
// exact double-precision predicate of k_pairs decides.  The resulting list is bit-identical.
// the reference's predicate in double precision (identical to k_pairs) for candidate t of atom s
__device__ __noinline__ bool pairs_exact(const Geo &g, const double4 *__restrict__ pos4, int s, int t, int code) {
  int s2x, s2y, s2z;
  atx_unpack_shift(code, s2x, s2y, s2z);
  const double4 pi = pos4[s], pj = pos4[t];
  double dx = __dsub_rn(pi.x, pj.x), dy = __dsub_rn(pi.y, pj.y), dz = __dsub_rn(pi.z, pj.z);
  if ((s2x | s2y | s2z) != 0) {
    double ax, ay, az;
    atx_image_vector(g.A, s2x, s2y, s2z, ax, ay, az);
    dx = __dadd_rn(dx, ax); dy = __dadd_rn(dy, ay); dz = __dadd_rn(dz, az);
  }
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return d2 < g.cutoff_sq;
}

template <bool FILL>
__global__ void __launch_bounds__(128)
k_pairs_f32(int nat, const __grid_constant__ Geo g, float lo2, float hi2, const float4 *__restrict__ posf,
            const double4 *__restrict__ pos4, const int4 *__restrict__ sshift,
            const int *__restrict__ cell_start, int *__restrict__ count,
            const long long *__restrict__ seed, int2 *__restrict__ list, long long *__restrict__ scal,
            int2 *__restrict__ rows, int rows_cap) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  const float4 fi = posf[s];
  const int4 cs = sshift[s];
  int ci[3];
  ci[2] = cs.x % g.n[2];
  ci[1] = (cs.x / g.n[2]) % g.n[1];
  ci[0] = cs.x / (g.n[2] * g.n[1]);
  long long w = FILL ? seed[s] : 0;
  int cnt = 0;
  for (int x = -g.sten[0]; x <= g.sten[0]; x++) {
    int cx = ci[0] + x, sx = cs.y;
    if (g.pbc[0]) {
      while (cx < 0) { cx += g.n[0]; sx += 1; }
      while (cx >= g.n[0]) { cx -= g.n[0]; sx -= 1; }
    } else if (cx < 0 || cx >= g.n[0]) continue;
    for (int y = -g.sten[1]; y <= g.sten[1]; y++) {
      int cy = ci[1] + y, sy = cs.z;
      if (g.pbc[1]) {
        while (cy < 0) { cy += g.n[1]; sy += 1; }
        while (cy >= g.n[1]) { cy -= g.n[1]; sy -= 1; }
      } else if (cy < 0 || cy >= g.n[1]) continue;
      const double oxy0 = x * g.cvec.m[0] + y * g.cvec.m[3], oxy1 = x * g.cvec.m[1] + y * g.cvec.m[4],
                   oxy2 = x * g.cvec.m[2] + y * g.cvec.m[5];
      for (int z = -g.sten[2]; z <= g.sten[2]; z++) {
        int cz = ci[2] + z, sz = cs.w;
        if (g.pbc[2]) {
          while (cz < 0) { cz += g.n[2]; sz += 1; }
          while (cz >= g.n[2]) { cz -= g.n[2]; sz -= 1; }
        } else if (cz < 0 || cz >= g.n[2]) continue;
        // the image of the candidate's cell next to i: offset (x,y,z) cells from i's cell, whatever
        // the wrapped index is; relative position of i as seen from that cell's origin
        const float qx = fi.x - (float)(oxy0 + z * g.cvec.m[6]);
        const float qy = fi.y - (float)(oxy1 + z * g.cvec.m[7]);
        const float qz = fi.z - (float)(oxy2 + z * g.cvec.m[8]);
        const int cid = (cx * g.n[1] + cy) * g.n[2] + cz;
        const int b = cell_start[cid], e = cell_start[cid + 1];
        // packed shift2 = (s + w) - cs_j as ONE integer subtraction: the biased bytes are base-256 digits and
        // every digit of the result stays inside [0, 255] (|cs| < 60 is checked when the records are built)
        const int psz = atx_pack_shift(sx, sy, sz) + ATX_SHIFT_ZERO;
#pragma unroll 2
        for (int t = b; t < e; t++) {
          const float4 fj = posf[t];
          const float dxf = qx - fj.x, dyf = qy - fj.y, dzf = qz - fj.z;
          const float d2f = dxf * dxf + dyf * dyf + dzf * dzf;
          if (d2f < hi2) {
            const int wj = __float_as_int(fj.w);
            const int code = psz - (wj & ATX_SHIFT_MASK);
            bool hit = (t != s) | (code != ATX_SHIFT_ZERO);
            if (d2f >= lo2 && hit) hit = pairs_exact(g, pos4, s, t, code);   // band around the cutoff (rare)
            if (hit) {
              const int2 ent = make_int2(t, code | (wj & 0x7f000000));
              if (FILL) list[w++] = ent;
              else if (rows && cnt < rows_cap) rows[(size_t)s * rows_cap + cnt] = ent;
              cnt++;
            }
          }
        }
      }
    }
  }
  if (!FILL) count[s] = cnt;
}


// Pair search for well-filled cells (metals: ~20 atoms per cell, ~500 candidates per atom): one WARP per
// cell.  The (at most 27) stencil cells of the warp's cell, in the reference's order, form one
// concatenated candidate sequence.
//   phase 1: lanes = candidates (two per lane, coalesced 16-byte records, the cell offset folded in once),
//            loop over the cell's atoms i (broadcast from shared memory): the single-precision test in its
//            scalar-product form  fi.pj - (|pj|^2 - hi2)/2 > |fi|^2/2  costs three FMA and a compare per
//            candidate; one ballot per 32 candidates is parked in shared memory.  Candidates in the band
//            around the cutoff (about one per 50 atoms) go through the exact predicate before the ballot.
//   phase 2: lanes = atoms i: every lane expands its own ballots in ascending candidate order -- stencil
//            order, ascending sorted index inside a cell: the reference's order -- and writes its entries.
// ~300 warp instructions per atom instead of ~590 in k_pairs_f32 (where a warp of 32 atoms straddles cells
// with different trip counts and executes the hit path with a few active lanes in nearly every iteration).
// Error bound of the scalar-product form: atx_neighbors_update.  Requires a stencil of +-1 cell.
#define NLC_WPB 4      // warps (cells) per block
#define NLC_CB 8       // 64-candidate groups per batch
#ifndef NLC_MINB
#define NLC_MINB 7
#endif
struct NlcWarp {
  int prefix[28];           // exclusive prefix of the segment lengths, [27] = total
  int tb[27];               // first sorted index of the segment minus its prefix
  int P[27];                // wrap of the segment in the linear form of atx_pack_shift
  float ox[27], oy[27], oz[27];
  float4 fi[32];            // atoms of the cell: x, y, z, |f|^2 / 2
  unsigned mw[NLC_CB * 2][32];   // ballots: word 2 cp + h holds candidates 64 cp + 32 h ... + 31 of atom [ii]
  int2 cand[NLC_CB * 64];   // per candidate of the batch: sorted index, entry word minus the atom's own shift
};

template <bool FILL>
__global__ void __launch_bounds__(32 * NLC_WPB, NLC_MINB)
k_pairs_coop(int ncell, const __grid_constant__ Geo g, float hi2, float bw, const float4 *__restrict__ posf,
             const double4 *__restrict__ pos4, const int4 *__restrict__ sshift,
             const int *__restrict__ cell_start, int *__restrict__ count,
             const long long *__restrict__ seed, int2 *__restrict__ list, int2 *__restrict__ rows, int rows_cap) {
  __shared__ NlcWarp wsm[NLC_WPB];
  const int lane = threadIdx.x & 31;
  NlcWarp &W = wsm[threadIdx.x >> 5];
  const int c = blockIdx.x * NLC_WPB + (threadIdx.x >> 5);
  if (c >= ncell) return;
  const int cb = cell_start[c], ce = cell_start[c + 1];
  if (cb == ce) return;
  // ---- the 27 segments (lane k: x = k / 9 - 1 outermost, z = k % 3 - 1 innermost, as the reference loops)
  {
    int len = 0, b = 0, P = 0;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (lane < 27) {
      const int d[3] = {lane / 9 - 1, (lane / 3) % 3 - 1, lane % 3 - 1};
      int cc[3] = {c / (g.n[2] * g.n[1]), (c / g.n[2]) % g.n[1], c % g.n[2]};
      int w[3] = {0, 0, 0};
      bool ok = true;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        cc[k] += d[k];
        if (g.pbc[k]) {
          if (cc[k] < 0) { cc[k] += g.n[k]; w[k] = 1; }
          else if (cc[k] >= g.n[k]) { cc[k] -= g.n[k]; w[k] = -1; }
        } else if (cc[k] < 0 || cc[k] >= g.n[k]) ok = false;
      }
      if (ok) {
        const int cid = (cc[0] * g.n[1] + cc[1]) * g.n[2] + cc[2];
        b = cell_start[cid];
        len = cell_start[cid + 1] - b;
      }
      P = w[0] + (w[1] << 8) + (w[2] << 16);
      ox = (float)(d[0] * g.cvec.m[0] + d[1] * g.cvec.m[3] + d[2] * g.cvec.m[6]);
      oy = (float)(d[0] * g.cvec.m[1] + d[1] * g.cvec.m[4] + d[2] * g.cvec.m[7]);
      oz = (float)(d[0] * g.cvec.m[2] + d[1] * g.cvec.m[5] + d[2] * g.cvec.m[8]);
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane < 27) {
      W.prefix[lane] = incl - len;
      W.tb[lane] = b - (incl - len);
      W.P[lane] = P;
      W.ox[lane] = ox; W.oy[lane] = oy; W.oz[lane] = oz;
    }
    if (lane == 26) W.prefix[27] = incl;
  }
  __syncwarp();
  const int total = W.prefix[27];
  const int home0 = W.prefix[13];   // segment 13 = the cell itself with zero wrap: starts at cb

  for (int i0 = cb; i0 < ce; i0 += 32) {
    const int ni = min(32, ce - i0);
    int psz_own = 0;
    if (lane < ni) {
      const float4 f = posf[i0 + lane];
      W.fi[lane] = make_float4(f.x, f.y, f.z, 0.5f * (f.x * f.x + f.y * f.y + f.z * f.z));
      const int4 cs = sshift[i0 + lane];
      psz_own = atx_pack_shift(cs.y, cs.z, cs.w) + ATX_SHIFT_ZERO;
    }
    // phase 2 lays lpa lanes over every atom: lane -> (atom ii, part sub of the batch's ballot words)
    const int lpa = min(NLC_CB, 32 / ni);
    const int sub = lane / ni, ii2 = lane - sub * ni;
    const bool act2 = sub < lpa;
    const int s2 = i0 + ii2;
    const int psz2 = __shfl_sync(0xffffffffu, psz_own, ii2);
    int2 *dst = nullptr;
    if (act2) dst = FILL ? list + seed[s2] : (rows ? rows + (size_t)s2 * rows_cap : nullptr);
    const int cap2 = FILL ? 0x7fffffff : (rows ? rows_cap : 0);
    __syncwarp();
    const int gself = home0 + (i0 + lane - cb);   // atom i0 + lane itself in the candidate sequence
    int cnt = 0, seg0 = 0, seg1 = 0;
    for (int g0 = 0; g0 < total; g0 += NLC_CB * 64) {
      const int nb = min(NLC_CB, (total - g0 + 63) >> 6);
      // ---- phase 1
      for (int cp = 0; cp < nb; cp++) {
        const int gi0 = g0 + cp * 64 + lane, gi1 = gi0 + 32;
        float p0x = 0.f, p0y = 0.f, p0z = 0.f, nh0 = -INFINITY, p1x = 0.f, p1y = 0.f, p1z = 0.f, nh1 = -INFINITY;
        int t0 = 0, t1 = 0, a0 = 0, a1 = 0;
        if (gi0 < total) {
          while (gi0 >= W.prefix[seg0 + 1]) seg0++;
          t0 = W.tb[seg0] + gi0;
          const float4 fj = posf[t0];
          p0x = fj.x + W.ox[seg0]; p0y = fj.y + W.oy[seg0]; p0z = fj.z + W.oz[seg0];
          nh0 = -0.5f * ((p0x * p0x + p0y * p0y + p0z * p0z) - hi2);
          const int wj = __float_as_int(fj.w);
          a0 = W.P[seg0] - (wj & ATX_SHIFT_MASK);
          W.cand[cp * 64 + lane] = make_int2(t0, a0 + (wj & 0x7f000000));
        }
        if (gi1 < total) {
          while (gi1 >= W.prefix[seg1 + 1]) seg1++;
          t1 = W.tb[seg1] + gi1;
          const float4 fj = posf[t1];
          p1x = fj.x + W.ox[seg1]; p1y = fj.y + W.oy[seg1]; p1z = fj.z + W.oz[seg1];
          nh1 = -0.5f * ((p1x * p1x + p1y * p1y + p1z * p1z) - hi2);
          const int wj = __float_as_int(fj.w);
          a1 = W.P[seg1] - (wj & ATX_SHIFT_MASK);
          W.cand[cp * 64 + 32 + lane] = make_int2(t1, a1 + (wj & 0x7f000000));
        }
#pragma unroll 2
        for (int ii = 0; ii < ni; ii++) {
          const float4 f = W.fi[ii];
          const float x0 = fmaf(f.z, p0z, fmaf(f.y, p0y, fmaf(f.x, p0x, nh0)));
          const float x1 = fmaf(f.z, p1z, fmaf(f.y, p1y, fmaf(f.x, p1x, nh1)));
          bool in0 = x0 > f.w, in1 = x1 > f.w;
          const float top = f.w + bw;
          const bool band0 = in0 && x0 <= top, band1 = in1 && x1 <= top;
          if (__any_sync(0xffffffffu, band0 || band1)) {
            // the band around the cutoff: the reference's double-precision predicate decides
            const int4 cs = sshift[i0 + ii];
            const int psz = atx_pack_shift(cs.y, cs.z, cs.w) + ATX_SHIFT_ZERO;
            if (band0) in0 = pairs_exact(g, pos4, i0 + ii, t0, psz + a0);
            if (band1) in1 = pairs_exact(g, pos4, i0 + ii, t1, psz + a1);
          }
          const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
          if (lane < 2) W.mw[cp * 2 + lane][ii] = lane ? m1 : m0;
        }
      }
      __syncwarp();
      // the atom itself (same cell, zero wrap) is not its own neighbour
      {
        const unsigned rel = (unsigned)(gself - g0);
        if (lane < ni && rel < (unsigned)(NLC_CB * 64)) W.mw[rel >> 5][lane] &= ~(1u << (rel & 31));
      }
      __syncwarp();
      // ---- phase 2: every lane expands its share of the ballot words of its atom, in ascending order
      if (act2) {
        const int q_lo = 2 * ((sub * nb) / lpa), q_hi = 2 * (((sub + 1) * nb) / lpa);
        int before = 0, tot = 0;
        for (int q = 0; q < 2 * nb; q++) {
          const int pc = __popc(W.mw[q][ii2]);
          tot += pc;
          if (q < q_lo) before += pc;
        }
        int pos = cnt + before;
        cnt += tot;
        int q = q_lo;
        unsigned cur = 0;
        for (;;) {
          while (cur == 0 && q < q_hi) cur = W.mw[q++][ii2];
          if (cur == 0) break;
          const int k = __ffs(cur) - 1;
          cur &= cur - 1;
          int2 e = W.cand[(q - 1) * 32 + k];
          e.y += psz2;
          if (pos < cap2) dst[pos] = e;
          pos++;
        }
      }
      __syncwarp();
    }
    if (!FILL && lane < ni) count[i0 + lane] = cnt;   // lanes < ni are part 0 of atom i0 + lane
    __syncwarp();
  }
}

// nebmax = max count, i_last = largest original index (1-based) of an atom that has a pair.
// (Per-atom atomicMax on two addresses used to cost more than the pair search itself: L2 atomics
// serialise per address.)
__global__ void __launch_bounds__(256)
k_count_stats(int nat, const int *__restrict__ count, const int *__restrict__ order,
              long long *__restrict__ scal) {
  int mc = 0, ml = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nat; s += gridDim.x * blockDim.x) {
    int c = count[s];
    mc = max(mc, c);
    if (c > 0) ml = max(ml, order[s] + 1);
  }
  mc = __reduce_max_sync(0xffffffffu, mc);
  ml = __reduce_max_sync(0xffffffffu, ml);
  if ((threadIdx.x & 31) == 0 && mc > 0) {
    atomicMax((unsigned long long *)&scal[1], (unsigned long long)mc);
    atomicMax((unsigned long long *)&scal[2], (unsigned long long)ml);
  }
}

// rows (fixed width, filled by the counting pass) -> CSR; 8 lanes per atom
__global__ void k_rows_to_csr(int nat, const int2 *__restrict__ rows, int rows_cap,
                              const long long *__restrict__ seed, int2 *__restrict__ list) {
  const int s = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const int lane = threadIdx.x & 7;
  if (s >= nat) return;
  const long long b = seed[s];
  const int n = (int)(seed[s + 1] - b);
  const int2 *row = rows + (size_t)s * rows_cap;
  for (int k = lane; k < n; k += 8) list[b + k] = row[k];
}

// refresh sorted positions after the atoms moved (list kept)
__global__ void k_refresh_pos(int nat, const double *__restrict__ r, const int *__restrict__ order,
                              double4 *__restrict__ pos4) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  int i = order[s];
  double4 v = pos4[s];
  v.x = r[3 * i];
  v.y = r[3 * i + 1];
  v.z = r[3 * i + 2];
  pos4[s] = v;
}

// library-mode Verlet shell: refresh the sorted positions and record the largest squared displacement
// since the list was built (bit pattern of a non-negative double orders like an unsigned integer)
__global__ void k_refresh_check(int nat, const double *__restrict__ r, const int *__restrict__ order,
                                const double4 *__restrict__ pos_build, double4 *__restrict__ pos4,
                                unsigned long long *__restrict__ max_d2) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (s < nat) {
    int i = order[s];
    double4 v = pos4[s], b = pos_build[s];
    v.x = r[3 * i];
    v.y = r[3 * i + 1];
    v.z = r[3 * i + 2];
    pos4[s] = v;
    double dx = v.x - b.x, dy = v.y - b.y, dz = v.z - b.z;
    d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 == d2)) d2 = 1e300;  // NaN positions force a rebuild (which reports them)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if ((threadIdx.x & 31) == 0 && d2 > 0.0) atomicMax(max_d2, (unsigned long long)__double_as_longlong(d2));
}

// reverse slot: for entry a = (i -> j, shift) the entry b = (j -> i, -shift)
__global__ void k_reverse_index(int nat, const long long *__restrict__ seed,
                                const int2 *__restrict__ list, int *__restrict__ rev) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  for (long long a = seed[s]; a < seed[s + 1]; a++) {
    int2 e = list[a];
    int sx, sy, sz;
    atx_unpack_shift(e.y, sx, sy, sz);
    int want = atx_pack_shift(-sx, -sy, -sz);
    int found = -1;
    for (long long b = seed[e.x]; b < seed[e.x + 1]; b++) {
      int2 q = list[b];
      if (q.x == s && (q.y & ATX_SHIFT_MASK) == want) { found = (int)b; break; }
    }
    rev[a] = found;
  }
}

// host-layout list ---------------------------------------------------------

__global__ void k_host_counts(int nat, const int *__restrict__ inv, const int *__restrict__ count,
                              int *__restrict__ hcount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nat) return;
  hcount[i] = count[inv[i]] + 1;  // + terminator slot
}

__global__ void k_host_fill(int nat, const int *__restrict__ inv, const int *__restrict__ order,
                            const long long *__restrict__ seed, const int2 *__restrict__ list,
                            const long long *__restrict__ hseed, long long *__restrict__ out_seed,
                            long long *__restrict__ out_last, int *__restrict__ out_nb,
                            int *__restrict__ out_dc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nat) return;
  if (i == nat) {
    out_seed[nat] = hseed[nat] + 1;
    return;
  }
  int s = inv[i];
  long long w = hseed[i];  // 0-based slot
  out_seed[i] = w + 1;
  for (long long a = seed[s]; a < seed[s + 1]; a++, w++) {
    int2 e = list[a];
    int sx, sy, sz;
    atx_unpack_shift(e.y, sx, sy, sz);
    out_nb[w] = order[e.x] + 1;
    out_dc[3 * w] = sx;
    out_dc[3 * w + 1] = sy;
    out_dc[3 * w + 2] = sz;
  }
  out_last[i] = w;  // 1-based inclusive == 0-based exclusive
  out_nb[w] = 0;
}

// ---------------------------------------------------------------------------
// API
// ---------------------------------------------------------------------------

extern "C" int atx_neighbors_create(atx_ctx *ctx, int avgn, atx_neighbors **nl) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (!ctx || !nl) return ATX_ERROR_UNSPECIFIED;
  *nl = new atx_neighbors();
  (*nl)->ctx = ctx;
  (*nl)->avgn = avgn;
  return 0;
}

extern "C" int atx_neighbors_destroy(atx_neighbors *nl) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  delete nl;
  return 0;
}

extern "C" int atx_neighbors_request_interaction_range(atx_neighbors *nl, double cutoff) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  // python_neighbors.f90:381-423: any request tears the list down
  if (cutoff > nl->interaction_range) nl->interaction_range = cutoff;
  nl->initialized = false;
  return 0;
}

// request_interaction_range(nl, cutoff, el1, el2) (lammps_neighbors.f90:223-251, python_neighbors.f90:381-423):
// the list itself is built with the largest range; the per-pair values are what neighbors_get_cutoff hands
// to a host that builds the list (LAMMPS sets cutsq(i,j) from them)
extern "C" int atx_neighbors_request_interaction_range_pair(atx_neighbors *nl, double cutoff, int el1, int el2) {
  if (!nl) return ATX_ERROR_UNSPECIFIED;
  if (el1 >= 1 && el1 < 32 && el2 >= 1 && el2 < 32) {
    if (cutoff > nl->pair_range[el1][el2]) nl->pair_range[el1][el2] = cutoff;
    if (cutoff > nl->pair_range[el2][el1]) nl->pair_range[el2][el1] = cutoff;
  }
  return atx_neighbors_request_interaction_range(nl, cutoff);
}

extern "C" int atx_neighbors_get_pair_range(atx_neighbors *nl, int el1, int el2, double *range) {
  if (!nl || !range) return ATX_ERROR_UNSPECIFIED;
  *range = (el1 >= 1 && el1 < 32 && el2 >= 1 && el2 < 32) ? nl->pair_range[el1][el2] : 0.0;
  return 0;
}

extern "C" int atx_neighbors_set_verlet_shell(atx_neighbors *nl, double verlet_shell) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  nl->verlet_shell = verlet_shell;
  nl->initialized = false;
  return 0;
}

static Geo make_geo(const atx_neighbors *nl, const atx_particles *p) {
  Geo g;
  g.rec = nl->rec_cell_size;
  g.A = p->Abox;
  for (int k = 0; k < 3; k++) {
    g.n[k] = nl->n_cells[k];
    g.pbc[k] = p->pbc[k];
    g.sten[k] = nl->sten[k];
  }
  for (int k = 0; k < 3; k++)
    for (int c = 0; c < 3; c++) g.cvec.m[3 * k + c] = p->Abox.m[3 * k + c] / nl->n_cells[k];
  g.cutoff_sq = nl->cutoff * nl->cutoff;
  return g;
}

extern "C" int atx_neighbors_update(atx_neighbors *nl, atx_particles *p) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  atx_ctx *ctx = nl->ctx;
  cudaStream_t st = ctx->stream;
  if (nl->external) {
    // the host owns the list (lammps_neighbors.f90:205-217: update is a no-op); only the sorted
    // position records follow the particles
    if (nl->bound != p || nl->nat != p->nat) {
      atx_set_error("The external neighbour list was set for another particles object; call atx_neighbors_set_external again.");
      return ATX_ERROR_UNSPECIFIED;
    }
    if (nl->p_rev != p->pos_rev) ATX_PASS(atx_neighbors_refresh_positions(nl, p));
    return 0;
  }
  if (nl->bound != p || nl->nat != p->nat) nl->initialized = false;
  if (nl->initialized && nl->p_rev == p->pos_rev && nl->cell_rev == p->cell_rev) return 0;

  const int nat = p->nat;
  ProfScope ps_build_(ctx, "nl_update");
  if (nl->initialized && nl->verlet_shell > 0.0 && nl->cell_rev == p->cell_rev && nl->p_rev >= 0 &&
      nl->el_rev == p->el_rev && nl->pos_build.cap >= (size_t)nat && nat > 0) {
    // Verlet shell (neighbors.f90:520-560 / python_neighbors.f90:570-600: the list is kept while
    // no atom moved further than verlet_shell/2).  The reference accumulates per-step maxima; here
    // the displacement since the build is measured directly, which is the exact form of that rule.
    ProfScope ps_(ctx, "nl_refresh_check");
    ATX_PASS(nl->scal.reserve(8));
    ATX_CUDA(cudaMemsetAsync(nl->scal.ptr + 4, 0, sizeof(long long), st));
    k_refresh_check<<<(nat + 255) / 256, 256, 0, st>>>(nat, p->rptr(), nl->order.ptr, nl->pos_build.ptr,
                                                       nl->pos4.ptr, (unsigned long long *)(nl->scal.ptr + 4));
    ATX_LAUNCHED();
    long long bits = 0;
    ATX_CUDA(cudaMemcpyAsync(&bits, nl->scal.ptr + 4, sizeof(long long), cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    double d2;
    memcpy(&d2, &bits, sizeof(double));
    if (d2 < 0.25 * nl->verlet_shell * nl->verlet_shell) {
      nl->p_rev = p->pos_rev;
      nl->nreused++;
      return 0;
    }
  }
  if (!nl->initialized) {
    nl->cutoff = nl->interaction_range + nl->verlet_shell;
    if (nl->cutoff <= 0.0) {
      atx_set_error("Cutoff needs to be larger than zero.");
      return ATX_ERROR_UNSPECIFIED;
    }
    nl->capacity = (long long)nat * nl->avgn;
    nl->bound = p;
    nl->nat = nat;
    nl->cell_rev = -1;
  }
  if (nl->cell_rev != p->cell_rev || !nl->initialized) neighbors_geometry(nl, p);
  nl->initialized = true;
  nl->rev_valid = false;

  long long ncell_ll = (long long)nl->n_cells[0] * nl->n_cells[1] * nl->n_cells[2];
  if (ncell_ll > 2000000000ll) {
    atx_set_error("Too many binning cells.");
    return ATX_ERROR_UNSPECIFIED;
  }
  int ncell = (int)ncell_ll;
  Geo g = make_geo(nl, p);

  ATX_PASS(nl->cellshift.reserve(nat + 1));
  ATX_PASS(nl->cell_count.reserve(ncell + 1));
  ATX_PASS(nl->cell_start.reserve(ncell + 1));
  ATX_PASS(nl->cell_fill.reserve(ncell + 1));
  ATX_PASS(nl->order.reserve(nat + 1));
  ATX_PASS(nl->inv.reserve(nat + 1));
  ATX_PASS(nl->pos4.reserve(nat + 1));
  ATX_PASS(nl->sshift.reserve(nat + 1));
  ATX_PASS(nl->count.reserve(nat + 1));
  ATX_PASS(nl->seed.reserve(nat + 2));
  ATX_PASS(nl->scal.reserve(8));
  static const bool f32_env = !(getenv("ATX_NL_F32") && atoi(getenv("ATX_NL_F32")) == 0);
  const bool use_f32 = f32_env;
  if (use_f32) ATX_PASS(nl->posf.reserve(nat + 1));

  ATX_CUDA(cudaMemsetAsync(nl->cell_count.ptr, 0, sizeof(int) * (ncell + 1), st));
  ATX_CUDA(cudaMemsetAsync(nl->cell_fill.ptr, 0, sizeof(int) * (ncell + 1), st));
  ATX_CUDA(cudaMemsetAsync(nl->scal.ptr, 0, sizeof(long long) * 8, st));
  ATX_CUDA(cudaMemsetAsync(nl->count.ptr, 0, sizeof(int) * (nat + 1), st));

  // single-precision pre-filter of the pair search: half-width of the band in which the exact
  // predicate decides (error analysis at k_pairs_f32; u = 2^-24, every term with a factor 2 of safety).
  // The bound on the cell-relative coordinates is the cell diagonal sum -- every atom inside the
  // (periodic or non-periodic) cell obeys it; the gather kernel checks it and the search is redone
  // with the exact kernel in the rare case it does not hold (no host round trip on the normal path).
  float lo2 = 0.f, hi2 = 0.f, hi2c = 0.f, bwc = 0.f;
  bool coop = false;
  double ext_max = 0.0;
  nl->f32_delta = -1.0;
  if (use_f32 && nat > 0) {
    const double *A = p->Abox.m;
    double offmax = 0.0;
    for (int k = 0; k < 3; k++) {
      double len = std::sqrt(A[3 * k] * A[3 * k] + A[3 * k + 1] * A[3 * k + 1] + A[3 * k + 2] * A[3 * k + 2]);
      offmax += nl->sten[k] * len / nl->n_cells[k];
      ext_max += 1.02 * len / nl->n_cells[k];
    }
    const double eps = std::ldexp(1.0, -23), R2 = g.cutoff_sq, R = std::sqrt(R2);
    const double e_d = eps * (5.0 * ext_max + 3.0 * offmax);
    const double delta = 2.0 * (2.0 * std::sqrt(3.0) * R * e_d * 1.01 + 3.0 * e_d * e_d + 2.0 * eps * R2 * 1.01) +
                         2.0 * eps * R2;
    if (delta <= 0.05 * R2) {
      nl->f32_delta = delta;
      lo2 = std::nextafterf((float)(R2 - delta), -1.0f);
      hi2 = std::nextafterf((float)(R2 + delta), 3.0e38f);
    }
    // warp-per-cell kernel (well-filled cells, stencil of +-1 cell): the test is evaluated in its
    // scalar-product form  d2 = |fi|^2 + |pj|^2 - 2 fi.pj  with  |fi| <= E = sqrt(3) ext_max (the gather kernel
    // checks the components against ext_max) and |pj| <= P = E + offmax.  With u = 2^-24:
    //   |fi|^2: 3 roundings + the rounding of fi itself                  <=  5 u E^2
    //   |pj|^2: 3 roundings + rounding of fj, of the offset, of the sum  <=  7 u P^2
    //   2 fi.pj: 3 FMA roundings + the roundings of both vectors         <= 12 u E P
    //   -(|pj|^2 - hi2)/2 and the three FMA partial sums (in d2 units)   <=  u (P^2 + R^2) + 3 u (P^2 + R^2 + 2 E P)
    // sum <= 34 u P^2 + 4 u R^2; the band half-width is twice that (+5 %), the exact predicate decides inside.
    const bool coop_env = !(getenv("ATX_NL_COOP") && atoi(getenv("ATX_NL_COOP")) == 0);
    if (nl->f32_delta > 0.0 && coop_env && nl->sten[0] == 1 && nl->sten[1] == 1 && nl->sten[2] == 1 &&
        (long long)nat >= 6ll * ncell) {
      const double u = std::ldexp(1.0, -24), E = std::sqrt(3.0) * ext_max, Pm = E + offmax;
      const double dc = 2.1 * (34.0 * u * Pm * Pm + 4.0 * u * R2);
      if (dc <= 0.05 * R2) {
        coop = true;
        hi2c = std::nextafterf((float)(R2 + dc), 3.0e38f);
        bwc = std::nextafterf((float)dc, 3.0e38f);
      }
    }
  }
  bool f32 = nl->f32_delta > 0.0;

  const int TB = 256;
  int gb = (nat + TB - 1) / TB;
  if (nat > 0) {
    k_cell_assign<<<gb, TB, 0, st>>>(nat, p->rptr(), g, nl->cellshift.ptr, nl->cell_count.ptr);
    ATX_LAUNCHED();
  }
  ATX_PASS(atx_scan_int(ctx, nl->cell_count.ptr, nl->cell_start.ptr, ncell + 1));
  if (nat > 0) {
    k_cell_scatter<<<gb, TB, 0, st>>>(nat, nl->cellshift.ptr, nl->cell_start.ptr, nl->cell_fill.ptr,
                                      nl->order.ptr);
    ATX_LAUNCHED();
    if ((long long)nat > 3ll * ncell) {
      const long long nthr = (long long)ncell * 32;
      k_cell_sort_warp<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(ncell, nl->cell_start.ptr, nl->order.ptr);
    } else {
      k_cell_sort<<<(ncell + TB - 1) / TB, TB, 0, st>>>(ncell, nl->cell_start.ptr, nl->order.ptr);
    }
    ATX_LAUNCHED();
    k_gather_sorted<<<gb, TB, 0, st>>>(nat, p->rptr(), p->el.cap ? p->el.ptr : nullptr,
                                       nl->cellshift.ptr, nl->order.ptr, nl->pos4.ptr,
                                       nl->sshift.ptr, nl->inv.ptr, g, f32 ? nl->posf.ptr : nullptr,
                                       nl->scal.ptr, ext_max);
    ATX_LAUNCHED();
  }
  long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // Single-pass build from the second build on: the previous build's longest list (+ margin) sizes
  // fixed-width rows the counting pass fills; if an atom outgrows its row the classic second
  // search pass runs instead.  Scratch is bounded to 8 GiB.
  int rows_cap = 0;
  if (nl->nbuilds > 0 && nl->nebmax > 0 && nat > 0) {
    int cap = nl->nebmax + (nl->nebmax / 8 > 8 ? nl->nebmax / 8 : 8);
    if ((size_t)nat * cap * sizeof(int2) <= ((size_t)8 << 30) && nl->rows.reserve((size_t)nat * cap) == 0)
      rows_cap = cap;
  }
  for (int attempt = 0; attempt < 2; attempt++) {
  if (nat > 0) {
    {
      ProfScope ps_(ctx, "nl_pairs_count");
      if (f32 && coop)
        k_pairs_coop<false><<<(ncell + NLC_WPB - 1) / NLC_WPB, 32 * NLC_WPB, 0, st>>>(
            ncell, g, hi2c, bwc, nl->posf.ptr, nl->pos4.ptr, nl->sshift.ptr, nl->cell_start.ptr, nl->count.ptr,
            nullptr, nullptr, rows_cap > 0 ? nl->rows.ptr : nullptr, rows_cap);
      else if (f32)
        k_pairs_f32<false><<<(nat + 127) / 128, 128, 0, st>>>(nat, g, lo2, hi2, nl->posf.ptr, nl->pos4.ptr,
                                                              nl->sshift.ptr, nl->cell_start.ptr, nl->count.ptr,
                                                              nullptr, nullptr, nl->scal.ptr,
                                                              rows_cap > 0 ? nl->rows.ptr : nullptr, rows_cap);
      else
      k_pairs<false><<<(nat + 127) / 128, 128, 0, st>>>(nat, g, nl->pos4.ptr, nl->sshift.ptr,
                                                        nl->cell_start.ptr, nl->order.ptr,
                                                        nl->count.ptr, nullptr, nullptr, nl->scal.ptr,
                                                        rows_cap > 0 ? nl->rows.ptr : nullptr, rows_cap);
    }
    ATX_LAUNCHED();
    k_count_stats<<<ctx->sm_count * 2, 256, 0, st>>>(nat, nl->count.ptr, nl->order.ptr, nl->scal.ptr);
    ATX_LAUNCHED();
  }
  ATX_PASS(atx_scan_int_to_ll(ctx, nl->count.ptr, nl->seed.ptr, nat + 1));
  ATX_CUDA(cudaMemcpyAsync(&h[0], nl->seed.ptr + nat, sizeof(long long), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaMemcpyAsync(&h[1], nl->scal.ptr + 1, 6 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  if (f32 && h[6]) {
    // the single-precision records were not usable (see k_gather_sorted): exact search instead
    f32 = false;
    nl->f32_delta = -1.0;
    ATX_CUDA(cudaMemsetAsync(nl->scal.ptr, 0, sizeof(long long) * 8, st));
    continue;
  }
  break;
  }
  nl->npairs = h[0];
  nl->nebmax = (int)h[1];
  long long i_last = h[2];  // 1-based original index of the last atom that has a pair
  // python_neighbors.f90:716-718: slot of the last pair is npairs + (i_last-1) (1-based) and must
  // stay below the fixed capacity nat*avgn
  if (nl->npairs > 0 && nl->npairs + (i_last - 1) >= nl->capacity) {
    atx_set_error("Neighbor list overflow. Current neighbor list position is " +
                  std::to_string(nl->npairs + i_last - 1) +
                  " while the size of this chunk runs from 1 to " + std::to_string(nl->capacity) +
                  ".");
    nl->initialized = false;
    return ATX_ERROR_UNSPECIFIED;
  }
  ATX_PASS(nl->list.reserve((size_t)nl->npairs + 1));
  if (nat > 0 && nl->npairs > 0) {
    {
      ProfScope ps_(ctx, "nl_pairs_fill");
      if (rows_cap > 0 && nl->nebmax <= rows_cap) {
        const long long nthr = (long long)nat * 8;
        k_rows_to_csr<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(nat, nl->rows.ptr, rows_cap, nl->seed.ptr,
                                                                      nl->list.ptr);
      } else if (f32 && coop) {
        k_pairs_coop<true><<<(ncell + NLC_WPB - 1) / NLC_WPB, 32 * NLC_WPB, 0, st>>>(
            ncell, g, hi2c, bwc, nl->posf.ptr, nl->pos4.ptr, nl->sshift.ptr, nl->cell_start.ptr, nl->count.ptr,
            nl->seed.ptr, nl->list.ptr, nullptr, 0);
      } else if (f32) {
        k_pairs_f32<true><<<(nat + 127) / 128, 128, 0, st>>>(nat, g, lo2, hi2, nl->posf.ptr, nl->pos4.ptr,
                                                             nl->sshift.ptr, nl->cell_start.ptr, nl->count.ptr,
                                                             nl->seed.ptr, nl->list.ptr, nl->scal.ptr, nullptr, 0);
      } else {
        k_pairs<true><<<(nat + 127) / 128, 128, 0, st>>>(nat, g, nl->pos4.ptr, nl->sshift.ptr,
                                                         nl->cell_start.ptr, nl->order.ptr,
                                                         nl->count.ptr, nl->seed.ptr, nl->list.ptr,
                                                         nl->scal.ptr, nullptr, 0);
      }
    }
    ATX_LAUNCHED();
    if (!f32) {
      // the exact kernels flag image shifts the 8-bit packing cannot hold (the single-precision path
      // never produces them: wrap counts beyond +-59 cells send the build to the exact kernels)
      ATX_CUDA(cudaMemcpyAsync(&h[3], nl->scal.ptr + 3, sizeof(long long), cudaMemcpyDeviceToHost, st));
      ATX_CUDA(cudaStreamSynchronize(st));
    }
    if (h[3]) {
      atx_set_error("Periodic image shift beyond +-127 cells; wrap the positions into the cell.");
      nl->initialized = false;
      return ATX_ERROR_UNSPECIFIED;
    }
  }
  ATX_CUDA(cudaGetLastError());
  if (nl->verlet_shell > 0.0 && nat > 0) {
    ATX_PASS(nl->pos_build.reserve(nat + 1));
    ATX_CUDA(cudaMemcpyAsync(nl->pos_build.ptr, nl->pos4.ptr, sizeof(double4) * nat, cudaMemcpyDeviceToDevice, st));
  }
  nl->p_rev = p->pos_rev;
  nl->cell_rev = p->cell_rev;
  nl->el_rev = p->el_rev;
  nl->nbuilds++;
  return 0;
}

// caller-supplied list -> internal records (identity order, zero shifts, neighbour element in the entry)
__global__ void k_external_list(int nat, int natloc, const double *__restrict__ r, const int *__restrict__ el,
                                const long long *__restrict__ seed, const int *__restrict__ flat,
                                double4 *__restrict__ pos4, int *__restrict__ order, int *__restrict__ inv,
                                int2 *__restrict__ list, unsigned char *__restrict__ role,
                                long long *__restrict__ scal) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  pos4[s] = make_double4(r[3 * s], r[3 * s + 1], r[3 * s + 2], (double)(el ? el[s] : 1));
  order[s] = s;
  inv[s] = s;
  role[s] = s < natloc ? 2 : 1;
  for (long long a = seed[s]; a < seed[s + 1]; a++) {
    int j = flat[a];
    if (j < 0 || j >= nat) {
      atomicMax((unsigned long long *)&scal[3], 1ull);
      j = s;
    }
    list[a] = make_int2(j, ATX_SHIFT_ZERO | ((el ? el[j] : 1) << 24));
  }
}

extern "C" int atx_neighbors_set_external(atx_neighbors *nl, atx_particles *p, int natloc, int inum,
                                          const int *ilist, const int *numneigh,
                                          const int *const *firstneigh) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  if (!nl || !p || inum < 0 || (inum > 0 && (!ilist || !numneigh || !firstneigh))) return ATX_ERROR_UNSPECIFIED;
  atx_ctx *ctx = nl->ctx;
  cudaStream_t st = ctx->stream;
  const int nat = p->nat;
  if (natloc < 0 || natloc > nat) {
    atx_set_error("atx_neighbors_set_external: natloc out of range.");
    return ATX_ERROR_UNSPECIFIED;
  }
  // flatten the per-atom pointers (pair_atomistica.cpp:426-456 maps them without copying; the
  // device needs one contiguous array)
  std::vector<long long> hseed((size_t)nat + 1, 0);
  for (int ii = 0; ii < inum; ii++) {
    const int i = ilist[ii];
    if (i < 0 || i >= nat) {
      atx_set_error("atx_neighbors_set_external: ilist entry out of range.");
      return ATX_ERROR_UNSPECIFIED;
    }
    hseed[i + 1] = numneigh[i];
  }
  int nebmax = 0;
  for (int i = 0; i < nat; i++) {
    nebmax = std::max(nebmax, (int)hseed[i + 1]);
    hseed[i + 1] += hseed[i];
  }
  const long long npairs = hseed[nat];
  std::vector<int> flat((size_t)npairs + 1);
  for (int ii = 0; ii < inum; ii++) {
    const int i = ilist[ii];
    if (numneigh[i] > 0) memcpy(flat.data() + hseed[i], firstneigh[i], sizeof(int) * numneigh[i]);
  }
  ATX_PASS(nl->seed.reserve((size_t)nat + 2));
  ATX_PASS(nl->ext_flat.reserve((size_t)npairs + 1));
  ATX_PASS(nl->list.reserve((size_t)npairs + 1));
  ATX_PASS(nl->pos4.reserve((size_t)nat + 1));
  ATX_PASS(nl->order.reserve((size_t)nat + 1));
  ATX_PASS(nl->inv.reserve((size_t)nat + 1));
  ATX_PASS(nl->role_ext.reserve((size_t)nat + 1));
  ATX_PASS(nl->scal.reserve(8));
  ATX_CUDA(cudaMemsetAsync(nl->scal.ptr, 0, sizeof(long long) * 8, st));
  ATX_CUDA(cudaMemcpyAsync(nl->seed.ptr, hseed.data(), sizeof(long long) * ((size_t)nat + 1), cudaMemcpyHostToDevice, st));
  if (npairs > 0)
    ATX_CUDA(cudaMemcpyAsync(nl->ext_flat.ptr, flat.data(), sizeof(int) * (size_t)npairs, cudaMemcpyHostToDevice, st));
  if (nat > 0) {
    k_external_list<<<(nat + 127) / 128, 128, 0, st>>>(nat, natloc, p->rptr(), p->el.cap ? p->el.ptr : nullptr,
                                                       nl->seed.ptr, nl->ext_flat.ptr, nl->pos4.ptr, nl->order.ptr,
                                                       nl->inv.ptr, nl->list.ptr, nl->role_ext.ptr, nl->scal.ptr);
    ATX_LAUNCHED();
  }
  long long bad = 0;
  ATX_CUDA(cudaMemcpyAsync(&bad, nl->scal.ptr + 3, sizeof(long long), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));   // also keeps hseed/flat alive until the copies are done
  if (bad) {
    atx_set_error("atx_neighbors_set_external: neighbour index out of range.");
    return ATX_ERROR_UNSPECIFIED;
  }
  nl->external = true;
  nl->natloc = natloc;
  nl->nat = nat;
  nl->npairs = npairs;
  nl->nebmax = nebmax;
  nl->bound = p;
  nl->initialized = true;
  nl->rev_valid = false;
  nl->p_rev = p->pos_rev;
  nl->cell_rev = p->cell_rev;
  nl->el_rev = p->el_rev;
  nl->nbuilds++;
  return 0;
}

extern "C" int atx_neighbors_rebuild(atx_neighbors *nl, atx_particles *p) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  nl->p_rev = -1;
  return atx_neighbors_update(nl, p);
}

int atx_neighbors_refresh_positions(atx_neighbors *nl, atx_particles *p) {
  int nat = nl->nat;
  if (nat > 0) {
    k_refresh_pos<<<(nat + 255) / 256, 256, 0, nl->ctx->stream>>>(nat, p->rptr(), nl->order.ptr,
                                                                 nl->pos4.ptr);
    ATX_LAUNCHED();
  }
  nl->p_rev = p->pos_rev;
  return 0;
}

int atx_neighbors_ensure_rev(atx_neighbors *nl) {
  if (nl->rev_valid) return 0;
  ProfScope ps_(nl->ctx, "nl_reverse_index");
  ATX_PASS(nl->rev.reserve((size_t)nl->npairs + 1));
  if (nl->nat > 0 && nl->npairs > 0) {
    k_reverse_index<<<(nl->nat + 127) / 128, 128, 0, nl->ctx->stream>>>(nl->nat, nl->seed.ptr,
                                                                       nl->list.ptr, nl->rev.ptr);
    ATX_LAUNCHED();
  }
  nl->rev_valid = true;
  return 0;
}

extern "C" int atx_neighbors_get_info(atx_neighbors *nl, long long *npairs, int *nebmax,
                                      int *n_cells, int *stencil) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  if (npairs) *npairs = nl->npairs;
  if (nebmax) *nebmax = nl->nebmax;
  for (int k = 0; k < 3; k++) {
    if (n_cells) n_cells[k] = nl->n_cells[k];
    if (stencil) stencil[k] = nl->sten[k];
  }
  return 0;
}

extern "C" int atx_neighbors_get_interaction_range(atx_neighbors *nl, double *range, double *verlet_shell) {
  if (!nl) return ATX_ERROR_UNSPECIFIED;
  if (range) *range = nl->interaction_range;
  if (verlet_shell) *verlet_shell = nl->verlet_shell;
  return 0;
}

extern "C" int atx_neighbors_get_counters(atx_neighbors *nl, long long *nbuilds, long long *nreused) {
  if (!nl) return ATX_ERROR_UNSPECIFIED;
  if (nbuilds) *nbuilds = nl->nbuilds;
  if (nreused) *nreused = nl->nreused;
  return 0;
}

// 0-based host-layout slot offsets per ORIGINAL atom (includes the terminator slots)
int atx_neighbors_host_seed(atx_neighbors *nl, DevBuf<long long> &hseed) {
  int nat = nl->nat;
  DevBuf<int> hcount;
  ATX_PASS(hcount.reserve(nat + 1));
  ATX_PASS(hseed.reserve(nat + 2));
  ATX_CUDA(cudaMemsetAsync(hcount.ptr, 0, sizeof(int) * (nat + 1), nl->ctx->stream));
  if (nat > 0) {
    k_host_counts<<<(nat + 255) / 256, 256, 0, nl->ctx->stream>>>(nat, nl->inv.ptr, nl->count.ptr,
                                                                  hcount.ptr);
    ATX_LAUNCHED();
  }
  ATX_PASS(atx_scan_int_to_ll(nl->ctx, hcount.ptr, hseed.ptr, nat + 1));
  ATX_CUDA(cudaStreamSynchronize(nl->ctx->stream));  // hcount is freed on return
  return 0;
}

extern "C" int atx_neighbors_copy_to_host(atx_neighbors *nl, intptr_t *seed, intptr_t *last,
                                          int *neighbors, int *dc, long long capacity) {
  if (nl && nl->ctx) cudaSetDevice(nl->ctx->device);
  static_assert(sizeof(intptr_t) == sizeof(long long), "NEIGHPTR_T must be 64 bit");
  if (!nl->initialized) {
    atx_set_error("Neighbor list not built.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_ctx *ctx = nl->ctx;
  cudaStream_t st = ctx->stream;
  int nat = nl->nat;
  long long need = nl->npairs + nat;
  if (need > capacity) {
    atx_set_error("Host neighbor arrays too small: need " + std::to_string(need) + " slots, have " +
                  std::to_string(capacity));
    return ATX_ERROR_UNSPECIFIED;
  }
  DevBuf<int> hcount, d_nb, d_dc;
  DevBuf<long long> hseed, d_seed, d_last;
  ATX_PASS(hcount.reserve(nat + 1));
  ATX_PASS(hseed.reserve(nat + 2));
  ATX_PASS(d_seed.reserve(nat + 2));
  ATX_PASS(d_last.reserve(nat + 2));
  ATX_PASS(d_nb.reserve(need + 1));
  ATX_PASS(d_dc.reserve(3 * (need + 1)));
  ATX_CUDA(cudaMemsetAsync(hcount.ptr, 0, sizeof(int) * (nat + 1), st));
  ATX_CUDA(cudaMemsetAsync(d_dc.ptr, 0, sizeof(int) * 3 * (need + 1), st));
  ATX_CUDA(cudaMemsetAsync(d_last.ptr, 0, sizeof(long long) * (nat + 1), st));
  if (nat > 0) {
    k_host_counts<<<(nat + 255) / 256, 256, 0, st>>>(nat, nl->inv.ptr, nl->count.ptr, hcount.ptr);
    ATX_LAUNCHED();
  }
  ATX_PASS(atx_scan_int_to_ll(ctx, hcount.ptr, hseed.ptr, nat + 1));
  k_host_fill<<<(nat + 256) / 256, 256, 0, st>>>(nat, nl->inv.ptr, nl->order.ptr, nl->seed.ptr,
                                                 nl->list.ptr, hseed.ptr, d_seed.ptr, d_last.ptr,
                                                 d_nb.ptr, d_dc.ptr);
  ATX_LAUNCHED();
  ATX_CUDA(cudaMemcpyAsync(seed, d_seed.ptr, sizeof(long long) * (nat + 1), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaMemcpyAsync(last, d_last.ptr, sizeof(long long) * nat, cudaMemcpyDeviceToHost, st));
  if (need > 0) {
    ATX_CUDA(cudaMemcpyAsync(neighbors, d_nb.ptr, sizeof(int) * need, cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaMemcpyAsync(dc, d_dc.ptr, sizeof(int) * 3 * need, cudaMemcpyDeviceToHost, st));
  }
  ATX_CUDA(cudaStreamSynchronize(st));
  return 0;
}
