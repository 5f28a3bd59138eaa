// List post-processing on the device-resident neighbour list: the helpers of the
// reference's src/python/c/analysis.c (py_pair_distribution :29-106, py_angle_distribution :108-206,
// py_bond_angles :208-318) and f_get_coordination_numbers (neighbors_wrap.f90:271-302) without copying
// the list back to the host.  The reference runs them over the (i, j, r) arrays of get_neighbors; here
// the same loops run over the CSR list in sorted numbering.  The histograms are accumulated as exact
// integers, so the normalised results are the reference's formulas applied to identical counts.
#include <cmath>
#include <vector>

#include "atx_internal.cuh"

// r_i - r_j + Abox.dc, the bond vector f_get_all_neighbors_vec returns (macros.inc:76)
__device__ __forceinline__ void an_bond(const Mat3 &A, const double4 &pi, const double4 *__restrict__ pos4, int2 en,
                                        double &dx, double &dy, double &dz) {
  const double4 pj = pos4[en.x];
  dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
  if (ATX_NONZERO_SHIFT(en.y)) {
    int sx, sy, sz;
    atx_unpack_shift(en.y, sx, sy, sz);
    double ax, ay, az;
    atx_image_vector(A, sx, sy, sz, ax, ay, az);
    dx += ax; dy += ay; dz += az;
  }
}

__global__ void k_an_coordination(int nat, Mat3 A, double cutoff_sq, const double4 *__restrict__ pos4,
                                  const long long *__restrict__ seed, const int2 *__restrict__ list,
                                  const int *__restrict__ order, int *__restrict__ c_orig) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  const double4 pi = pos4[s];
  int c = 0;
  for (long long a = seed[s]; a < seed[s + 1]; a++) {
    double dx, dy, dz;
    an_bond(A, pi, pos4, list[a], dx, dy, dz);
    if (dx * dx + dy * dy + dz * dz < cutoff_sq) c++;
  }
  c_orig[order[s]] = c;
}

// One warp per atom: per-atom histogram in shared memory, then sum and sum of squares over the atoms.
// MODE 0: pair distances, bin = int(nbins*r/cutoff); MODE 1: angles between all ordered pairs of bonds
// shorter than the cutoff, bin = int(nbins*angle/pi) wrapped into [0, nbins).
// acc[0..nbins) = sum of counts, acc[nbins..2 nbins) = sum of squared counts, acc[2 nbins] = atoms with
// at least one list entry, acc[2 nbins + 1] = number of angles.
template <int MODE>
__global__ void __launch_bounds__(256)
k_an_histogram(int nat, Mat3 A, int nbins, double cutoff, const double4 *__restrict__ pos4,
               const long long *__restrict__ seed, const int2 *__restrict__ list,
               unsigned long long *__restrict__ acc) {
  extern __shared__ int sh[];   // 8 per-warp histograms, then the block's two accumulators (as 2 ints per entry)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int *hist = sh + warp * nbins;
  unsigned long long *bh = reinterpret_cast<unsigned long long *>(sh + 8 * nbins + (nbins & 1));
  unsigned long long *bh2 = bh + nbins;
  for (int b = threadIdx.x; b < 2 * nbins; b += blockDim.x) bh[b] = 0ull;
  for (int b = lane; b < nbins; b += 32) hist[b] = 0;
  __syncthreads();
  unsigned long long natoms = 0ull, nangles = 0ull;
  const double cutoff_sq = cutoff * cutoff;
  for (int s = blockIdx.x * 8 + warp; s < nat; s += gridDim.x * 8) {
    const long long b0 = seed[s], b1 = seed[s + 1];
    if (b1 == b0) continue;
    if (lane == 0) natoms++;
    const double4 pi = pos4[s];
    for (long long a = b0 + lane; a < b1; a += 32) {
      double dx, dy, dz;
      an_bond(A, pi, pos4, list[a], dx, dy, dz);
      const double n = dx * dx + dy * dy + dz * dz;
      if (MODE == 0) {
        const int bin = (int)(nbins * sqrt(n) / cutoff);
        if (bin >= 0 && bin < nbins) atomicAdd(&hist[bin], 1);
      } else if (n < cutoff_sq) {
        for (long long a2 = b0; a2 < b1; a2++) {
          if (a2 == a) continue;
          double ex, ey, ez;
          an_bond(A, pi, pos4, list[a2], ex, ey, ez);
          const double n2 = ex * ex + ey * ey + ez * ez;
          if (n2 < cutoff_sq) {
            const double angle = acos((dx * ex + dy * ey + dz * ez) / sqrt(n * n2));
            int bin = (int)(nbins * angle / 3.14159265358979323846);
            while (bin < 0) bin += nbins;
            while (bin >= nbins) bin -= nbins;
            atomicAdd(&hist[bin], 1);
            nangles++;
          }
        }
      }
    }
    __syncwarp();
    for (int b = lane; b < nbins; b += 32) {
      const unsigned long long c = (unsigned long long)hist[b];
      if (c) {
        atomicAdd(&bh[b], c);
        atomicAdd(&bh2[b], c * c);
        hist[b] = 0;
      }
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) nangles += __shfl_xor_sync(0xffffffffu, nangles, o);
  if (lane == 0) {
    if (natoms) atomicAdd(&acc[2 * nbins], natoms);
    if (nangles) atomicAdd(&acc[2 * nbins + 1], nangles);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 2 * nbins; b += blockDim.x)
    if (bh[b]) atomicAdd(&acc[b], bh[b]);
}

// per-atom moment of the bond-angle distribution (analysis.c:208-318)
__global__ void k_an_bond_angles(int nat, Mat3 A, int moment, double cutoff_sq, const double4 *__restrict__ pos4,
                                 const long long *__restrict__ seed, const int2 *__restrict__ list,
                                 const int *__restrict__ order, double *__restrict__ m_orig) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  const double4 pi = pos4[s];
  const long long b0 = seed[s], b1 = seed[s + 1];
  double accum = 0.0;
  long long nangle = 0;
  for (long long a = b0; a < b1; a++) {
    double dx, dy, dz;
    an_bond(A, pi, pos4, list[a], dx, dy, dz);
    const double n = dx * dx + dy * dy + dz * dz;
    if (!(n < cutoff_sq)) continue;
    for (long long a2 = b0; a2 < b1; a2++) {
      if (a2 == a) continue;
      double ex, ey, ez;
      an_bond(A, pi, pos4, list[a2], ex, ey, ez);
      const double n2 = ex * ex + ey * ey + ez * ez;
      if (n2 < cutoff_sq) {
        const double angle = acos((dx * ex + dy * ey + dz * ez) / sqrt(n * n2));
        accum += pow(angle, (double)moment);
        nangle++;
      }
    }
  }
  m_orig[order[s]] = nangle > 0 ? accum / (double)nangle : 0.0;
}

static int an_ready(atx_neighbors *nl, atx_particles *p) {
  if (!nl || !p) return ATX_ERROR_UNSPECIFIED;
  cudaSetDevice(nl->ctx->device);
  ATX_PASS(atx_neighbors_update(nl, p));
  return 0;
}

extern "C" int atx_neighbors_coordination_numbers(atx_neighbors *nl, atx_particles *p, double cutoff, int *c) {
  ATX_PASS(an_ready(nl, p));
  const int nat = nl->nat;
  if (nat == 0) return 0;
  cudaStream_t st = nl->ctx->stream;
  DevBuf<int> d;
  ATX_PASS(d.reserve(nat + 1));
  k_an_coordination<<<(nat + 127) / 128, 128, 0, st>>>(nat, p->Abox, cutoff * cutoff, nl->pos4.ptr, nl->seed.ptr,
                                                       nl->list.ptr, nl->order.ptr, d.ptr);
  ATX_LAUNCHED();
  ATX_CUDA(cudaMemcpyAsync(c, d.ptr, sizeof(int) * nat, cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

static int an_histogram(atx_neighbors *nl, atx_particles *p, int mode, int nbins, double cutoff, double *h,
                        double *h2) {
  ATX_PASS(an_ready(nl, p));
  if (nbins < 1 || nbins > 4096 || !(cutoff > 0.0)) {
    atx_set_error("analysis: nbins must be 1 .. 4096 and cutoff positive.");
    return ATX_ERROR_UNSPECIFIED;
  }
  cudaStream_t st = nl->ctx->stream;
  DevBuf<unsigned long long> acc;
  ATX_PASS(acc.reserve(2 * (size_t)nbins + 2));
  ATX_CUDA(cudaMemsetAsync(acc.ptr, 0, sizeof(unsigned long long) * (2 * (size_t)nbins + 2), st));
  const size_t smem = sizeof(int) * (8 * (size_t)nbins + 2) + sizeof(unsigned long long) * 2 * (size_t)nbins;
  const int nat = nl->nat;
  int blocks = (nat + 7) / 8;
  if (blocks > nl->ctx->sm_count * 8) blocks = nl->ctx->sm_count * 8;
  if (blocks < 1) blocks = 1;
  if (mode == 0) {
    ATX_CUDA(cudaFuncSetAttribute(k_an_histogram<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_an_histogram<0><<<blocks, 256, smem, st>>>(nat, p->Abox, nbins, cutoff, nl->pos4.ptr, nl->seed.ptr,
                                                  nl->list.ptr, acc.ptr);
  } else {
    ATX_CUDA(cudaFuncSetAttribute(k_an_histogram<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_an_histogram<1><<<blocks, 256, smem, st>>>(nat, p->Abox, nbins, cutoff, nl->pos4.ptr, nl->seed.ptr,
                                                  nl->list.ptr, acc.ptr);
  }
  ATX_LAUNCHED();
  std::vector<unsigned long long> host(2 * (size_t)nbins + 2);
  ATX_CUDA(cudaMemcpyAsync(host.data(), acc.ptr, sizeof(unsigned long long) * host.size(), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ATX_CUDA(cudaGetLastError());
  // normalisation of analysis.c:88-99 (pairs: per atom and shell volume) / :190-200 (angles: nangle starts at 1)
  const double natoms = host[2 * nbins] > 0 ? (double)host[2 * nbins] : 1.0;
  const double nangle = 1.0 + (double)host[2 * nbins + 1];
  for (int b = 0; b < nbins; b++) {
    double norm, binvol;
    if (mode == 0) {
      const double r1 = b * cutoff / nbins, r2 = (b + 1) * cutoff / nbins;
      binvol = 4 * 3.14159265358979323846 / 3 * (r2 * r2 * r2 - r1 * r1 * r1);
      norm = natoms;
    } else {
      binvol = 3.14159265358979323846 / nbins;
      norm = nangle;
    }
    h[b] = (double)host[b] / (norm * binvol);
    h2[b] = (double)host[nbins + b] / (norm * binvol * binvol) - h[b] * h[b];
  }
  return 0;
}

extern "C" int atx_neighbors_pair_distribution(atx_neighbors *nl, atx_particles *p, int nbins, double cutoff,
                                               double *h, double *h2) {
  return an_histogram(nl, p, 0, nbins, cutoff, h, h2);
}

extern "C" int atx_neighbors_angle_distribution(atx_neighbors *nl, atx_particles *p, int nbins, double cutoff,
                                                double *h, double *h2) {
  return an_histogram(nl, p, 1, nbins, cutoff, h, h2);
}

extern "C" int atx_neighbors_bond_angles(atx_neighbors *nl, atx_particles *p, int moment, double cutoff, double *m) {
  ATX_PASS(an_ready(nl, p));
  const int nat = nl->nat;
  if (nat == 0) return 0;
  cudaStream_t st = nl->ctx->stream;
  DevBuf<double> d;
  ATX_PASS(d.reserve(nat + 1));
  k_an_bond_angles<<<(nat + 127) / 128, 128, 0, st>>>(nat, p->Abox, moment, cutoff * cutoff, nl->pos4.ptr,
                                                      nl->seed.ptr, nl->list.ptr, nl->order.ptr, d.ptr);
  ATX_LAUNCHED();
  ATX_CUDA(cudaMemcpyAsync(m, d.ptr, sizeof(double) * nat, cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  return 0;
}
