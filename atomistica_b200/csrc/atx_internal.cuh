// Internal declarations shared by the CUDA translation units of libatomistica_b200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/atomistica_b200.h"

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------

void atx_set_error(const std::string &msg);
extern long long g_atx_launches;

#define ATX_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      atx_set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                    ":" + std::to_string(__LINE__) + ")");                               \
      return ATX_ERROR_DEVICE;                                                           \
    }                                                                                    \
  } while (0)

#define ATX_PASS(call)         \
  do {                         \
    int r_ = (call);           \
    if (r_ != 0) return r_;    \
  } while (0)

#define ATX_LAUNCHED() (++g_atx_launches)

// ---------------------------------------------------------------------------
// grow-only device buffer
// ---------------------------------------------------------------------------

template <typename T>
struct DevBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc((void **)&ptr, want * sizeof(T));
    if (e != cudaSuccess) {
      cap = 0;
      atx_set_error(std::string("cudaMalloc of ") + std::to_string(want * sizeof(T)) +
                    " bytes failed: " + cudaGetErrorString(e));
      return ATX_ERROR_DEVICE;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
};

// pinned host staging buffer
template <typename T>
struct PinBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMallocHost((void **)&ptr, want * sizeof(T));
    if (e != cudaSuccess) {
      cap = 0;
      atx_set_error(std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
      return ATX_ERROR_DEVICE;
    }
    cap = want;
    return 0;
  }
  ~PinBuf() {
    if (ptr) cudaFreeHost(ptr);
  }
};

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------

struct ProfSlot {
  std::string name;
  std::vector<cudaEvent_t> ev;  // pairs: start, stop
  size_t used = 0;
};

struct atx_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  DevBuf<char> cub_tmp;
  bool prof_on = false;
  std::vector<ProfSlot> prof;
};

// RAII event pair around a kernel launch; free when profiling is off
struct ProfScope {
  atx_ctx *ctx;
  cudaEvent_t stop = nullptr;
  ProfScope(atx_ctx *c, const char *name);
  ~ProfScope();
};

// 3x3 column-major matrix passed by value to kernels
struct Mat3 {
  double m[9];
};

struct atx_particles {
  atx_ctx *ctx = nullptr;
  int nat = 0;
  Mat3 Abox{}, Bbox{};
  int pbc[3] = {1, 1, 1};
  DevBuf<double> r;   // (3,nat) original atom order
  const double *r_ext = nullptr;  // externally owned device positions (set_positions_device)
  DevBuf<int> el;     // particle element ids, original order
  PinBuf<double> stage;
  long long pos_rev = 0, cell_rev = 0, el_rev = 0;
  const double *rptr() const { return r_ext ? r_ext : r.ptr; }
};

// list entry word .y: periodic shift, 8 bits per component (bias 128), in bits 0..23 and the
// particle element id of the neighbour in bits 24..30
#define ATX_SHIFT_BIAS 128
#define ATX_SHIFT_ZERO (ATX_SHIFT_BIAS | (ATX_SHIFT_BIAS << 8) | (ATX_SHIFT_BIAS << 16))
#define ATX_SHIFT_MASK 0xFFFFFF
#define ATX_NONZERO_SHIFT(p) (((p) & ATX_SHIFT_MASK) != ATX_SHIFT_ZERO)
#define ATX_ENTRY_EL(p) ((p) >> 24)

struct atx_neighbors {
  atx_ctx *ctx = nullptr;
  int avgn = 100;
  double interaction_range = 0.0;
  double pair_range[32][32] = {{0.0}};   // largest request per pair of particle element ids (1-based), LAMMPS hosts ask for it
  double verlet_shell = 0.0;
  double cutoff = 0.0;
  bool initialized = false;
  long long capacity = 0;  // nat*avgn, fixed at first build (python_neighbors.f90:499)
  long long p_rev = -1, cell_rev = -1;
  const atx_particles *bound = nullptr;

  // geometry (binning_init)
  int n_cells[3] = {0, 0, 0};
  int sten[3] = {1, 1, 1};
  Mat3 rec_cell_size{};

  // per build
  int nat = 0;
  long long npairs = 0;
  int nebmax = 0;
  long long nbuilds = 0;
  long long nreused = 0;    // updates answered by the Verlet-shell check without a rebuild
  long long el_rev = -1;

  DevBuf<int4> cellshift;   // per original atom: cell id, wrap shift
  DevBuf<int> cell_count;   // ncell+1
  DevBuf<int> cell_start;   // ncell+1
  DevBuf<int> cell_fill;    // ncell
  DevBuf<int> order;        // sorted slot -> original atom
  DevBuf<int> inv;          // original atom -> sorted slot
  DevBuf<double4> pos4;     // sorted: x,y,z, (w = element id as double bits)
  DevBuf<double4> pos_build;  // pos4 at the last build (library-mode Verlet shell only)
  DevBuf<int4> sshift;      // sorted: cell id + wrap shift
  DevBuf<float4> posf;      // sorted: position relative to the atom's cell origin (float), w = bits of
                            // (element << 24 | packed wrap shift): 16-byte candidate record of the
                            // single-precision pre-filter of the pair search
  double f32_delta = -1.0;  // > 0: half-width of the band around cutoff^2 inside which the exact test runs
  DevBuf<int> count;        // nat+1 neighbour counts (sorted order)
  DevBuf<long long> seed;   // nat+1 exclusive offsets (sorted order), 0-based
  DevBuf<int2> list;        // device list: {sorted j, packed shift}
  DevBuf<int2> rows;        // fixed-width per-atom rows of the single-pass build (scratch)
  DevBuf<int> rev;          // optional reverse-slot index
  bool rev_valid = false;
  // caller-supplied list (LAMMPS host, atx_neighbors_set_external): atoms >= natloc are ghosts that
  // appear as separate atoms, no periodic shifts; role_ext = 2 for owned atoms, 1 for ghosts
  bool external = false;
  int natloc = -1;
  DevBuf<unsigned char> role_ext;
  DevBuf<int> ext_flat;
  DevBuf<long long> scal;   // small scalar scratch (npairs, nebmax, flags)
};

// neighbour-list internals used by the potentials
int atx_neighbors_ensure_rev(atx_neighbors *nl);
// refresh pos4 from p->r without rebuilding the list (positions moved less than the skin)
int atx_neighbors_refresh_positions(atx_neighbors *nl, atx_particles *p);

// device exclusive scan helpers (cub)
int atx_scan_int_to_ll(atx_ctx *ctx, const int *in, long long *out, size_t n);
int atx_scan_int(atx_ctx *ctx, const int *in, int *out, size_t n);

// host helper: accumulate device array (sorted or original order) into a host array
int atx_accumulate_to_host(atx_ctx *ctx, const double *dev, double *host, size_t n,
                           PinBuf<double> &stage);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------

__device__ __forceinline__ int atx_pack_shift(int sx, int sy, int sz) {
  return (sx + ATX_SHIFT_BIAS) | ((sy + ATX_SHIFT_BIAS) << 8) | ((sz + ATX_SHIFT_BIAS) << 16);
}
__device__ __forceinline__ void atx_unpack_shift(int p, int &sx, int &sy, int &sz) {
  sx = (p & 255) - ATX_SHIFT_BIAS;
  sy = ((p >> 8) & 255) - ATX_SHIFT_BIAS;
  sz = ((p >> 16) & 255) - ATX_SHIFT_BIAS;
}

// matmul(Abox, shift) with the reference's association order and no FMA contraction
__device__ __forceinline__ void atx_image_vector(const Mat3 &A, int sx, int sy, int sz, double &ax,
                                                 double &ay, double &az) {
  double s0 = (double)sx, s1 = (double)sy, s2 = (double)sz;
  ax = __dadd_rn(__dadd_rn(__dmul_rn(A.m[0], s0), __dmul_rn(A.m[3], s1)), __dmul_rn(A.m[6], s2));
  ay = __dadd_rn(__dadd_rn(__dmul_rn(A.m[1], s0), __dmul_rn(A.m[4], s1)), __dmul_rn(A.m[7], s2));
  az = __dadd_rn(__dadd_rn(__dmul_rn(A.m[2], s0), __dmul_rn(A.m[5], s1)), __dmul_rn(A.m[8], s2));
}

// read-only 32-byte load as ONE 256-bit instruction (LDG.E.ENL2.256.CONSTANT on sm_100a):
// one L1 request per gathered record instead of two
__device__ __forceinline__ double4 atx_ld4(const double4 *p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
}
// coherent variant for data written earlier in the same kernel sequence by other kernels is not
// needed (kernel boundaries order the writes); this one is for arrays updated in place.
__device__ __forceinline__ double4 atx_ld4_ca(const double4 *p) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
}

template <typename T>
__device__ __forceinline__ T atx_warp_sum(T v, int width = 32) {
  for (int o = width / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
  return v;
}
