// Spatial domain decomposition over the GPUs of one node: slab decomposition, ghost-atom halo
// exchange and global rebuild decision over NCCL (NVLink 5 / NVSwitch).
//
// Behavioural model: src/standalone/domain_decomposition.f90 (communicate_particles :494-640,
// communicate_ghosts :699-826, global OR of the rebuild flag standalone/neighbors.f90:552-557).
// Differences by design:
//   * 1-D slabs along the first cell vector: on an NVSwitch box every peer is one hop away and the
//     halo messages are sub-MB, so two large messages per step beat 26 small ones.
//   * The halo is 2*(rc + skin) thick and every rank evaluates the density / bond-order terms of
//     its inner ghosts itself; the reference's reverse force communication (communicate_forces
//     :887-970) and the EAM embedding-derivative exchange disappear, one forward position exchange
//     per step remains.
//   * The rebuild flag is all-reduced on the device into the stop flag of the optimistic step
//     batches (atx_md.cu), so there is no per-step host synchronisation either.
// NCCL is loaded with dlopen("libnccl.so.2"): the library has no link-time dependency on it and
// single-GPU users never touch it.
#include <cub/device/device_select.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "atx_potential_common.cuh"

#define ATX_ACCEL_CONV 9.648533212331e-3

int atx_eam_compute_device(atx_eam *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o);
int atx_bop_compute_device(atx_bop *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o);
int atx_rebo2_compute_device(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, const PotOut &o);
int atx_bop_check_overflow(atx_bop *pot);
bool atx_bop_supports_split(atx_bop *pot);
int atx_rebo2_check_overflow(atx_rebo2 *pot);

// ---------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.h) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    atx_set_error(std::string("Cannot load libnccl.so.2: ") + dlerror());
    return ATX_ERROR_MPI;
  }
#define LOAD(field, name)                                            \
  *(void **)(&g_nccl.field) = dlsym(h, name);                        \
  if (!g_nccl.field) {                                               \
    atx_set_error(std::string("libnccl lacks symbol ") + name);     \
    return ATX_ERROR_MPI;                                            \
  }
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(AllGather, "ncclAllGather")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  g_nccl.h = h;
  return 0;
}

#define ATX_NCCL(call)                                                                      \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess) {                                                                \
      atx_set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r_));                 \
      return ATX_ERROR_MPI;                                                                 \
    }                                                                                       \
  } while (0)

struct atx_dd {
  atx_ctx *ctx = nullptr;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
};

extern "C" int atx_dd_get_unique_id(char *id128) {
  ATX_PASS(nccl_load());
  ncclUniqueId id;
  ATX_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
}

extern "C" int atx_dd_create(atx_ctx *ctx, int rank, int nranks, const char *id128, atx_dd **out) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  ATX_PASS(nccl_load());
  ATX_CUDA(cudaSetDevice(ctx->device));
  atx_dd *dd = new atx_dd();
  dd->ctx = ctx;
  dd->rank = rank;
  dd->nranks = nranks;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ATX_NCCL(g_nccl.CommInitRank(&dd->comm, nranks, id, rank));
  *out = dd;
  return 0;
}

extern "C" int atx_dd_destroy(atx_dd *dd) {
  if (dd && dd->ctx) cudaSetDevice(dd->ctx->device);
  if (!dd) return 0;
  if (dd->comm) g_nccl.CommDestroy(dd->comm);
  delete dd;
  return 0;
}

// ---------------------------------------------------------------------------
// domain-decomposed MD
// ---------------------------------------------------------------------------

struct DdCtrl {
  int stop;          // global (all-reduced) rebuild flag
  int want;          // local rebuild wish of the current step
  int steps_done;
  unsigned int counter_drift;
  unsigned int counter_kick;
  unsigned long long stepmax_bits;
  double accum_max_dr;
  double verlet_shell;
  double epot;       // local: sum of per-atom energies of owned atoms (last step)
  double ekin;
  unsigned long long seq;     // executed guarded steps since create (stamps the peer-to-peer signals)
  unsigned int counter_pack;
  int err;                    // 1: a peer's signal did not arrive in time
};

// Peer-to-peer halo exchange (the step path when every rank could map its peers' mailboxes):
// every rank owns one cudaMalloc'ed mailbox, exported with cudaIpcGetMemHandle and mapped by all
// peers.  Layout: sig[2][DD_MAXP] signal words, then the ghost receive areas recvL[2][3*capG] and
// recvR[2][3*capG] (positions arriving from the left / right slab; two parities, see below).
// Per step the pack kernel STORES the ghost positions straight into the neighbours' receive areas
// over NVLink, fences, and then stores (seq << 1 | rebuild wish) into the sig slot [seq & 1][me] of
// EVERY rank: one message carries the halo-arrival flag and this rank's share of the global OR of
// the rebuild rule (standalone/neighbors.f90:552-557), replacing ncclSend/ncclRecv + ncclAllReduce.
// A rank can run at most one step ahead of a peer (it cannot pass wait(s) without the peer's
// signal of step s), so two parities of every slot and receive area are enough.
#define DD_MAXP 16
#define DD_SIG_BYTES (2 * DD_MAXP * sizeof(unsigned long long))

#define DD_ROW 9  // migration record: id, el, r(3), v(3), minv

struct atx_ddmd {
  atx_dd *dd = nullptr;
  atx_ctx *ctx = nullptr;
  int pot_kind = 0;
  void *pot = nullptr;
  Mat3 A{}, B{};
  int pbc[3] = {1, 1, 1};
  double slo = 0, shi = 1, hfrac = 0;
  double slo_own = 0, shi_own = 1;  // ownership interval (open ended at non-periodic faces)
  double a1[3] = {0, 0, 0};
  double torig[3] = {0, 0, 0};
  int left = -1, right = -1;
  double wrapL = 0.0, wrapR = 0.0;  // multiples of a1 added to positions sent left / right
  double dt = 1.0, rc = 0.0, skin = 0.0;
  int nown = 0, ngl = 0, ngr = 0, nsendL = 0, nsendR = 0;
  DevBuf<double> r, v, f, minv, tmpd, bufL, bufR, epa, sums, kin_partials, mig_send, mig_recv;
  DevBuf<double> idd;  // atom ids stored as doubles (exact up to 2^53)
  DevBuf<int> el, sendL, sendR, flags, flags2, flags3, sel, tmpi, cnt;
  DevBuf<double> r2, v2, f2, minv2, idd2;  // double buffers for the migration compaction
  DevBuf<int> el2;
  DevBuf<unsigned char> role, role_loc;      // 2 owned, 1 inner ghost, 0 outer ghost (sorted / local order)
  DevBuf<DdCtrl> ctrl;
  PinBuf<DdCtrl> hctrl;
  PinBuf<double> stage;
  PinBuf<int> hcnt;
  atx_particles ploc;
  atx_neighbors *nl = nullptr;
  long long nrebuilds = 0;
  double last_ms = 0.0;
  int batch = 16;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // peer-to-peer step path
  bool p2p = false;
  size_t capG = 0;                       // ghost capacity per side (atoms), identical on all ranks
  char *mbox = nullptr;                  // own mailbox
  char *peer_mbox[DD_MAXP] = {nullptr};  // mapped mailboxes ([rank] = own)
  DevBuf<unsigned long long *> d_peer_sig;
  double rebuild_host_ms[8] = {0};       // host wall time per rebuild phase (accumulated)
  // sorted range of the interior atoms (owned, no ghost in their list): evaluated while the halo is in flight
  int split_lo = 0, split_hi = 0;
  bool split = false;
};

// inv / pos4 (peer-to-peer step path): the new position also goes straight into the atom's sorted
// record, so that the refresh pass only has to place the ghosts
__global__ void k_dd_drift(int nown, double dt, double *__restrict__ r, double *__restrict__ v,
                           const double *__restrict__ f, const double *__restrict__ minv,
                           DdCtrl *__restrict__ ctrl, const int *__restrict__ inv = nullptr,
                           double4 *__restrict__ pos4 = nullptr) {
  if (ctrl->stop) return;
  __shared__ double red[8];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (i < nown) {
    double a = 0.5 * minv[i] * ATX_ACCEL_CONV * dt;
    double vx = v[3 * i] + a * f[3 * i], vy = v[3 * i + 1] + a * f[3 * i + 1], vz = v[3 * i + 2] + a * f[3 * i + 2];
    v[3 * i] = vx; v[3 * i + 1] = vy; v[3 * i + 2] = vz;
    double dx = vx * dt, dy = vy * dt, dz = vz * dt;
    const double x = r[3 * i] + dx, y = r[3 * i + 1] + dy, z = r[3 * i + 2] + dz;
    r[3 * i] = x; r[3 * i + 1] = y; r[3 * i + 2] = z;
    if (pos4) {
      double *q = reinterpret_cast<double *>(pos4 + inv[i]);
      q[0] = x; q[1] = y; q[2] = z;
    }
    d2 = dx * dx + dy * dy + dz * dz;
  }
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, red[w]);
    atomicMax(&ctrl->stepmax_bits, (unsigned long long)__double_as_longlong(m));
    __threadfence();
    unsigned int done = atomicAdd(&ctrl->counter_drift, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double mx = __longlong_as_double((long long)atomicAdd(&ctrl->stepmax_bits, 0ull));
      double acc = ctrl->accum_max_dr + sqrt(mx);
      ctrl->accum_max_dr = acc;
      ctrl->stepmax_bits = 0ull;
      ctrl->counter_drift = 0u;
      ctrl->want = (2.0 * acc >= ctrl->verlet_shell) ? 1 : 0;  // all-reduced into ctrl->stop
      ctrl->seq += 1ull;
    }
  }
}

// Velocity-Verlet in leapfrog form for the steps inside a run: the second half kick of step n and the
// first half kick of step n + 1 use the same force, so one kernel applies both (scale = 1; 0.5 for the
// first step of a run, whose state has whole-step velocities) and drifts.  The force is read from the
// potential's sorted output, the position goes to the local array (halo packing, rebuilds) and to
// the sorted record.  The whole-step velocity, the local force array and the energy sums are
// produced once, by k_dd_kick at the end of the run.
__global__ void k_dd_kickdrift(int nown, double dt, double scale, double *__restrict__ r, double *__restrict__ v,
                               const double *__restrict__ fs, const double *__restrict__ minv,
                               DdCtrl *__restrict__ ctrl, const int *__restrict__ inv,
                               double4 *__restrict__ pos4) {
  if (ctrl->stop) return;
  __shared__ double red[8];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (i < nown) {
    const int s = inv[i];
    double a = scale * minv[i] * ATX_ACCEL_CONV * dt;
    double vx = v[3 * i] + a * fs[3 * s], vy = v[3 * i + 1] + a * fs[3 * s + 1], vz = v[3 * i + 2] + a * fs[3 * s + 2];
    v[3 * i] = vx; v[3 * i + 1] = vy; v[3 * i + 2] = vz;
    double dx = vx * dt, dy = vy * dt, dz = vz * dt;
    const double x = r[3 * i] + dx, y = r[3 * i + 1] + dy, z = r[3 * i + 2] + dz;
    r[3 * i] = x; r[3 * i + 1] = y; r[3 * i + 2] = z;
    double *q = reinterpret_cast<double *>(pos4 + s);
    q[0] = x; q[1] = y; q[2] = z;
    d2 = dx * dx + dy * dy + dz * dz;
  }
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, red[w]);
    atomicMax(&ctrl->stepmax_bits, (unsigned long long)__double_as_longlong(m));
    __threadfence();
    unsigned int done = atomicAdd(&ctrl->counter_drift, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double mx = __longlong_as_double((long long)atomicAdd(&ctrl->stepmax_bits, 0ull));
      double acc = ctrl->accum_max_dr + sqrt(mx);
      ctrl->accum_max_dr = acc;
      ctrl->stepmax_bits = 0ull;
      ctrl->counter_drift = 0u;
      ctrl->want = (2.0 * acc >= ctrl->verlet_shell) ? 1 : 0;
      ctrl->seq += 1ull;
    }
  }
}

// ranks with no owned atoms still have to publish a wish
__global__ void k_dd_nowish(DdCtrl *ctrl) {
  if (!ctrl->stop) {
    ctrl->want = 0;
    ctrl->seq += 1ull;
  }
}

// ---- peer-to-peer step path -------------------------------------------------------------------

// Ghost positions of this step straight into the neighbours' receive areas (remote stores over
// NVLink), then the step's signal word into every rank's mailbox.  dstL / dstR: parity-0 base of the
// LEFT neighbour's recvR area / the RIGHT neighbour's recvL area as mapped into this process.
__global__ void __launch_bounds__(256)
k_dd_pack_p2p(int nL, int nR, const int *__restrict__ idxL, const int *__restrict__ idxR,
              const double *__restrict__ r, double lx, double ly, double lz, double rx, double ry, double rz,
              double *dstL, double *dstR, size_t par_stride, unsigned long long *const *peer_sig, int me,
              int P, DdCtrl *ctrl) {
  if (ctrl->stop) return;
  __shared__ bool is_last;
  const unsigned long long seq = ctrl->seq;
  const size_t par = (size_t)(seq & 1ull) * par_stride;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nL) {
    const int i = idxL[t];
    double *o = dstL + par + 3 * (size_t)t;
    o[0] = r[3 * i] + lx; o[1] = r[3 * i + 1] + ly; o[2] = r[3 * i + 2] + lz;
  } else if (t < nL + nR) {
    const int u = t - nL, i = idxR[u];
    double *o = dstR + par + 3 * (size_t)u;
    o[0] = r[3 * i] + rx; o[1] = r[3 * i + 1] + ry; o[2] = r[3 * i + 2] + rz;
  }
  __threadfence_system();   // this thread's remote stores before anything that follows
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int done = atomicAdd(&ctrl->counter_pack, 1u);
    is_last = (done == gridDim.x - 1);
    __threadfence_system();
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < P) {
      const unsigned long long word = (seq << 1) | (ctrl->want ? 1ull : 0ull);
      __threadfence_system();
      *((volatile unsigned long long *)(peer_sig[threadIdx.x] + (seq & 1ull) * DD_MAXP + me)) = word;
    }
    if (threadIdx.x == 0) ctrl->counter_pack = 0u;
  }
}

// One warp: lane p waits for rank p's signal of this step; the OR of the wishes becomes the stop flag
// (identical on every rank).  A signal that does not arrive within ~4 s (a peer failed) raises
// ctrl->err and stops the batch instead of hanging the device.
__global__ void k_dd_wait(const unsigned long long *sig, int P, DdCtrl *ctrl, int count_step) {
  if (ctrl->stop) return;
  const unsigned long long seq = ctrl->seq;
  const int lane = threadIdx.x;
  int want = 0, bad = 0;
  if (lane < P) {
    const volatile unsigned long long *s = sig + (seq & 1ull) * DD_MAXP + lane;
    const long long t0 = clock64();
    unsigned long long v = *s;
    while ((v >> 1) != seq) {
      if (clock64() - t0 > 8000000000ll) { bad = 1; break; }
      __nanosleep(40);
      v = *s;
    }
    want = (int)(v & 1ull);
  }
  __threadfence_system();
  want = __any_sync(0xffffffffu, want);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    if (bad) { ctrl->err = 1; ctrl->stop = 1; }
    else ctrl->stop = want;
    // leapfrog form: nothing can stop this step any more, so it is counted here (no kick kernel follows)
    if (count_step && !bad && !want) ctrl->steps_done += 1;
  }
}

// sorted position records of the ghosts from the receive areas of this step (the owned atoms were
// placed by the drift); one thread per ghost
__global__ void k_dd_refresh_p2p(int ng, int n, int ngl, const double *recvL, const double *recvR,
                                 size_t par_stride, const int *__restrict__ inv, double4 *__restrict__ pos4,
                                 const DdCtrl *ctrl) {
  if (ctrl->stop) return;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const size_t par = (size_t)(ctrl->seq & 1ull) * par_stride;
  const double *src = g < ngl ? recvL + par + 3 * (size_t)g : recvR + par + 3 * (size_t)(g - ngl);
  double *q = reinterpret_cast<double *>(pos4 + inv[n + g]);
  q[0] = src[0]; q[1] = src[1]; q[2] = src[2];
}

__global__ void k_dd_pack(int n, const int *__restrict__ idx, const double *__restrict__ r,
                          double sx, double sy, double sz, double *__restrict__ buf,
                          const int *__restrict__ stop) {
  if (stop && *stop) return;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = idx[t];
  buf[3 * t] = r[3 * i] + sx;
  buf[3 * t + 1] = r[3 * i + 1] + sy;
  buf[3 * t + 2] = r[3 * i + 2] + sz;
}

__global__ void k_dd_refresh(int n, const double *__restrict__ r, const int *__restrict__ order,
                             double4 *__restrict__ pos4, const int *__restrict__ stop) {
  if (stop && *stop) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = order[s];
  double4 v = pos4[s];
  v.x = r[3 * i]; v.y = r[3 * i + 1]; v.z = r[3 * i + 2];
  pos4[s] = v;
}

// kick owned atoms with the forces of the sorted arrays; kinetic energy and owned potential energy
__global__ void k_dd_kick(int nown, double dt, double *__restrict__ v, const double *__restrict__ fs,
                          const double *__restrict__ epa_s, const int *__restrict__ inv,
                          const double *__restrict__ minv, double *__restrict__ f_loc,
                          double *__restrict__ partials, DdCtrl *__restrict__ ctrl, int count_step) {
  if (ctrl->stop) return;
  __shared__ double red[16];
  __shared__ bool is_last;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ek = 0.0, ep = 0.0;
  if (i < nown) {
    int s = inv[i];
    double mi = minv[i];
    double a = 0.5 * mi * ATX_ACCEL_CONV * dt;
    double fx = fs[3 * s], fy = fs[3 * s + 1], fz = fs[3 * s + 2];
    f_loc[3 * i] = fx; f_loc[3 * i + 1] = fy; f_loc[3 * i + 2] = fz;
    double vx = v[3 * i] + a * fx, vy = v[3 * i + 1] + a * fy, vz = v[3 * i + 2] + a * fz;
    v[3 * i] = vx; v[3 * i + 1] = vy; v[3 * i + 2] = vz;
    ek = 0.5 * (vx * vx + vy * vy + vz * vz) / (mi * ATX_ACCEL_CONV);
    ep = epa_s[s];
  }
  for (int o = 16; o > 0; o >>= 1) {
    ek += __shfl_xor_sync(0xffffffffu, ek, o);
    ep += __shfl_xor_sync(0xffffffffu, ep, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = ek; red[8 + (threadIdx.x >> 5)] = ep; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, u = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { t += red[w]; u += red[8 + w]; }
    partials[2 * blockIdx.x] = t;
    partials[2 * blockIdx.x + 1] = u;
    __threadfence();
    unsigned int done = atomicAdd(&ctrl->counter_kick, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0, u = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
      t += ((volatile double *)partials)[2 * b];
      u += ((volatile double *)partials)[2 * b + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
      t += __shfl_xor_sync(0xffffffffu, t, o);
      u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = t; red[8 + (threadIdx.x >> 5)] = u; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0, uu = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) { tt += red[w]; uu += red[8 + w]; }
      ctrl->ekin = tt;
      ctrl->epot = uu;
      ctrl->counter_kick = 0u;
      ctrl->steps_done += count_step;
    }
  }
}

__global__ void k_dd_empty_step(DdCtrl *ctrl, int count_step) {
  if (ctrl->stop) return;
  ctrl->ekin = 0.0;
  ctrl->epot = 0.0;
  ctrl->steps_done += count_step;
}

// flags for migration / ghost selection; s = Bbox(1,:) . (r + torig) is the fractional coordinate
__global__ void k_dd_flags(int nown, const double *__restrict__ r, double b0, double b1, double b2,
                           double t0, double t1, double t2, double lo, double hi, int mode,
                           int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nown) return;
  double s = b0 * (r[3 * i] + t0) + b1 * (r[3 * i + 1] + t1) + b2 * (r[3 * i + 2] + t2);
  int fl;
  if (mode == 0) fl = (s >= lo && s < hi);  // stay
  else if (mode == 1) fl = (s < lo);        // below
  else fl = (s >= hi);                      // above or equal
  flags[i] = fl;
}

// role per SORTED atom: 2 owned, 1 inner ghost (within rc+skin of the slab), 0 outer ghost
__global__ void k_dd_roles(int nloc, int nown, const int *__restrict__ order, const double *__restrict__ r,
                           double b0, double b1, double b2, double t0, double t1, double t2, double lo,
                           double hi, unsigned char *__restrict__ role) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  int i = order[s];
  unsigned char q = 2;
  if (i >= nown) {
    double x = b0 * (r[3 * i] + t0) + b1 * (r[3 * i + 1] + t1) + b2 * (r[3 * i + 2] + t2);
    q = (x >= lo && x < hi) ? 1 : 0;
  }
  role[s] = q;
}

__global__ void k_dd_iota(int n, int *a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

__global__ void k_dd_pack_rows(int n, const int *__restrict__ idx, const double *__restrict__ idd,
                               const int *__restrict__ el, const double *__restrict__ r,
                               const double *__restrict__ v, const double *__restrict__ minv,
                               double sx, double sy, double sz, double *__restrict__ rows) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = idx[t];
  double *o = rows + (size_t)DD_ROW * t;
  o[0] = idd[i];
  o[1] = (double)el[i];
  o[2] = r[3 * i] + sx; o[3] = r[3 * i + 1] + sy; o[4] = r[3 * i + 2] + sz;
  o[5] = v[3 * i]; o[6] = v[3 * i + 1]; o[7] = v[3 * i + 2];
  o[8] = minv[i];
}

__global__ void k_dd_unpack_rows(int n, int at, const double *__restrict__ rows, double sx, double sy,
                                 double sz, double *__restrict__ idd, int *__restrict__ el,
                                 double *__restrict__ r, double *__restrict__ v,
                                 double *__restrict__ minv) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = at + t;
  const double *o = rows + (size_t)DD_ROW * t;
  idd[i] = o[0];
  el[i] = (int)o[1];
  r[3 * i] = o[2] + sx; r[3 * i + 1] = o[3] + sy; r[3 * i + 2] = o[4] + sz;
  v[3 * i] = o[5]; v[3 * i + 1] = o[6]; v[3 * i + 2] = o[7];
  minv[i] = o[8];
}

// ghost record: el, r(3)
__global__ void k_dd_pack_ghost(int n, const int *__restrict__ idx, const int *__restrict__ el,
                                const double *__restrict__ r, double sx, double sy, double sz,
                                double *__restrict__ rows) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = idx[t];
  rows[4 * t] = (double)el[i];
  rows[4 * t + 1] = r[3 * i] + sx; rows[4 * t + 2] = r[3 * i + 1] + sy; rows[4 * t + 3] = r[3 * i + 2] + sz;
}

__global__ void k_dd_unpack_ghost(int n, int at, const double *__restrict__ rows, int *__restrict__ el,
                                  double *__restrict__ r) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = at + t;
  el[i] = (int)rows[4 * t];
  r[3 * i] = rows[4 * t + 1]; r[3 * i + 1] = rows[4 * t + 2]; r[3 * i + 2] = rows[4 * t + 3];
}

__global__ void k_dd_gather_owned(int n, const int *__restrict__ idx, const double *__restrict__ idd,
                                  const int *__restrict__ el, const double *__restrict__ r,
                                  const double *__restrict__ v, const double *__restrict__ minv,
                                  double *__restrict__ idd2, int *__restrict__ el2,
                                  double *__restrict__ r2, double *__restrict__ v2,
                                  double *__restrict__ minv2) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = idx[t];
  idd2[t] = idd[i];
  el2[t] = el[i];
  for (int c = 0; c < 3; c++) { r2[3 * t + c] = r[3 * i + c]; v2[3 * t + c] = v[3 * i + c]; }
  minv2[t] = minv[i];
}

template <typename T>
static void swapbuf(DevBuf<T> &a, DevBuf<T> &b) {
  std::swap(a.ptr, b.ptr);
  std::swap(a.cap, b.cap);
}

// stable selection of the indices i < n with flags[i] != 0 into out; the count lands in
// md->cnt[8 + slot] on the device (read back by dd_counts_roundtrip)
static int dd_select(atx_ddmd *md, int n, const int *flags, int *out, int slot) {
  atx_ctx *ctx = md->ctx;
  ATX_PASS(md->tmpi.reserve(n + 1));
  ATX_PASS(md->cnt.reserve(16));
  if (n == 0) {
    ATX_CUDA(cudaMemsetAsync(md->cnt.ptr + 8 + slot, 0, sizeof(int), ctx->stream));
    return 0;
  }
  k_dd_iota<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, md->tmpi.ptr);
  ATX_LAUNCHED();
  size_t bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, bytes, md->tmpi.ptr, flags, out, md->cnt.ptr + 8 + slot, n, ctx->stream);
  ATX_PASS(ctx->cub_tmp.reserve(bytes));
  ATX_CUDA(cub::DeviceSelect::Flagged(ctx->cub_tmp.ptr, bytes, md->tmpi.ptr, flags, out,
                                      md->cnt.ptr + 8 + slot, n, ctx->stream));
  g_atx_launches += 2;
  return 0;
}

// counts of what I send left / right (device slots cnt[8], cnt[9], written by dd_select) go to the slab
// neighbours straight from device memory; what they send me lands in cnt[12] (from the left) and cnt[13]
// (from the right); ONE host round trip then returns cnt[8 .. 13] (own counts incl. slot 10, received counts)
static int dd_counts_roundtrip(atx_ddmd *md, int *own3, int *recvL, int *recvR) {
  atx_dd *dd = md->dd;
  cudaStream_t st = md->ctx->stream;
  int *d = md->cnt.ptr;
  ATX_CUDA(cudaMemsetAsync(d + 12, 0, 2 * sizeof(int), st));
  ATX_NCCL(g_nccl.GroupStart());
  if (md->left >= 0) ATX_NCCL(g_nccl.Send(d + 8, 1, ncclInt, md->left, dd->comm, st));
  if (md->right >= 0) ATX_NCCL(g_nccl.Send(d + 9, 1, ncclInt, md->right, dd->comm, st));
  if (md->left >= 0 && md->left == md->right) {
    // two ranks, periodic: the peer's first message is what it sent to ITS left, i.e. my right
    ATX_NCCL(g_nccl.Recv(d + 13, 1, ncclInt, md->right, dd->comm, st));
    ATX_NCCL(g_nccl.Recv(d + 12, 1, ncclInt, md->left, dd->comm, st));
  } else {
    if (md->left >= 0) ATX_NCCL(g_nccl.Recv(d + 12, 1, ncclInt, md->left, dd->comm, st));
    if (md->right >= 0) ATX_NCCL(g_nccl.Recv(d + 13, 1, ncclInt, md->right, dd->comm, st));
  }
  ATX_NCCL(g_nccl.GroupEnd());
  ATX_CUDA(cudaMemcpyAsync(md->hcnt.ptr + 8, d + 8, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 3; k++) own3[k] = md->hcnt.ptr[8 + k];
  *recvL = md->hcnt.ptr[12];
  *recvR = md->hcnt.ptr[13];
  return 0;
}

// Note on send/recv matching with two ranks and periodic x: left == right == the other rank.  My
// "send left" is the peer's "receive from right"; NCCL matches operations to the same peer in issue
// order, so sends are issued (left, right) and receives (right-of-peer == from left first...) --
// to keep both sides consistent every exchange issues: send L, send R, recv from R, recv from L when
// left == right, which pairs my L-send with the peer's R-recv.
static int dd_sendrecv(atx_ddmd *md, const double *sL, size_t nL, const double *sR, size_t nR,
                       double *rL, size_t mL, double *rR, size_t mR) {
  atx_dd *dd = md->dd;
  cudaStream_t st = md->ctx->stream;
  ATX_NCCL(g_nccl.GroupStart());
  if (md->left >= 0) ATX_NCCL(g_nccl.Send(sL, nL, ncclDouble, md->left, dd->comm, st));
  if (md->right >= 0) ATX_NCCL(g_nccl.Send(sR, nR, ncclDouble, md->right, dd->comm, st));
  if (md->left >= 0 && md->left == md->right) {
    ATX_NCCL(g_nccl.Recv(rR, mR, ncclDouble, md->right, dd->comm, st));
    ATX_NCCL(g_nccl.Recv(rL, mL, ncclDouble, md->left, dd->comm, st));
  } else {
    if (md->left >= 0) ATX_NCCL(g_nccl.Recv(rL, mL, ncclDouble, md->left, dd->comm, st));
    if (md->right >= 0) ATX_NCCL(g_nccl.Recv(rR, mR, ncclDouble, md->right, dd->comm, st));
  }
  ATX_NCCL(g_nccl.GroupEnd());
  return 0;
}

static int dd_reserve_local(atx_ddmd *md, size_t n) {
  // grow all per-local-atom arrays (main and alternate set), preserving the owned part
  if (n + 1 <= md->r.cap / 3 && n + 1 <= md->el.cap && n + 1 <= md->r2.cap / 3 && n + 1 <= md->el2.cap) return 0;
  size_t want = n + n / 4 + 4096;
  cudaStream_t st = md->ctx->stream;
  auto grow_d = [&](DevBuf<double> &b, size_t per, size_t keep) -> int {
    if (b.cap >= per * want) return 0;
    DevBuf<double> nb;
    ATX_PASS(nb.reserve(per * want));
    if (keep && b.ptr) ATX_CUDA(cudaMemcpyAsync(nb.ptr, b.ptr, sizeof(double) * keep, cudaMemcpyDeviceToDevice, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    swapbuf(b, nb);
    return 0;
  };
  size_t no = md->nown;
  ATX_PASS(grow_d(md->r, 3, 3 * no));
  ATX_PASS(grow_d(md->v, 3, 3 * no));
  ATX_PASS(grow_d(md->f, 3, 3 * no));
  ATX_PASS(grow_d(md->minv, 1, no));
  ATX_PASS(grow_d(md->idd, 1, no));
  ATX_PASS(grow_d(md->r2, 3, 0));
  ATX_PASS(grow_d(md->v2, 3, 0));
  ATX_PASS(grow_d(md->f2, 3, 0));
  ATX_PASS(grow_d(md->minv2, 1, 0));
  ATX_PASS(grow_d(md->idd2, 1, 0));
  if (md->el.cap < want) {
    DevBuf<int> nb;
    ATX_PASS(nb.reserve(want));
    if (no && md->el.ptr) ATX_CUDA(cudaMemcpyAsync(nb.ptr, md->el.ptr, sizeof(int) * no, cudaMemcpyDeviceToDevice, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    swapbuf(md->el, nb);
  }
  ATX_PASS(md->el2.reserve(want));
  ATX_PASS(md->flags.reserve(want));
  ATX_PASS(md->flags2.reserve(want));
  ATX_PASS(md->flags3.reserve(want));
  ATX_PASS(md->sel.reserve(want));
  ATX_PASS(md->sendL.reserve(want));
  ATX_PASS(md->sendR.reserve(want));
  ATX_PASS(md->role.reserve(want));
  ATX_PASS(md->role_loc.reserve(want));
  return 0;
}

static int dd_compute(atx_ddmd *md, bool guarded, int phase = 0, bool need_energy = true) {
  PotOut o;
  o.phase = phase;
  o.split_lo = md->split_lo;
  o.split_hi = md->split_hi;
  int nloc = md->nown + md->ngl + md->ngr;
  ATX_PASS(md->tmpd.reserve(3 * (size_t)nloc + 3));
  ATX_PASS(md->epa.reserve((size_t)nloc + 1));
  o.f = md->tmpd.ptr;  // sorted order
  o.epa = need_energy ? md->epa.ptr : nullptr;   // per-atom energies: only the kick that ends a run sums them
  o.sums = md->sums.ptr;
  o.stop = guarded ? &md->ctrl.ptr->stop : nullptr;
  o.want_virial = false;
  o.want_sums = false;   // the kick kernel sums the per-atom energies of the owned atoms itself
  o.role = md->dd->nranks > 1 ? md->role.ptr : nullptr;
  switch (md->pot_kind) {
    case ATX_POT_EAM:
      return atx_eam_compute_device((atx_eam *)md->pot, &md->ploc, md->nl, nullptr, o);
    case ATX_POT_BOP:
      return atx_bop_compute_device((atx_bop *)md->pot, &md->ploc, md->nl, nullptr, o);
    case ATX_POT_REBO2:
      return atx_rebo2_compute_device((atx_rebo2 *)md->pot, &md->ploc, md->nl, o);
  }
  atx_set_error("atx_dd_md: unknown potential kind");
  return ATX_ERROR_UNSPECIFIED;
}

// migration + ghost construction + local neighbour list
struct DdTimer {
  atx_ddmd *md;
  bool exact;
  std::chrono::steady_clock::time_point t;
  explicit DdTimer(atx_ddmd *m) : md(m) {
    static const bool ex = getenv("ATX_DD_PROFILE") && atoi(getenv("ATX_DD_PROFILE")) != 0;
    exact = ex;
    if (exact) cudaStreamSynchronize(md->ctx->stream);
    t = std::chrono::steady_clock::now();
  }
  void lap(int slot) {
    if (exact) cudaStreamSynchronize(md->ctx->stream);
    auto n = std::chrono::steady_clock::now();
    md->rebuild_host_ms[slot] += std::chrono::duration<double, std::milli>(n - t).count();
    t = n;
  }
};

static int dd_rebuild(atx_ddmd *md) {
  atx_ctx *ctx = md->ctx;
  cudaStream_t st = ctx->stream;
  DdTimer tm(md);
  const double b0 = md->B.m[0], b1 = md->B.m[3], b2 = md->B.m[6];  // Bbox(1,:)
  const double *t = md->torig;
  int n = md->nown;
  int gb = (n + 255) / 256;
  ATX_PASS(dd_reserve_local(md, n));

  // ---- migration: atoms outside [slo, shi) go to the neighbouring slab
  int nstay = n, nL = 0, nR = 0;
  if (md->dd->nranks > 1) {
    if (n > 0) {
      k_dd_flags<<<gb, 256, 0, st>>>(n, md->r.ptr, b0, b1, b2, t[0], t[1], t[2], md->slo_own, md->shi_own, 1, md->flags.ptr);
      k_dd_flags<<<gb, 256, 0, st>>>(n, md->r.ptr, b0, b1, b2, t[0], t[1], t[2], md->slo_own, md->shi_own, 2, md->flags2.ptr);
      k_dd_flags<<<gb, 256, 0, st>>>(n, md->r.ptr, b0, b1, b2, t[0], t[1], t[2], md->slo_own, md->shi_own, 0, md->flags3.ptr);
      g_atx_launches += 3;
    }
    ATX_PASS(dd_select(md, n, md->flags.ptr, md->sendL.ptr, 0));
    ATX_PASS(dd_select(md, n, md->flags2.ptr, md->sendR.ptr, 1));
    ATX_PASS(dd_select(md, n, md->flags3.ptr, md->sel.ptr, 2));
    int c3[3];
    int inL = 0, inR = 0;
    ATX_PASS(dd_counts_roundtrip(md, c3, &inL, &inR));
    nL = c3[0]; nR = c3[1]; nstay = c3[2];
    tm.lap(0);
    if ((md->left < 0 && nL > 0) || (md->right < 0 && nR > 0)) {
      atx_set_error("Particle outside simulation domain (left the non-periodic box along x).");
      return ATX_ERROR_UNSPECIFIED;
    }
    tm.lap(1);
    ATX_PASS(md->mig_send.reserve((size_t)DD_ROW * (nL + nR) + DD_ROW));
    ATX_PASS(md->mig_recv.reserve((size_t)DD_ROW * (inL + inR) + DD_ROW));
    // positions travel in the GLOBAL frame (+ periodic wrap), the receiver subtracts its origin
    if (nL > 0) {
      k_dd_pack_rows<<<(nL + 127) / 128, 128, 0, st>>>(nL, md->sendL.ptr, md->idd.ptr, md->el.ptr, md->r.ptr,
                                                       md->v.ptr, md->minv.ptr, t[0] + md->wrapL * md->a1[0],
                                                       t[1] + md->wrapL * md->a1[1], t[2] + md->wrapL * md->a1[2],
                                                       md->mig_send.ptr);
      ATX_LAUNCHED();
    }
    if (nR > 0) {
      k_dd_pack_rows<<<(nR + 127) / 128, 128, 0, st>>>(nR, md->sendR.ptr, md->idd.ptr, md->el.ptr, md->r.ptr,
                                                       md->v.ptr, md->minv.ptr, t[0] + md->wrapR * md->a1[0],
                                                       t[1] + md->wrapR * md->a1[1], t[2] + md->wrapR * md->a1[2],
                                                       md->mig_send.ptr + (size_t)DD_ROW * nL);
      ATX_LAUNCHED();
    }
    if (nL + nR + inL + inR > 0 || true) {
      ATX_PASS(dd_sendrecv(md, md->mig_send.ptr, (size_t)DD_ROW * nL, md->mig_send.ptr + (size_t)DD_ROW * nL,
                           (size_t)DD_ROW * nR, md->mig_recv.ptr, (size_t)DD_ROW * inL,
                           md->mig_recv.ptr + (size_t)DD_ROW * inL, (size_t)DD_ROW * inR));
    }
    // compact the stayers into the alternate arrays, append the arrivals, swap
    int nnew = nstay + inL + inR;
    ATX_PASS(dd_reserve_local(md, (size_t)nnew));
    if (nstay > 0) {
      k_dd_gather_owned<<<(nstay + 255) / 256, 256, 0, st>>>(nstay, md->sel.ptr, md->idd.ptr, md->el.ptr, md->r.ptr,
                                                            md->v.ptr, md->minv.ptr, md->idd2.ptr, md->el2.ptr,
                                                            md->r2.ptr, md->v2.ptr, md->minv2.ptr);
      ATX_LAUNCHED();
    }
    if (inL > 0) {
      k_dd_unpack_rows<<<(inL + 127) / 128, 128, 0, st>>>(inL, nstay, md->mig_recv.ptr, -t[0], -t[1], -t[2],
                                                          md->idd2.ptr, md->el2.ptr, md->r2.ptr, md->v2.ptr,
                                                          md->minv2.ptr);
      ATX_LAUNCHED();
    }
    if (inR > 0) {
      k_dd_unpack_rows<<<(inR + 127) / 128, 128, 0, st>>>(inR, nstay + inL, md->mig_recv.ptr + (size_t)DD_ROW * inL,
                                                          -t[0], -t[1], -t[2], md->idd2.ptr, md->el2.ptr, md->r2.ptr,
                                                          md->v2.ptr, md->minv2.ptr);
      ATX_LAUNCHED();
    }
    swapbuf(md->r, md->r2); swapbuf(md->v, md->v2); swapbuf(md->f, md->f2); swapbuf(md->minv, md->minv2);
    swapbuf(md->idd, md->idd2); swapbuf(md->el, md->el2);
    md->nown = n = nnew;
    gb = (n + 255) / 256;
    tm.lap(2);
  }

  // ---- ghosts: owned atoms within the halo of a face are sent to that neighbour
  md->nsendL = md->nsendR = md->ngl = md->ngr = 0;
  if (md->dd->nranks > 1) {
    ATX_CUDA(cudaMemsetAsync(md->cnt.ptr + 8, 0, 2 * sizeof(int), st));
    if (md->left >= 0) {
      if (n > 0) {
        k_dd_flags<<<gb, 256, 0, st>>>(n, md->r.ptr, b0, b1, b2, t[0], t[1], t[2], md->slo + md->hfrac, 2.0, 1,
                                       md->flags.ptr);
        ATX_LAUNCHED();
      }
      ATX_PASS(dd_select(md, n, md->flags.ptr, md->sendL.ptr, 0));
    }
    if (md->right >= 0) {
      if (n > 0) {
        k_dd_flags<<<gb, 256, 0, st>>>(n, md->r.ptr, b0, b1, b2, t[0], t[1], t[2], -1.0, md->shi - md->hfrac, 2,
                                       md->flags2.ptr);
        ATX_LAUNCHED();
      }
      ATX_PASS(dd_select(md, n, md->flags2.ptr, md->sendR.ptr, 1));
    }
    {
      int c3b[3];
      ATX_PASS(dd_counts_roundtrip(md, c3b, &md->ngl, &md->ngr));
      md->nsendL = c3b[0];
      md->nsendR = c3b[1];
    }
    tm.lap(3);
    tm.lap(4);
    int nloc = n + md->ngl + md->ngr;
    if (md->p2p && ((size_t)md->ngl > md->capG || (size_t)md->ngr > md->capG)) {
      atx_set_error("Domain decomposition: more ghost atoms than the peer-to-peer receive areas hold (" +
                    std::to_string(md->ngl) + " / " + std::to_string(md->ngr) + " of " + std::to_string(md->capG) +
                    "); set ATX_DD_P2P=0 to use the NCCL halo path.");
      return ATX_ERROR_MPI;
    }
    ATX_PASS(dd_reserve_local(md, nloc));
    ATX_PASS(md->bufL.reserve(4 * (size_t)md->nsendL + 4));
    ATX_PASS(md->bufR.reserve(4 * (size_t)md->nsendR + 4));
    ATX_PASS(md->mig_recv.reserve(4 * (size_t)(md->ngl + md->ngr) + 4));
    // ghost positions travel in the receiver's... no: global frame + wrap; receiver subtracts origin
    if (md->nsendL > 0) {
      k_dd_pack_ghost<<<(md->nsendL + 127) / 128, 128, 0, st>>>(
          md->nsendL, md->sendL.ptr, md->el.ptr, md->r.ptr, t[0] + md->wrapL * md->a1[0],
          t[1] + md->wrapL * md->a1[1], t[2] + md->wrapL * md->a1[2], md->bufL.ptr);
      ATX_LAUNCHED();
    }
    if (md->nsendR > 0) {
      k_dd_pack_ghost<<<(md->nsendR + 127) / 128, 128, 0, st>>>(
          md->nsendR, md->sendR.ptr, md->el.ptr, md->r.ptr, t[0] + md->wrapR * md->a1[0],
          t[1] + md->wrapR * md->a1[1], t[2] + md->wrapR * md->a1[2], md->bufR.ptr);
      ATX_LAUNCHED();
    }
    ATX_PASS(dd_sendrecv(md, md->bufL.ptr, 4 * (size_t)md->nsendL, md->bufR.ptr, 4 * (size_t)md->nsendR,
                         md->mig_recv.ptr, 4 * (size_t)md->ngl, md->mig_recv.ptr + 4 * (size_t)md->ngl,
                         4 * (size_t)md->ngr));
    if (md->ngl + md->ngr > 0) {
      int ng = md->ngl + md->ngr;
      k_dd_unpack_ghost<<<(ng + 127) / 128, 128, 0, st>>>(ng, n, md->mig_recv.ptr, md->el.ptr, md->r.ptr);
      ATX_LAUNCHED();
      // receiver frame: subtract the local origin (ghost rows hold global positions)
    }
    tm.lap(5);
  }
  return 0;
}

__global__ void k_dd_shift(int n, int at, double *__restrict__ r, double sx, double sy, double sz) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int i = at + t;
  r[3 * i] += sx; r[3 * i + 1] += sy; r[3 * i + 2] += sz;
}

// ---- local numbering follows the sorted order -------------------------------------------------
// After a list build the owned atoms are renumbered so that their local index grows with their
// sorted (cell-ordered) index: the per-step gathers / scatters between the local arrays (r, v, f)
// and the sorted records (pos4, forces of the potentials) then touch memory almost sequentially.
__global__ void k_dd_owned_flag(int nloc, int n, const int *__restrict__ order, int *__restrict__ flag) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s <= nloc) flag[s] = (s < nloc && order[s] < n) ? 1 : 0;
}
__global__ void k_dd_newidx(int n, const int *__restrict__ inv, const int *__restrict__ rank,
                            int *__restrict__ newidx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) newidx[i] = rank[inv[i]];
}
__global__ void k_dd_permute_owned(int n, const int *__restrict__ newidx, const double *__restrict__ idd,
                                   const int *__restrict__ el, const double *__restrict__ r,
                                   const double *__restrict__ v, const double *__restrict__ minv,
                                   double *__restrict__ idd2, int *__restrict__ el2, double *__restrict__ r2,
                                   double *__restrict__ v2, double *__restrict__ minv2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = newidx[i];
  idd2[k] = idd[i];
  el2[k] = el[i];
  minv2[k] = minv[i];
  for (int c = 0; c < 3; c++) { r2[3 * k + c] = r[3 * i + c]; v2[3 * k + c] = v[3 * i + c]; }
}
__global__ void k_dd_renumber(int nloc, int n, const int *__restrict__ newidx, int *__restrict__ order,
                              int *__restrict__ inv) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  int o = order[s];
  if (o < n) {
    o = newidx[o];
    order[s] = o;
  }
  inv[o] = s;
}
__global__ void k_dd_remap(int m, const int *__restrict__ newidx, int *__restrict__ idx) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) idx[t] = newidx[idx[t]];
}

static int dd_build_list(atx_ddmd *md) {
  atx_ctx *ctx = md->ctx;
  cudaStream_t st = ctx->stream;
  DdTimer tm(md);
  int n = md->nown, ng = md->ngl + md->ngr, nloc = n + ng;
  const double *t = md->torig;
  if (ng > 0) {
    k_dd_shift<<<(ng + 127) / 128, 128, 0, st>>>(ng, n, md->r.ptr, -t[0], -t[1], -t[2]);
    ATX_LAUNCHED();
  }
  // local particles: alias the arrays
  md->ploc.nat = nloc;
  md->ploc.r_ext = md->r.ptr;
  std::swap(md->ploc.el.ptr, md->el.ptr);  // ploc.el aliases md->el while the list is built
  std::swap(md->ploc.el.cap, md->el.cap);
  md->ploc.pos_rev++;
  md->nl->p_rev = -1;  // force a real rebuild (see md_rebuild in atx_md.cu)
  int err = atx_neighbors_update(md->nl, &md->ploc);
  std::swap(md->ploc.el.ptr, md->el.ptr);
  std::swap(md->ploc.el.cap, md->el.cap);
  if (err) return err;
  if (md->dd->nranks > 1 && n > 0) {
    // renumber the owned atoms in sorted order (flags -> ranks -> permutation of the local arrays)
    const int gl = (nloc + 1 + 255) / 256, go = (n + 255) / 256;
    ATX_PASS(md->flags.reserve((size_t)nloc + 2));
    ATX_PASS(md->flags2.reserve((size_t)nloc + 2));
    ATX_PASS(md->flags3.reserve((size_t)nloc + 2));
    k_dd_owned_flag<<<gl, 256, 0, st>>>(nloc, n, md->nl->order.ptr, md->flags.ptr);
    ATX_PASS(atx_scan_int(ctx, md->flags.ptr, md->flags2.ptr, (size_t)nloc + 1));
    k_dd_newidx<<<go, 256, 0, st>>>(n, md->nl->inv.ptr, md->flags2.ptr, md->flags3.ptr);
    k_dd_permute_owned<<<go, 256, 0, st>>>(n, md->flags3.ptr, md->idd.ptr, md->el.ptr, md->r.ptr, md->v.ptr,
                                           md->minv.ptr, md->idd2.ptr, md->el2.ptr, md->r2.ptr, md->v2.ptr,
                                           md->minv2.ptr);
    if (ng > 0) {
      ATX_CUDA(cudaMemcpyAsync(md->r2.ptr + 3 * (size_t)n, md->r.ptr + 3 * (size_t)n, sizeof(double) * 3 * ng,
                               cudaMemcpyDeviceToDevice, st));
      ATX_CUDA(cudaMemcpyAsync(md->el2.ptr + n, md->el.ptr + n, sizeof(int) * ng, cudaMemcpyDeviceToDevice, st));
    }
    k_dd_renumber<<<(nloc + 255) / 256, 256, 0, st>>>(nloc, n, md->flags3.ptr, md->nl->order.ptr, md->nl->inv.ptr);
    if (md->nsendL > 0) k_dd_remap<<<(md->nsendL + 255) / 256, 256, 0, st>>>(md->nsendL, md->flags3.ptr, md->sendL.ptr);
    if (md->nsendR > 0) k_dd_remap<<<(md->nsendR + 255) / 256, 256, 0, st>>>(md->nsendR, md->flags3.ptr, md->sendR.ptr);
    g_atx_launches += 6;
    swapbuf(md->r, md->r2); swapbuf(md->v, md->v2); swapbuf(md->f, md->f2); swapbuf(md->minv, md->minv2);
    swapbuf(md->idd, md->idd2); swapbuf(md->el, md->el2);
    md->ploc.r_ext = md->r.ptr;
  }
  if (md->dd->nranks > 1 && nloc > 0) {
    // inner ghosts: within rc + skin (= half the halo) of the slab; + a small margin
    const double band = 0.5 * md->hfrac * 1.02;
    k_dd_roles<<<(nloc + 255) / 256, 256, 0, st>>>(nloc, n, md->nl->order.ptr, md->r.ptr, md->B.m[0], md->B.m[3],
                                                   md->B.m[6], t[0], t[1], t[2], md->slo - band, md->shi + band,
                                                   md->role.ptr);
    ATX_LAUNCHED();
  }
  // Opt-in (ATX_DD_SPLIT=1): measured at N=8 on C4 the three centre launches per step lose in partial
  // waves (0.82 vs 0.77 ms of centre kernels per step) what the hidden halo wait (0.07 ms) gains.
  // interior atoms: owned atoms further than rc + skin (half the halo) from both slab faces at build
  // time have only owned atoms in their lists.  Cells are ordered x-major, so the atoms of the cell
  // planes that lie completely inside that region form ONE range of the sorted numbering.
  md->split = false;
  md->split_lo = md->split_hi = 0;
  if (md->dd->nranks > 1 && md->p2p && md->pot_kind == ATX_POT_BOP && atx_bop_supports_split((atx_bop *)md->pot) &&
      nloc > 0 && getenv("ATX_DD_SPLIT") && atoi(getenv("ATX_DD_SPLIT")) != 0) {
    const double alpha = 1.0 / md->dd->nranks + 2.0 * md->hfrac;
    const double u_lo = (md->hfrac + 0.51 * md->hfrac) / alpha, u_hi = (md->hfrac + 1.0 / md->dd->nranks - 0.51 * md->hfrac) / alpha;
    const int n0 = md->nl->n_cells[0], plane = md->nl->n_cells[1] * md->nl->n_cells[2];
    int c_lo = (int)std::ceil(u_lo * n0), c_hi = (int)std::floor(u_hi * n0);
    if (c_lo < 0) c_lo = 0;
    if (c_hi > n0) c_hi = n0;
    if (c_hi > c_lo) {
      int h2[2] = {0, 0};
      ATX_CUDA(cudaMemcpyAsync(&h2[0], md->nl->cell_start.ptr + (size_t)c_lo * plane, sizeof(int), cudaMemcpyDeviceToHost, st));
      ATX_CUDA(cudaMemcpyAsync(&h2[1], md->nl->cell_start.ptr + (size_t)c_hi * plane, sizeof(int), cudaMemcpyDeviceToHost, st));
      ATX_CUDA(cudaStreamSynchronize(st));
      if (h2[1] > h2[0]) {
        md->split_lo = h2[0];
        md->split_hi = h2[1];
        md->split = true;
      }
    }
  }
  md->nrebuilds++;
  tm.lap(6);
  return 0;
}

static int dd_reset_ctrl(atx_ddmd *md) {
  DdCtrl *h = md->hctrl.ptr;
  h->stop = 0;
  h->want = 0;
  h->accum_max_dr = 1e-6;
  h->stepmax_bits = 0;
  h->counter_drift = 0;
  h->counter_kick = 0;
  ATX_CUDA(cudaMemcpyAsync(md->ctrl.ptr, h, sizeof(DdCtrl), cudaMemcpyHostToDevice, md->ctx->stream));
  return 0;
}

static int dd_kick(atx_ddmd *md, int count_step = 1) {
  ProfScope ps_(md->ctx, "dd_kick");
  cudaStream_t st = md->ctx->stream;
  int n = md->nown;
  if (n > 0) {
    int gb = (n + 255) / 256;
    ATX_PASS(md->kin_partials.reserve(2 * (size_t)gb + 2));
    k_dd_kick<<<gb, 256, 0, st>>>(n, md->dt, md->v.ptr, md->tmpd.ptr, md->epa.ptr, md->nl->inv.ptr, md->minv.ptr,
                                  md->f.ptr, md->kin_partials.ptr, md->ctrl.ptr, count_step);
  } else {
    k_dd_empty_step<<<1, 1, 0, st>>>(md->ctrl.ptr, count_step);
  }
  ATX_LAUNCHED();
  return 0;
}

// Map every peer's mailbox (collective over the communicator).  Falls back to the NCCL step path
// (md->p2p stays false) when any rank cannot export / map the handles; ATX_DD_P2P=0 forces that.
static int dd_p2p_setup(atx_ddmd *md) {
  atx_dd *dd = md->dd;
  cudaStream_t st = md->ctx->stream;
  const int P = dd->nranks, me = dd->rank;
  if (P < 2 || P > DD_MAXP) return 0;
  int enable = 1;
  if (const char *v = getenv("ATX_DD_P2P")) enable = atoi(v) != 0;
  // scratch: [0] capacity wish -> max, [1] ok flag -> min, then P ipc handles
  DevBuf<int> d_i;
  DevBuf<char> d_h;
  ATX_PASS(d_i.reserve(4));
  ATX_PASS(d_h.reserve((size_t)P * sizeof(cudaIpcMemHandle_t)));
  int g = md->ngl > md->ngr ? md->ngl : md->ngr;
  if (md->nsendL > g) g = md->nsendL;
  if (md->nsendR > g) g = md->nsendR;
  int h_i[2] = {g + g / 2 + 4096, 0};
  ATX_CUDA(cudaMemcpyAsync(d_i.ptr, h_i, sizeof(int), cudaMemcpyHostToDevice, st));
  ATX_NCCL(g_nccl.AllReduce(d_i.ptr, d_i.ptr, 1, ncclInt, ncclMax, dd->comm, st));
  ATX_CUDA(cudaMemcpyAsync(h_i, d_i.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  md->capG = (size_t)h_i[0];
  const size_t bytes = DD_SIG_BYTES + 4 * 3 * md->capG * sizeof(double);
  int ok = enable;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc((void **)&md->mbox, bytes) != cudaSuccess) { ok = 0; md->mbox = nullptr; cudaGetLastError(); }
  if (ok) {
    ATX_CUDA(cudaMemsetAsync(md->mbox, 0, bytes, st));
    if (cudaIpcGetMemHandle(&mine, md->mbox) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  ATX_CUDA(cudaMemcpyAsync(d_h.ptr + (size_t)me * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  ATX_NCCL(g_nccl.AllGather(d_h.ptr + (size_t)me * sizeof(mine), d_h.ptr, sizeof(mine), ncclChar, dd->comm, st));
  std::vector<cudaIpcMemHandle_t> all(P);
  ATX_CUDA(cudaMemcpyAsync(all.data(), d_h.ptr, (size_t)P * sizeof(mine), cudaMemcpyDeviceToHost, st));
  h_i[1] = ok;
  ATX_CUDA(cudaMemcpyAsync(d_i.ptr + 1, h_i + 1, sizeof(int), cudaMemcpyHostToDevice, st));
  ATX_NCCL(g_nccl.AllReduce(d_i.ptr + 1, d_i.ptr + 1, 1, ncclInt, ncclMin, dd->comm, st));
  ATX_CUDA(cudaMemcpyAsync(h_i + 1, d_i.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ok = h_i[1];
  if (ok) {
    for (int p = 0; p < P; p++) {
      if (p == me) { md->peer_mbox[p] = md->mbox; continue; }
      void *q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
        break;
      }
      md->peer_mbox[p] = (char *)q;
    }
  }
  // everybody must have mapped everybody before the first remote store / the fallback decision
  h_i[1] = ok;
  ATX_CUDA(cudaMemcpyAsync(d_i.ptr + 1, h_i + 1, sizeof(int), cudaMemcpyHostToDevice, st));
  ATX_NCCL(g_nccl.AllReduce(d_i.ptr + 1, d_i.ptr + 1, 1, ncclInt, ncclMin, dd->comm, st));
  ATX_CUDA(cudaMemcpyAsync(h_i + 1, d_i.ptr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ok = h_i[1];
  if (!ok) {
    for (int p = 0; p < P; p++) {
      if (p != me && md->peer_mbox[p]) cudaIpcCloseMemHandle(md->peer_mbox[p]);
      md->peer_mbox[p] = nullptr;
    }
    if (md->mbox) cudaFree(md->mbox);
    md->mbox = nullptr;
    cudaGetLastError();
    md->p2p = false;
    return 0;
  }
  std::vector<unsigned long long *> sigs(DD_MAXP, nullptr);
  for (int p = 0; p < P; p++) sigs[p] = (unsigned long long *)md->peer_mbox[p];
  ATX_PASS(md->d_peer_sig.reserve(DD_MAXP));
  ATX_CUDA(cudaMemcpyAsync(md->d_peer_sig.ptr, sigs.data(), sizeof(unsigned long long *) * DD_MAXP,
                           cudaMemcpyHostToDevice, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  md->p2p = true;
  return 0;
}

extern "C" int atx_dd_md_create(atx_dd *dd, int pot_kind, void *pot, const double *Abox,
                                const double *Bbox, const int *pbc, double rc, double skin, int avgn,
                                int nown, const long long *id, const int *el, const double *r,
                                const double *v, const double *mass, double dt, atx_ddmd **out) {
  if (dd && dd->ctx) cudaSetDevice(dd->ctx->device);
  if (!dd || !pot || !out) return ATX_ERROR_UNSPECIFIED;
  atx_ctx *ctx = dd->ctx;
  cudaStream_t st = ctx->stream;
  ATX_CUDA(cudaSetDevice(ctx->device));
  atx_ddmd *md = new atx_ddmd();
  md->dd = dd;
  md->ctx = ctx;
  md->pot_kind = pot_kind;
  md->pot = pot;
  for (int i = 0; i < 9; i++) { md->A.m[i] = Abox[i]; md->B.m[i] = Bbox[i]; }
  for (int k = 0; k < 3; k++) { md->pbc[k] = pbc[k] != 0; md->a1[k] = Abox[k]; }
  md->dt = dt; md->rc = rc; md->skin = skin;
  const int P = dd->nranks, me = dd->rank;
  md->slo = (double)me / P;
  md->shi = (double)(me + 1) / P;
  const double slo_own = (!md->pbc[0] && me == 0) ? -1e300 : md->slo;      // open ends own the rest
  const double shi_own = (!md->pbc[0] && me == P - 1) ? 1e300 : md->shi;
  // thickness of the whole cell between the faces spanned by a2, a3: 1/|Bbox(1,:)|
  double bn = std::sqrt(Bbox[0] * Bbox[0] + Bbox[3] * Bbox[3] + Bbox[6] * Bbox[6]);
  // EAM / BOP gather: 2 cutoffs.  REBO2 evaluates every bond among its local atoms and keeps the owned
  // share (rb_force_atom<ROLES>): the forces on an owned atom depend on atoms up to five bonds away.
  const double hmul = pot_kind == ATX_POT_REBO2 ? 5.0 : 2.0;
  double halo = P > 1 ? hmul * (rc + skin) : 0.0;
  md->hfrac = halo * bn;
  if (P > 1 && md->hfrac > 1.0 / P) {
    atx_set_error("Domain decomposition: slabs are thinner than the halo (2*(rc+skin), REBO2: 5*(rc+skin)); use "
                  "fewer ranks.");
    delete md;
    return ATX_ERROR_UNSPECIFIED;
  }
  if (P > 1) {
    md->left = me > 0 ? me - 1 : (md->pbc[0] ? P - 1 : -1);
    md->right = me < P - 1 ? me + 1 : (md->pbc[0] ? 0 : -1);
    md->wrapL = me == 0 ? 1.0 : 0.0;       // leaving through the lower face of the global cell
    md->wrapR = me == P - 1 ? -1.0 : 0.0;
  }
  md->slo_own = slo_own;
  md->shi_own = shi_own;
  for (int k = 0; k < 3; k++) md->torig[k] = (md->slo - md->hfrac) * md->a1[k];
  // local cell: (ds + 2h) a1, a2, a3; non-periodic along a1 when decomposed
  double alpha = P > 1 ? (1.0 / P + 2.0 * md->hfrac) : 1.0;
  md->ploc.ctx = ctx;
  md->ploc.Abox = md->A;
  md->ploc.Bbox = md->B;
  for (int k = 0; k < 3; k++) {
    md->ploc.Abox.m[k] = md->A.m[k] * alpha;          // column 1
    md->ploc.Bbox.m[3 * k] = md->B.m[3 * k] / alpha;  // row 1
    md->ploc.pbc[k] = md->pbc[k];
  }
  if (P > 1) md->ploc.pbc[0] = 0;
  md->ploc.cell_rev = 1;
  ATX_PASS(atx_neighbors_create(ctx, avgn, &md->nl));
  ATX_PASS(atx_neighbors_request_interaction_range(md->nl, rc));
  ATX_PASS(atx_neighbors_set_verlet_shell(md->nl, skin));

  md->nown = 0;
  ATX_PASS(dd_reserve_local(md, (size_t)nown + nown / 2 + 1024));
  ATX_PASS(md->sums.reserve(ATX_NSUM));
  ATX_PASS(md->ctrl.reserve(1));
  ATX_PASS(md->hctrl.reserve(1));
  ATX_PASS(md->hcnt.reserve(16));
  ATX_PASS(md->stage.reserve(11 * (size_t)nown + 64));
  double *h = md->stage.ptr;
  // local frame = global - torig
  for (int i = 0; i < nown; i++) {
    // periodic along a1: take the image nearest to this rank's slab centre.  Callers either assign
    // owners by the wrapped fractional coordinate and pass unwrapped positions, or generate a slab
    // and displace atoms slightly across its faces; both end up within one hop of their owner.
    double wrap = 0.0;
    if (md->pbc[0])
      wrap = std::floor(Bbox[0] * r[3 * i] + Bbox[3] * r[3 * i + 1] + Bbox[6] * r[3 * i + 2] -
                        0.5 * (md->slo + md->shi) + 0.5);
    for (int c = 0; c < 3; c++) h[3 * i + c] = r[3 * i + c] - wrap * md->a1[c] - md->torig[c];
    h[3 * (size_t)nown + i] = 1.0 / mass[i];
    h[4 * (size_t)nown + i] = (double)id[i];
    for (int c = 0; c < 3; c++) h[5 * (size_t)nown + 3 * i + c] = v ? v[3 * i + c] : 0.0;
  }
  if (nown > 0) {
    ATX_CUDA(cudaMemcpyAsync(md->r.ptr, h, sizeof(double) * 3 * nown, cudaMemcpyHostToDevice, st));
    ATX_CUDA(cudaMemcpyAsync(md->minv.ptr, h + 3 * (size_t)nown, sizeof(double) * nown, cudaMemcpyHostToDevice, st));
    ATX_CUDA(cudaMemcpyAsync(md->idd.ptr, h + 4 * (size_t)nown, sizeof(double) * nown, cudaMemcpyHostToDevice, st));
    ATX_CUDA(cudaMemcpyAsync(md->v.ptr, h + 5 * (size_t)nown, sizeof(double) * 3 * nown, cudaMemcpyHostToDevice, st));
    ATX_CUDA(cudaMemcpyAsync(md->el.ptr, el, sizeof(int) * nown, cudaMemcpyHostToDevice, st));
  }
  ATX_CUDA(cudaStreamSynchronize(st));
  md->nown = nown;
  DdCtrl c{};
  c.accum_max_dr = 1e-6;
  c.verlet_shell = skin;
  *md->hctrl.ptr = c;
  ATX_CUDA(cudaMemcpyAsync(md->ctrl.ptr, md->hctrl.ptr, sizeof(DdCtrl), cudaMemcpyHostToDevice, st));
  ATX_CUDA(cudaEventCreate(&md->ev0));
  ATX_CUDA(cudaEventCreate(&md->ev1));
  int err = dd_rebuild(md);
  if (!err) err = dd_build_list(md);
  if (!err) err = dd_compute(md, false);
  if (!err) {
    // forces of the owned atoms in local order for the first drift
    int n = md->nown;
    if (n > 0) err = atx_unsort(ctx, md->nown + md->ngl + md->ngr, 3, md->nl->order.ptr, md->tmpd.ptr, md->f.ptr);
  }
  if (!err) err = dd_p2p_setup(md);
  if (err) {
    delete md;
    return err;
  }
  ATX_CUDA(cudaStreamSynchronize(st));
  *out = md;
  return 0;
}

extern "C" int atx_dd_md_destroy(atx_ddmd *md) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  if (!md) return 0;
  if (md->ev0) cudaEventDestroy(md->ev0);
  if (md->ev1) cudaEventDestroy(md->ev1);
  md->ploc.r_ext = nullptr;
  if (md->nl) atx_neighbors_destroy(md->nl);
  if (md->p2p) {
    cudaStreamSynchronize(md->ctx->stream);
    for (int p = 0; p < md->dd->nranks && p < DD_MAXP; p++)
      if (p != md->dd->rank && md->peer_mbox[p]) cudaIpcCloseMemHandle(md->peer_mbox[p]);
    if (md->mbox) cudaFree(md->mbox);
  }
  delete md;
  return 0;
}

// scale: 0 = separate drift and kick kernels; 0.5 / 1 = leapfrog form (first / later step of a run)
static int dd_enqueue_step(atx_ddmd *md, double scale, bool need_energy = true) {
  ProfScope ps_step(md->ctx, "dd_step");
  atx_dd *dd = md->dd;
  cudaStream_t st = md->ctx->stream;
  const int n = md->nown, nloc = n + md->ngl + md->ngr;
  const int *stop = &md->ctrl.ptr->stop;
  {
    ProfScope ps_(md->ctx, "dd_drift");
    if (n > 0 && scale > 0.0) {
      k_dd_kickdrift<<<(n + 255) / 256, 256, 0, st>>>(n, md->dt, scale, md->r.ptr, md->v.ptr, md->tmpd.ptr,
                                                      md->minv.ptr, md->ctrl.ptr, md->nl->inv.ptr, md->nl->pos4.ptr);
    } else if (n > 0) {
      const bool fuse = dd->nranks > 1 && md->p2p;
      k_dd_drift<<<(n + 255) / 256, 256, 0, st>>>(n, md->dt, md->r.ptr, md->v.ptr, md->f.ptr, md->minv.ptr,
                                                  md->ctrl.ptr, fuse ? md->nl->inv.ptr : nullptr,
                                                  fuse ? md->nl->pos4.ptr : nullptr);
    } else {
      k_dd_nowish<<<1, 1, 0, st>>>(md->ctrl.ptr);
    }
    ATX_LAUNCHED();
  }
  if (dd->nranks > 1 && md->p2p) {
    const double dL = 1.0 / dd->nranks, dR = -1.0 / dd->nranks;
    const size_t par_stride = 3 * md->capG;
    double *recvL = (double *)(md->mbox + DD_SIG_BYTES), *recvR = recvL + 2 * par_stride;
    // my left neighbour receives in ITS recvR, my right neighbour in ITS recvL
    double *dstL = md->left >= 0 ? (double *)(md->peer_mbox[md->left] + DD_SIG_BYTES) + 2 * par_stride : nullptr;
    double *dstR = md->right >= 0 ? (double *)(md->peer_mbox[md->right] + DD_SIG_BYTES) : nullptr;
    const int nsend = md->nsendL + md->nsendR;
    {
      ProfScope ps_(md->ctx, "dd_halo");
      k_dd_pack_p2p<<<nsend > 0 ? (nsend + 255) / 256 : 1, 256, 0, st>>>(
          md->nsendL, md->nsendR, md->sendL.ptr, md->sendR.ptr, md->r.ptr, dL * md->a1[0], dL * md->a1[1],
          dL * md->a1[2], dR * md->a1[0], dR * md->a1[1], dR * md->a1[2], dstL, dstR, par_stride,
          md->d_peer_sig.ptr, dd->rank, dd->nranks, md->ctrl.ptr);
      ATX_LAUNCHED();
      // interior centres do not depend on this step's ghost positions: they run while the halo travels
      if (md->split) ATX_PASS(dd_compute(md, true, 1, need_energy));
      k_dd_wait<<<1, 32, 0, st>>>((const unsigned long long *)md->mbox, dd->nranks, md->ctrl.ptr, scale > 0.0 ? 1 : 0);
      ATX_LAUNCHED();
    }
    if (nloc > n) {
      ProfScope ps_(md->ctx, "dd_refresh");
      k_dd_refresh_p2p<<<(nloc - n + 255) / 256, 256, 0, st>>>(nloc - n, n, md->ngl, recvL, recvR, par_stride,
                                                               md->nl->inv.ptr, md->nl->pos4.ptr, md->ctrl.ptr);
      ATX_LAUNCHED();
    }
    ATX_PASS(dd_compute(md, true, md->split ? 2 : 0, need_energy));
    if (scale == 0.0) ATX_PASS(dd_kick(md));
    return 0;
  }
  if (dd->nranks > 1) {
    // global OR of the rebuild wish, written straight into the stop flag (no host involvement)
    {
      ProfScope ps_(md->ctx, "dd_allreduce");
      ATX_NCCL(g_nccl.AllReduce(&md->ctrl.ptr->want, &md->ctrl.ptr->stop, 1, ncclInt, ncclMax, dd->comm, st));
    }
    // ghost positions: my frame -> global (+ periodic wrap) -> receiver frame.  The local origins
    // of neighbouring slabs differ by a1/P once the wrap is included, so the whole frame change is
    // a constant shift applied by the sender.
    const double dL = 1.0 / dd->nranks;    // to the left slab
    const double dR = -1.0 / dd->nranks;   // to the right slab
    if (md->nsendL > 0) {
      k_dd_pack<<<(md->nsendL + 127) / 128, 128, 0, st>>>(md->nsendL, md->sendL.ptr, md->r.ptr, dL * md->a1[0],
                                                          dL * md->a1[1], dL * md->a1[2], md->bufL.ptr, stop);
      ATX_LAUNCHED();
    }
    if (md->nsendR > 0) {
      k_dd_pack<<<(md->nsendR + 127) / 128, 128, 0, st>>>(md->nsendR, md->sendR.ptr, md->r.ptr, dR * md->a1[0],
                                                          dR * md->a1[1], dR * md->a1[2], md->bufR.ptr, stop);
      ATX_LAUNCHED();
    }
    {
      ProfScope ps_(md->ctx, "dd_halo");
      ATX_PASS(dd_sendrecv(md, md->bufL.ptr, 3 * (size_t)md->nsendL, md->bufR.ptr, 3 * (size_t)md->nsendR,
                           md->r.ptr + 3 * (size_t)n, 3 * (size_t)md->ngl,
                           md->r.ptr + 3 * (size_t)(n + md->ngl), 3 * (size_t)md->ngr));
    }
  } else {
    // single rank: want -> stop
    ATX_CUDA(cudaMemcpyAsync(&md->ctrl.ptr->stop, &md->ctrl.ptr->want, sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  if (nloc > 0) {
    k_dd_refresh<<<(nloc + 255) / 256, 256, 0, st>>>(nloc, md->r.ptr, md->nl->order.ptr, md->nl->pos4.ptr, stop);
    ATX_LAUNCHED();
  }
  ATX_PASS(dd_compute(md, true));
  ATX_PASS(dd_kick(md));
  return 0;
}

extern "C" int atx_dd_md_run(atx_ddmd *md, int nsteps, double *epot, double *ekin) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  atx_dd *dd = md->dd;
  cudaStream_t st = md->ctx->stream;
  ATX_CUDA(cudaEventRecord(md->ev0, st));
  md->hctrl.ptr->steps_done = 0;
  ATX_CUDA(cudaMemcpyAsync(&md->ctrl.ptr->steps_done, &md->hctrl.ptr->steps_done, sizeof(int),
                           cudaMemcpyHostToDevice, st));
  int remaining = nsteps, done_total = 0;
  // leapfrog form of the step (one kick-and-drift kernel, no per-step kick / energy reduction) on the
  // peer-to-peer path; ATX_DD_FUSED=0 keeps the separate kernels
  static const bool fused_ok = !(getenv("ATX_DD_FUSED") && atoi(getenv("ATX_DD_FUSED")) == 0);
  const bool fused = fused_ok && dd->nranks > 1 && md->p2p && nsteps > 0;
  bool first = true;
  while (remaining > 0) {
    int batch = remaining < md->batch ? remaining : md->batch;
    for (int b = 0; b < batch; b++) {
      // leapfrog form: only the evaluation that can be the last one of the run needs per-atom energies
      ATX_PASS(dd_enqueue_step(md, fused ? (first ? 0.5 : 1.0) : 0.0, !fused || (b == batch - 1 && remaining == batch)));
      first = false;
    }
    ATX_CUDA(cudaMemcpyAsync(md->hctrl.ptr, md->ctrl.ptr, sizeof(DdCtrl), cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    ATX_CUDA(cudaGetLastError());
    if (md->pot_kind == ATX_POT_BOP) ATX_PASS(atx_bop_check_overflow((atx_bop *)md->pot));
    if (md->pot_kind == ATX_POT_REBO2) ATX_PASS(atx_rebo2_check_overflow((atx_rebo2 *)md->pot));
    DdCtrl hc = *md->hctrl.ptr;
    if (hc.err) {
      atx_set_error("Domain decomposition: a peer's halo signal did not arrive (peer-to-peer step path timed out).");
      return ATX_ERROR_MPI;
    }
    int done = hc.steps_done - done_total;
    done_total = hc.steps_done;
    remaining -= done;
    if (hc.stop) {
      // the stop flag is global: every rank arrives here after the same step
      ATX_PASS(dd_rebuild(md));
      ATX_PASS(dd_build_list(md));
      if (fused) md->hctrl.ptr->steps_done += 1;   // no kick kernel counts the step the rebuild completes
      ATX_PASS(dd_reset_ctrl(md));
      {
        DdTimer tm(md);
        ATX_PASS(dd_compute(md, false));
        if (!fused) ATX_PASS(dd_kick(md));
        tm.lap(7);
      }
      done_total += 1;
      remaining -= 1;
    }
  }
  if (fused) ATX_PASS(dd_kick(md, 0));   // whole-step velocities, local force array, energy sums
  ATX_CUDA(cudaMemcpyAsync(md->hctrl.ptr, md->ctrl.ptr, sizeof(DdCtrl), cudaMemcpyDeviceToHost, st));
  // global energies
  ATX_PASS(md->sums.reserve(ATX_NSUM));
  if (dd->nranks > 1) {
    ATX_NCCL(g_nccl.AllReduce(&md->ctrl.ptr->epot, md->sums.ptr, 2, ncclDouble, ncclSum, dd->comm, st));
  } else {
    ATX_CUDA(cudaMemcpyAsync(md->sums.ptr, &md->ctrl.ptr->epot, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  ATX_CUDA(cudaMemcpyAsync(md->stage.ptr, md->sums.ptr, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaEventRecord(md->ev1, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ATX_CUDA(cudaGetLastError());
  float ms = 0.f;
  ATX_CUDA(cudaEventElapsedTime(&ms, md->ev0, md->ev1));
  md->last_ms = ms;
  if (epot) *epot = md->stage.ptr[0];
  if (ekin) *ekin = md->stage.ptr[1];
  return 0;
}

extern "C" int atx_dd_md_get_count(atx_ddmd *md, int *nown, int *nghost) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  if (nown) *nown = md->nown;
  if (nghost) *nghost = md->ngl + md->ngr;
  return 0;
}

extern "C" int atx_dd_md_get_state(atx_ddmd *md, long long *id, double *r, double *v, double *f) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  cudaStream_t st = md->ctx->stream;
  int n = md->nown;
  ATX_PASS(md->stage.reserve(10 * (size_t)n + 16));
  double *h = md->stage.ptr;
  if (n > 0) {
    ATX_CUDA(cudaMemcpyAsync(h, md->r.ptr, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaMemcpyAsync(h + 3 * (size_t)n, md->v.ptr, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaMemcpyAsync(h + 6 * (size_t)n, md->f.ptr, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaMemcpyAsync(h + 9 * (size_t)n, md->idd.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  }
  ATX_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < n; i++) {
    if (id) id[i] = (long long)h[9 * (size_t)n + i];
    for (int c = 0; c < 3; c++) {
      if (r) r[3 * i + c] = h[3 * i + c] + md->torig[c];
      if (v) v[3 * i + c] = h[3 * (size_t)n + 3 * i + c];
      if (f) f[3 * i + c] = h[6 * (size_t)n + 3 * i + c];
    }
  }
  return 0;
}

extern "C" int atx_dd_md_get_stats(atx_ddmd *md, long long *nrebuilds, double *last_run_ms) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  if (nrebuilds) *nrebuilds = md->nrebuilds;
  if (last_run_ms) *last_run_ms = md->last_ms;
  return 0;
}

// host wall time (ms, accumulated since create) of the rebuild phases: 0 migration select + counts,
// 1 count exchange, 2 migration transfer + compaction, 3 ghost select + counts, 4 count exchange,
// 5 ghost transfer, 6 local neighbour list + roles, 7 first force evaluation.  Exact per phase only
// with ATX_DD_PROFILE=1 (a stream synchronisation at every phase boundary); p2p: 1 when the
// peer-to-peer step path is active.
extern "C" int atx_dd_md_get_profile(atx_ddmd *md, double *ms8, int *p2p) {
  if (!md) return ATX_ERROR_UNSPECIFIED;
  if (ms8) for (int k = 0; k < 8; k++) ms8[k] = md->rebuild_host_ms[k];
  if (p2p) *p2p = md->p2p ? 1 : 0;
  return 0;
}
