// Context, error stack, particles_t device mirror, scan helpers.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstring>
#include <mutex>

#include "atx_internal.cuh"

static thread_local std::string g_last_error;
long long g_atx_launches = 0;

void atx_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" int atx_last_error(char *buf, int len) {
  if (!buf || len <= 0) return ATX_ERROR_UNSPECIFIED;
  std::strncpy(buf, g_last_error.c_str(), (size_t)len - 1);
  buf[len - 1] = 0;
  return 0;
}

extern "C" const char *atx_version(void) { return "atomistica_b200 0.1 (sm_100a)"; }

extern "C" long long atx_kernel_launches(int reset) {
  long long v = g_atx_launches;
  if (reset) g_atx_launches = 0;
  return v;
}

extern "C" int atx_ctx_create(int device, atx_ctx **out) {
  if (!out) return ATX_ERROR_UNSPECIFIED;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    atx_set_error(std::string("No CUDA device available (") + cudaGetErrorString(e) +
                  "); atomistica_b200 has no CPU fallback.");
    return ATX_ERROR_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    atx_set_error("Invalid CUDA device ordinal " + std::to_string(device));
    return ATX_ERROR_DEVICE;
  }
  ATX_CUDA(cudaSetDevice(device));
  atx_ctx *c = new atx_ctx();
  c->device = device;
  ATX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  ATX_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  *out = c;
  return 0;
}

extern "C" int atx_ctx_destroy(atx_ctx *c) {
  if (c) cudaSetDevice(c->device);  // entry points do not assume the caller kept the device current
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int atx_ctx_synchronize(atx_ctx *c) {
  if (c) cudaSetDevice(c->device);  // entry points do not assume the caller kept the device current
  ATX_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// particles
// ---------------------------------------------------------------------------

extern "C" int atx_particles_create(atx_ctx *ctx, atx_particles **p) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (!ctx || !p) return ATX_ERROR_UNSPECIFIED;
  *p = new atx_particles();
  (*p)->ctx = ctx;
  return 0;
}

extern "C" int atx_particles_destroy(atx_particles *p) {
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  delete p;
  return 0;
}

extern "C" int atx_particles_set_cell(atx_particles *p, const double *Abox, const double *Bbox,
                                      const int *pbc) {
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  for (int i = 0; i < 9; i++) {
    p->Abox.m[i] = Abox[i];
    p->Bbox.m[i] = Bbox[i];
  }
  for (int i = 0; i < 3; i++) p->pbc[i] = pbc[i] != 0;
  p->cell_rev++;
  return 0;
}

extern "C" int atx_particles_set_positions(atx_particles *p, int nat, const double *r) {
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  if (nat < 0) return ATX_ERROR_UNSPECIFIED;
  p->nat = nat;
  p->r_ext = nullptr;
  ATX_PASS(p->r.reserve((size_t)3 * nat + 3));
  if (nat > 0) {
    ATX_CUDA(cudaMemcpyAsync(p->r.ptr, r, sizeof(double) * 3 * nat, cudaMemcpyHostToDevice,
                             p->ctx->stream));
    // the host buffer may be pageable and reused by the caller right after we return
    ATX_CUDA(cudaStreamSynchronize(p->ctx->stream));
  }
  p->pos_rev++;
  return 0;
}

extern "C" int atx_particles_set_positions_device(atx_particles *p, int nat, const double *r_dev) {
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  p->nat = nat;
  p->r_ext = r_dev;
  p->pos_rev++;
  return 0;
}

extern "C" int atx_particles_set_elements(atx_particles *p, int nat, const int *el) {
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  ATX_PASS(p->el.reserve((size_t)nat + 1));
  if (nat > 0) {
    ATX_CUDA(cudaMemcpyAsync(p->el.ptr, el, sizeof(int) * nat, cudaMemcpyHostToDevice,
                             p->ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(p->ctx->stream));
  }
  p->el_rev++;
  p->pos_rev++;  // element ids travel with the sorted positions
  return 0;
}

extern "C" int atx_host_alloc_pinned(size_t bytes, void **ptr) {
  if (!ptr) return ATX_ERROR_UNSPECIFIED;
  ATX_CUDA(cudaMallocHost(ptr, bytes ? bytes : 8));
  return 0;
}

extern "C" int atx_host_free_pinned(void *ptr) {
  if (ptr) ATX_CUDA(cudaFreeHost(ptr));
  return 0;
}

extern "C" int atx_host_register(void *ptr, size_t bytes, int *already) {
  if (already) *already = 0;
  if (!ptr || !bytes) return ATX_ERROR_UNSPECIFIED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    atx_set_error("No CUDA device available: host memory cannot be page-locked (no CPU fallback).");
    return ATX_ERROR_UNSPECIFIED;
  }
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) {
    if (already) *already = 1;   // cudaMallocHost / cudaHostAlloc memory, or registered by somebody else
    return 0;
  }
  cudaGetLastError();
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    if (already) *already = 1;
    return 0;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    atx_set_error(std::string("cudaHostRegister failed: ") + cudaGetErrorString(e));
    return ATX_ERROR_UNSPECIFIED;
  }
  return 0;
}

extern "C" int atx_host_unregister(void *ptr) {
  if (ptr && cudaHostUnregister(ptr) != cudaSuccess) cudaGetLastError();
  return 0;
}

// ---------------------------------------------------------------------------
// scans
// ---------------------------------------------------------------------------

struct IntToLL {
  __host__ __device__ long long operator()(int v) const { return (long long)v; }
};

int atx_scan_int_to_ll(atx_ctx *ctx, const int *in, long long *out, size_t n) {
  size_t bytes = 0;
  cub::TransformInputIterator<long long, IntToLL, const int *> it(in, IntToLL());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, ctx->stream);
  ATX_PASS(ctx->cub_tmp.reserve(bytes));
  ATX_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, bytes, it, out, (int)n, ctx->stream));
  g_atx_launches += 2;
  return 0;
}

int atx_scan_int(atx_ctx *ctx, const int *in, int *out, size_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream);
  ATX_PASS(ctx->cub_tmp.reserve(bytes));
  ATX_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, bytes, in, out, (int)n, ctx->stream));
  g_atx_launches += 2;
  return 0;
}

int atx_accumulate_to_host(atx_ctx *ctx, const double *dev, double *host, size_t n,
                           PinBuf<double> &stage) {
  if (n == 0) return 0;
  ATX_PASS(stage.reserve(n));
  ATX_CUDA(cudaMemcpyAsync(stage.ptr, dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; i++) host[i] += stage.ptr[i];
  return 0;
}

// ---------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------

ProfScope::ProfScope(atx_ctx *c, const char *name) : ctx(c) {
  if (!c->prof_on) return;
  ProfSlot *slot = nullptr;
  for (auto &s : c->prof)
    if (s.name == name) slot = &s;
  if (!slot) {
    c->prof.push_back(ProfSlot());
    slot = &c->prof.back();
    slot->name = name;
  }
  if (slot->used + 2 > slot->ev.size()) {
    size_t old = slot->ev.size();
    slot->ev.resize(old + 256);
    for (size_t i = old; i < slot->ev.size(); i++) cudaEventCreate(&slot->ev[i]);
  }
  cudaEventRecord(slot->ev[slot->used], c->stream);
  stop = slot->ev[slot->used + 1];
  slot->used += 2;
}

ProfScope::~ProfScope() {
  if (stop) cudaEventRecord(stop, ctx->stream);
}

extern "C" int atx_profile_enable(atx_ctx *ctx, int on) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  ctx->prof_on = on != 0;
  if (on)
    for (auto &s : ctx->prof) s.used = 0;
  return 0;
}

extern "C" int atx_profile_read(atx_ctx *ctx, const char *name, double *total_ms, long long *count) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  ATX_CUDA(cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  long long n = 0;
  for (auto &s : ctx->prof)
    if (s.name == name)
      for (size_t i = 0; i + 1 < s.used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.ev[i], s.ev[i + 1]) == cudaSuccess) {
          tot += ms;
          n++;
        }
      }
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return 0;
}

__global__ void k_fp64_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;
}

extern "C" int atx_measure_fp64_peak(atx_ctx *ctx, double *tflops) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  DevBuf<double> out;
  ATX_PASS(out.reserve(8));
  cudaEvent_t e0, e1;
  ATX_CUDA(cudaEventCreate(&e0));
  ATX_CUDA(cudaEventCreate(&e1));
  const int iters = 8192, blocks = ctx->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    ATX_CUDA(cudaEventRecord(e0, ctx->stream));
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(out.ptr, iters);
    ATX_CUDA(cudaEventRecord(e1, ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    ATX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    double t = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && t > best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return 0;
}

__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}

extern "C" int atx_measure_copy_bandwidth(atx_ctx *ctx, double *gbs) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  const size_t n = (size_t)1 << 26;  // 64 Mi double2 = 1 GiB per buffer
  DevBuf<double2> a, b;
  ATX_PASS(a.reserve(n));
  ATX_PASS(b.reserve(n));
  ATX_CUDA(cudaMemsetAsync(a.ptr, 0, n * sizeof(double2), ctx->stream));
  cudaEvent_t e0, e1;
  ATX_CUDA(cudaEventCreate(&e0));
  ATX_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    ATX_CUDA(cudaEventRecord(e0, ctx->stream));
    k_copy<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>(a.ptr, b.ptr, n);
    ATX_CUDA(cudaEventRecord(e1, ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    ATX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double t = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
    if (rep > 0 && t > best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gbs = best;
  return 0;
}
