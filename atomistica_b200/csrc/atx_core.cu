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
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int atx_ctx_synchronize(atx_ctx *c) {
  ATX_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// particles
// ---------------------------------------------------------------------------

extern "C" int atx_particles_create(atx_ctx *ctx, atx_particles **p) {
  if (!ctx || !p) return ATX_ERROR_UNSPECIFIED;
  *p = new atx_particles();
  (*p)->ctx = ctx;
  return 0;
}

extern "C" int atx_particles_destroy(atx_particles *p) {
  delete p;
  return 0;
}

extern "C" int atx_particles_set_cell(atx_particles *p, const double *Abox, const double *Bbox,
                                      const int *pbc) {
  for (int i = 0; i < 9; i++) {
    p->Abox.m[i] = Abox[i];
    p->Bbox.m[i] = Bbox[i];
  }
  for (int i = 0; i < 3; i++) p->pbc[i] = pbc[i] != 0;
  p->cell_rev++;
  return 0;
}

extern "C" int atx_particles_set_positions(atx_particles *p, int nat, const double *r) {
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
  p->nat = nat;
  p->r_ext = r_dev;
  p->pos_rev++;
  return 0;
}

extern "C" int atx_particles_set_elements(atx_particles *p, int nat, const int *el) {
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
