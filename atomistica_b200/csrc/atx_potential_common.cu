#include "atx_potential_common.cuh"

// partials are stored component-major: partials[comp * nblocks + block].
// One block per component, fixed summation tree -> deterministic.
__global__ void __launch_bounds__(256)
k_reduce_partials(const double *__restrict__ partials, int nblocks, double *__restrict__ sums,
                  const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[8];
  const int comp = blockIdx.x;
  const double *src = partials + (size_t)comp * nblocks;
  double x = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) x += src[b];
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    sums[comp] = t;
  }
}

int atx_reduce_partials(atx_ctx *ctx, const double *partials, int nblocks, double *sums,
                        const int *stop) {
  k_reduce_partials<<<ATX_NSUM, 256, 0, ctx->stream>>>(partials, nblocks, sums, stop);
  ATX_LAUNCHED();
  return 0;
}

__global__ void k_unsort(int nat, int ncomp, const int *__restrict__ order,
                         const double *__restrict__ in, double *__restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nat * ncomp) return;
  int s = (int)(t / ncomp), c = (int)(t % ncomp);
  out[(size_t)order[s] * ncomp + c] = in[t];
}

int atx_unsort(atx_ctx *ctx, int nat, int ncomp, const int *order, const double *in, double *out) {
  long long n = (long long)nat * ncomp;
  if (n == 0) return 0;
  k_unsort<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(nat, ncomp, order, in, out);
  ATX_LAUNCHED();
  return 0;
}

__global__ void k_sort_int(int nat, const int *__restrict__ order, const int *__restrict__ in,
                           int *__restrict__ out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nat) out[s] = in[order[s]];
}

int atx_sort_int(atx_ctx *ctx, int nat, const int *order, const int *in, int *out) {
  if (nat == 0) return 0;
  k_sort_int<<<(nat + 255) / 256, 256, 0, ctx->stream>>>(nat, order, in, out);
  ATX_LAUNCHED();
  return 0;
}

int atx_prepare_out(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, bool want_epa, bool want_wpa,
                    PotOut &o) {
  size_t nat = nl->nat;
  ATX_PASS(sc.f.reserve(3 * nat + 3));
  ATX_PASS(sc.sums.reserve(ATX_NSUM));
  o.f = sc.f.ptr;
  o.sums = sc.sums.ptr;
  if (want_epa) {
    ATX_PASS(sc.epa.reserve(nat + 1));
    o.epa = sc.epa.ptr;
  }
  if (want_wpa) {
    ATX_PASS(sc.wpa.reserve(9 * nat + 9));
    o.wpa = sc.wpa.ptr;
  }
  return 0;
}

int atx_prepare_mask(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, const int *mask_host,
                     const int **mask_sorted) {
  *mask_sorted = nullptr;
  if (!mask_host) return 0;
  size_t nat = nl->nat;
  ATX_PASS(sc.mask_in.reserve(nat + 1));
  ATX_PASS(sc.mask_sorted.reserve(nat + 1));
  ATX_CUDA(cudaMemcpyAsync(sc.mask_in.ptr, mask_host, sizeof(int) * nat, cudaMemcpyHostToDevice,
                           ctx->stream));
  ATX_PASS(atx_sort_int(ctx, (int)nat, nl->order.ptr, sc.mask_in.ptr, sc.mask_sorted.ptr));
  ATX_CUDA(cudaStreamSynchronize(ctx->stream));
  *mask_sorted = sc.mask_sorted.ptr;
  return 0;
}

int atx_finish_to_host(atx_ctx *ctx, atx_neighbors *nl, PotScratch &sc, const PotOut &o,
                       double *epot, double *f, double *wpot, double *epot_per_at,
                       double *wpot_per_at) {
  size_t nat = nl->nat;
  size_t per = 3 + (o.epa && epot_per_at ? 1 : 0) + (o.wpa && wpot_per_at ? 9 : 0);
  ATX_PASS(sc.out.reserve(per * nat + ATX_NSUM + 16));
  double *d = sc.out.ptr;
  ATX_PASS(atx_unsort(ctx, (int)nat, 3, nl->order.ptr, o.f, d));
  size_t off = 3 * nat;
  size_t off_epa = 0, off_wpa = 0;
  if (o.epa && epot_per_at) {
    off_epa = off;
    ATX_PASS(atx_unsort(ctx, (int)nat, 1, nl->order.ptr, o.epa, d + off));
    off += nat;
  }
  if (o.wpa && wpot_per_at) {
    off_wpa = off;
    ATX_PASS(atx_unsort(ctx, (int)nat, 9, nl->order.ptr, o.wpa, d + off));
    off += 9 * nat;
  }
  ATX_CUDA(cudaMemcpyAsync(d + off, o.sums, sizeof(double) * ATX_NSUM, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  size_t total = off + ATX_NSUM;
  if (sc.store_outputs) {
    // device -> caller's arrays directly (page-locked arrays make these true asynchronous DMAs)
    ATX_PASS(sc.stage_small.reserve(ATX_NSUM));
    if (f) ATX_CUDA(cudaMemcpyAsync(f, d, sizeof(double) * 3 * nat, cudaMemcpyDeviceToHost, ctx->stream));
    if (o.epa && epot_per_at)
      ATX_CUDA(cudaMemcpyAsync(epot_per_at, d + off_epa, sizeof(double) * nat, cudaMemcpyDeviceToHost, ctx->stream));
    if (o.wpa && wpot_per_at)
      ATX_CUDA(cudaMemcpyAsync(wpot_per_at, d + off_wpa, sizeof(double) * 9 * nat, cudaMemcpyDeviceToHost,
                               ctx->stream));
    ATX_CUDA(cudaMemcpyAsync(sc.stage_small.ptr, o.sums, sizeof(double) * ATX_NSUM, cudaMemcpyDeviceToHost,
                             ctx->stream));
    ATX_CUDA(cudaStreamSynchronize(ctx->stream));
    ATX_CUDA(cudaGetLastError());
    const double *hs = sc.stage_small.ptr;
    if (epot) *epot += hs[0];
    if (wpot)
      for (int k = 0; k < 9; k++) wpot[k] += hs[1 + k];
    return 0;
  }
  ATX_PASS(sc.stage.reserve(total));
  ATX_CUDA(cudaMemcpyAsync(sc.stage.ptr, d, sizeof(double) * total, cudaMemcpyDeviceToHost,
                           ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(ctx->stream));
  ATX_CUDA(cudaGetLastError());
  const double *h = sc.stage.ptr;
  if (f)
    for (size_t i = 0; i < 3 * nat; i++) f[i] += h[i];
  if (o.epa && epot_per_at)
    for (size_t i = 0; i < nat; i++) epot_per_at[i] += h[off_epa + i];
  if (o.wpa && wpot_per_at)
    for (size_t i = 0; i < 9 * nat; i++) wpot_per_at[i] += h[off_wpa + i];
  if (epot) *epot += h[off];
  if (wpot)
    for (int k = 0; k < 9; k++) wpot[k] += h[off + 1 + k];
  return 0;
}
