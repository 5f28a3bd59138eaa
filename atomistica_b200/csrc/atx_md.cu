// Device-resident NVE driver: velocity-Verlet + Verlet-shell neighbour maintenance.
//
// Behavioural model:
//   verlet_step1  src/standalone/verlet.f90:100-177   v += f/(2m) dt; r += v dt; accum_max_dr += max|dr|
//   refresh rule  src/standalone/neighbors.f90:552-590 rebuild iff 2*accum_max_dr >= verlet_shell,
//                                                      accum_max_dr reset to 1d-6 on rebuild
//   verlet_step2  src/standalone/verlet.f90:183-235   v += f/(2m) dt
//
// B200 design: positions live in the neighbour list's cell-sorted 32-byte records and are advanced
// in place; velocities/forces/masses are kept in the same sorted order and re-permuted at every
// rebuild, so the per-step kernels touch only contiguous data.  Steps are enqueued optimistically
// in batches without any host synchronisation: the drift kernel evaluates the rebuild rule on the
// device and raises a stop flag, after which the remaining kernels of the batch return
// immediately; the host then rebuilds the list, finishes the interrupted step and continues.
// The rebuild schedule is therefore exactly the reference's.
#include "atx_potential_common.cuh"

// eV/(A*amu) -> A/fs^2
#define ATX_ACCEL_CONV 9.648533212331e-3

int atx_eam_compute_device(atx_eam *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o);
int atx_bop_compute_device(atx_bop *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o);
int atx_rebo2_compute_device(atx_rebo2 *pot, atx_particles *p, atx_neighbors *nl, const PotOut &o);
int atx_bop_check_overflow(atx_bop *pot);
int atx_rebo2_check_overflow(atx_rebo2 *pot);

struct MdCtrl {
  int stop;
  int steps_done;
  unsigned int counter_drift;
  unsigned int counter_kick;
  unsigned long long stepmax_bits;
  double accum_max_dr;
  double verlet_shell;
  double epot;
  double ekin;
};

struct atx_md {
  atx_ctx *ctx = nullptr;
  int pot_kind = 0;
  void *pot = nullptr;
  atx_particles *p_user = nullptr;
  atx_neighbors *nl = nullptr;
  atx_particles pint;  // particles in the driver's internal (sorted) order
  int nat = 0;
  double dt = 1.0;
  DevBuf<double> r_int, v, f, minv, tmp3, tmp1, sums, kin_partials;
  DevBuf<int> id, tmpi, el_int;
  DevBuf<MdCtrl> ctrl;
  PinBuf<MdCtrl> hctrl;
  PinBuf<double> stage;
  long long nrebuilds = 0;
  double last_ms = 0.0;
  bool forces_valid = false;
  int batch = 16;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

__global__ void k_md_drift(int nat, double dt, double4 *__restrict__ pos4, double *__restrict__ v,
                           const double *__restrict__ f, const double *__restrict__ minv,
                           MdCtrl *__restrict__ ctrl) {
  if (ctrl->stop) return;
  __shared__ double red[8];
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (s < nat) {
    double a = 0.5 * minv[s] * ATX_ACCEL_CONV * dt;
    double vx = v[3 * s] + a * f[3 * s];
    double vy = v[3 * s + 1] + a * f[3 * s + 1];
    double vz = v[3 * s + 2] + a * f[3 * s + 2];
    v[3 * s] = vx; v[3 * s + 1] = vy; v[3 * s + 2] = vz;
    double dx = vx * dt, dy = vy * dt, dz = vz * dt;
    double4 p = pos4[s];
    p.x += dx; p.y += dy; p.z += dz;
    pos4[s] = p;
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
      // last block: accum_max_dr += sqrt(max dr^2); rebuild rule
      __threadfence();
      double mx = __longlong_as_double((long long)atomicAdd(&ctrl->stepmax_bits, 0ull));
      double acc = ctrl->accum_max_dr + sqrt(mx);
      ctrl->accum_max_dr = acc;
      ctrl->stepmax_bits = 0ull;
      ctrl->counter_drift = 0u;
      if (2.0 * acc >= ctrl->verlet_shell) ctrl->stop = 1;
    }
  }
}

// Velocity-Verlet in leapfrog form for the steps inside a run (verlet.f90:100-235 applies two half
// kicks with the same force around the force evaluation; inside a run they are one update): scale = 1,
// or 0.5 for the first step of a run, whose state has whole-step velocities.  The last block applies the
// rebuild rule; a step that is not stopped here will complete, so it is counted here.  The
// whole-step velocities and the energies are produced once, by k_md_kick at the end of the run.
__global__ void k_md_kickdrift(int nat, double dt, double scale, double4 *__restrict__ pos4, double *__restrict__ v,
                               const double *__restrict__ f, const double *__restrict__ minv,
                               MdCtrl *__restrict__ ctrl) {
  if (ctrl->stop) return;
  __shared__ double red[8];
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (s < nat) {
    double a = scale * minv[s] * ATX_ACCEL_CONV * dt;
    double vx = v[3 * s] + a * f[3 * s];
    double vy = v[3 * s + 1] + a * f[3 * s + 1];
    double vz = v[3 * s + 2] + a * f[3 * s + 2];
    v[3 * s] = vx; v[3 * s + 1] = vy; v[3 * s + 2] = vz;
    double dx = vx * dt, dy = vy * dt, dz = vz * dt;
    double4 p = pos4[s];
    p.x += dx; p.y += dy; p.z += dz;
    pos4[s] = p;
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
      if (2.0 * acc >= ctrl->verlet_shell) ctrl->stop = 1;
      else ctrl->steps_done += 1;
    }
  }
}

__global__ void k_md_kick(int nat, double dt, double *__restrict__ v, const double *__restrict__ f,
                          const double *__restrict__ minv, const double *__restrict__ sums,
                          double *__restrict__ kin_partials, MdCtrl *__restrict__ ctrl, int count_step) {
  if (ctrl->stop) return;
  __shared__ double red[8];
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double ek = 0.0;
  if (s < nat) {
    double mi = minv[s];
    double a = 0.5 * mi * ATX_ACCEL_CONV * dt;
    double vx = v[3 * s] + a * f[3 * s];
    double vy = v[3 * s + 1] + a * f[3 * s + 1];
    double vz = v[3 * s + 2] + a * f[3 * s + 2];
    v[3 * s] = vx; v[3 * s + 1] = vy; v[3 * s + 2] = vz;
    ek = 0.5 * (vx * vx + vy * vy + vz * vz) / (mi * ATX_ACCEL_CONV);
  }
  for (int o = 16; o > 0; o >>= 1) ek += __shfl_xor_sync(0xffffffffu, ek, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ek;
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    kin_partials[blockIdx.x] = t;
    __threadfence();
    unsigned int done = atomicAdd(&ctrl->counter_kick, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    // last block: deterministic tree sum of the per-block kinetic energies
    __threadfence();
    double tot = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      tot += ((volatile double *)kin_partials)[b];
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
      ctrl->ekin = t;
      ctrl->epot = sums[0];
      ctrl->counter_kick = 0u;
      ctrl->steps_done += count_step;
    }
  }
}

__global__ void k_md_extract(int nat, const double4 *__restrict__ pos4, double *__restrict__ r,
                             int *__restrict__ el) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  double4 p = pos4[s];
  r[3 * s] = p.x; r[3 * s + 1] = p.y; r[3 * s + 2] = p.z;
  el[s] = (int)p.w;
}

__global__ void k_md_permute(int nat, const int *__restrict__ order, const double *__restrict__ v,
                             const double *__restrict__ minv, const int *__restrict__ id,
                             double *__restrict__ v2, double *__restrict__ minv2,
                             int *__restrict__ id2) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  int o = order[s];
  v2[3 * s] = v[3 * o]; v2[3 * s + 1] = v[3 * o + 1]; v2[3 * s + 2] = v[3 * o + 2];
  minv2[s] = minv[o];
  id2[s] = id[o];
}

__global__ void k_md_iota(int n, int *a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

__global__ void k_md_gather_state(int nat, const int *__restrict__ id,
                                  const double4 *__restrict__ pos4, const double *__restrict__ v,
                                  const double *__restrict__ f, double *__restrict__ out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nat) return;
  int i = id[s];
  double4 p = pos4[s];
  double *r_o = out, *v_o = out + 3 * (size_t)nat, *f_o = out + 6 * (size_t)nat;
  r_o[3 * i] = p.x; r_o[3 * i + 1] = p.y; r_o[3 * i + 2] = p.z;
  for (int c = 0; c < 3; c++) {
    v_o[3 * i + c] = v[3 * s + c];
    f_o[3 * i + c] = f[3 * s + c];
  }
}

static int md_compute(atx_md *md, bool guarded, bool want_sums = true) {
  PotOut o;
  o.f = md->f.ptr;
  o.sums = md->sums.ptr;
  o.stop = guarded ? &md->ctrl.ptr->stop : nullptr;
  o.want_virial = false;
  o.want_sums = want_sums;
  switch (md->pot_kind) {
    case ATX_POT_EAM:
      return atx_eam_compute_device((atx_eam *)md->pot, &md->pint, md->nl, nullptr, o);
    case ATX_POT_BOP:
      return atx_bop_compute_device((atx_bop *)md->pot, &md->pint, md->nl, nullptr, o);
    case ATX_POT_REBO2:
      return atx_rebo2_compute_device((atx_rebo2 *)md->pot, &md->pint, md->nl, o);
  }
  atx_set_error("atx_md: unknown potential kind");
  return ATX_ERROR_UNSPECIFIED;
}

// full neighbour rebuild from the current sorted positions, then re-permute the state
static int md_rebuild(atx_md *md, bool first) {
  atx_ctx *ctx = md->ctx;
  cudaStream_t st = ctx->stream;
  int nat = md->nat, gb = (nat + 255) / 256;
  atx_neighbors *nl = md->nl;
  if (!first) {
    k_md_extract<<<gb, 256, 0, st>>>(nat, nl->pos4.ptr, md->r_int.ptr, md->pint.el.ptr);
    ATX_LAUNCHED();
  }
  md->pint.pos_rev++;
  // the driver's rule (accumulated per-step maxima) has fired: a REAL rebuild, not the library-mode
  // displacement check of atx_neighbors_update, which would keep the list (and its build positions)
  // whenever the true displacement is still below shell/2
  nl->p_rev = -1;
  ATX_PASS(atx_neighbors_update(nl, &md->pint));
  ATX_PASS(md->tmp3.reserve(3 * (size_t)nat + 3));
  ATX_PASS(md->tmp1.reserve(nat + 1));
  ATX_PASS(md->tmpi.reserve(nat + 1));
  k_md_permute<<<gb, 256, 0, st>>>(nat, nl->order.ptr, md->v.ptr, md->minv.ptr, md->id.ptr,
                                   md->tmp3.ptr, md->tmp1.ptr, md->tmpi.ptr);
  ATX_LAUNCHED();
  std::swap(md->v.ptr, md->tmp3.ptr); std::swap(md->v.cap, md->tmp3.cap);
  std::swap(md->minv.ptr, md->tmp1.ptr); std::swap(md->minv.cap, md->tmp1.cap);
  std::swap(md->id.ptr, md->tmpi.ptr); std::swap(md->id.cap, md->tmpi.cap);
  // internal order == sorted order from now on
  k_md_iota<<<gb, 256, 0, st>>>(nat, nl->order.ptr);
  ATX_LAUNCHED();
  k_md_iota<<<gb, 256, 0, st>>>(nat, nl->inv.ptr);
  ATX_LAUNCHED();
  md->nrebuilds++;
  return 0;
}

extern "C" int atx_md_create(atx_ctx *ctx, int pot_kind, void *pot, atx_particles *p,
                             atx_neighbors *nl, const double *mass, const double *v, double dt,
                             atx_md **out) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (!ctx || !pot || !p || !nl || !mass || !out) return ATX_ERROR_UNSPECIFIED;
  int nat = p->nat;
  if (nat <= 0) {
    atx_set_error("atx_md_create: particles hold no positions.");
    return ATX_ERROR_UNSPECIFIED;
  }
  atx_md *md = new atx_md();
  md->ctx = ctx;
  md->pot_kind = pot_kind;
  md->pot = pot;
  md->p_user = p;
  md->nl = nl;
  md->nat = nat;
  md->dt = dt;
  cudaStream_t st = ctx->stream;
  int gb = (nat + 255) / 256;
  ATX_PASS(md->r_int.reserve(3 * (size_t)nat + 3));
  ATX_PASS(md->v.reserve(3 * (size_t)nat + 3));
  ATX_PASS(md->f.reserve(3 * (size_t)nat + 3));
  ATX_PASS(md->minv.reserve(nat + 1));
  ATX_PASS(md->id.reserve(nat + 1));
  ATX_PASS(md->sums.reserve(ATX_NSUM));
  ATX_PASS(md->kin_partials.reserve(gb + 1));
  ATX_PASS(md->ctrl.reserve(1));
  ATX_PASS(md->hctrl.reserve(1));
  ATX_PASS(md->stage.reserve(9 * (size_t)nat + 16));
  // internal particles: same cell, positions/elements copied in original order
  md->pint.ctx = ctx;
  md->pint.nat = nat;
  md->pint.Abox = p->Abox;
  md->pint.Bbox = p->Bbox;
  for (int k = 0; k < 3; k++) md->pint.pbc[k] = p->pbc[k];
  md->pint.cell_rev = 1;
  ATX_PASS(md->pint.el.reserve(nat + 1));
  ATX_CUDA(cudaMemcpyAsync(md->r_int.ptr, p->rptr(), sizeof(double) * 3 * nat,
                           cudaMemcpyDeviceToDevice, st));
  if (p->el.cap) {
    ATX_CUDA(cudaMemcpyAsync(md->pint.el.ptr, p->el.ptr, sizeof(int) * nat, cudaMemcpyDeviceToDevice, st));
  } else {
    atx_set_error("atx_md_create: particle elements have not been set.");
    delete md;
    return ATX_ERROR_UNSPECIFIED;
  }
  md->pint.r_ext = md->r_int.ptr;
  double *h = md->stage.ptr;
  for (int i = 0; i < nat; i++) h[i] = 1.0 / mass[i];
  ATX_CUDA(cudaMemcpyAsync(md->minv.ptr, h, sizeof(double) * nat, cudaMemcpyHostToDevice, st));
  if (v) {
    for (size_t i = 0; i < 3 * (size_t)nat; i++) h[nat + i] = v[i];
    ATX_CUDA(cudaMemcpyAsync(md->v.ptr, h + nat, sizeof(double) * 3 * nat, cudaMemcpyHostToDevice, st));
  } else {
    ATX_CUDA(cudaMemsetAsync(md->v.ptr, 0, sizeof(double) * 3 * nat, st));
  }
  k_md_iota<<<gb, 256, 0, st>>>(nat, md->id.ptr);
  ATX_LAUNCHED();
  MdCtrl c{};
  c.accum_max_dr = 1e-6;
  c.verlet_shell = nl->verlet_shell;
  *md->hctrl.ptr = c;
  ATX_CUDA(cudaMemcpyAsync(md->ctrl.ptr, md->hctrl.ptr, sizeof(MdCtrl), cudaMemcpyHostToDevice, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ATX_CUDA(cudaEventCreate(&md->ev0));
  ATX_CUDA(cudaEventCreate(&md->ev1));
  int err = md_rebuild(md, true);
  if (!err) err = md_compute(md, false);
  if (err) {
    delete md;
    return err;
  }
  ATX_CUDA(cudaStreamSynchronize(st));
  md->forces_valid = true;
  *out = md;
  return 0;
}

extern "C" int atx_md_destroy(atx_md *md) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  if (!md) return 0;
  if (md->ev0) cudaEventDestroy(md->ev0);
  if (md->ev1) cudaEventDestroy(md->ev1);
  md->pint.r_ext = nullptr;
  delete md;
  return 0;
}

static int md_reset_ctrl_after_rebuild(atx_md *md) {
  // stop = 0, accum_max_dr = 1d-6 (standalone/neighbors.f90:575)
  MdCtrl *h = md->hctrl.ptr;
  h->stop = 0;
  h->accum_max_dr = 1e-6;
  h->stepmax_bits = 0;
  h->counter_drift = 0;
  h->counter_kick = 0;
  ATX_CUDA(cudaMemcpyAsync(md->ctrl.ptr, h, sizeof(MdCtrl), cudaMemcpyHostToDevice, md->ctx->stream));
  return 0;
}

extern "C" int atx_md_run(atx_md *md, int nsteps, double *epot, double *ekin) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  atx_ctx *ctx = md->ctx;
  cudaStream_t st = ctx->stream;
  int nat = md->nat, gb = (nat + 255) / 256;
  atx_neighbors *nl = md->nl;
  ATX_CUDA(cudaEventRecord(md->ev0, st));
  int remaining = nsteps;
  // steps_done counts from zero in every run
  md->hctrl.ptr->steps_done = 0;
  {
    MdCtrl *h = md->hctrl.ptr;
    ATX_CUDA(cudaMemcpyAsync(&md->ctrl.ptr->steps_done, &h->steps_done, sizeof(int),
                             cudaMemcpyHostToDevice, st));
  }
  int done_total = 0;
  // leapfrog form of the step (one kick-and-drift kernel per step, final half kick and energy sums once
  // per run); ATX_MD_FUSED=0 keeps the separate drift and kick kernels
  static const bool fused_ok = !(getenv("ATX_MD_FUSED") && atoi(getenv("ATX_MD_FUSED")) == 0);
  const bool fused = fused_ok && nsteps > 0 && nat > 0;
  bool first = true;
  while (remaining > 0) {
    int batch = remaining < md->batch ? remaining : md->batch;
    for (int b = 0; b < batch; b++) {
      if (fused) {
        k_md_kickdrift<<<gb, 256, 0, st>>>(nat, md->dt, first ? 0.5 : 1.0, nl->pos4.ptr, md->v.ptr, md->f.ptr,
                                           md->minv.ptr, md->ctrl.ptr);
        ATX_LAUNCHED();
        first = false;
        // the potential energy is read once, by the kick that ends the run: only the evaluation that can
        // be the last one reduces its partial sums (a rebuild re-evaluates with sums)
        ATX_PASS(md_compute(md, true, b == batch - 1 && remaining == batch));
        continue;
      }
      k_md_drift<<<gb, 256, 0, st>>>(nat, md->dt, nl->pos4.ptr, md->v.ptr, md->f.ptr, md->minv.ptr,
                                     md->ctrl.ptr);
      ATX_LAUNCHED();
      ATX_PASS(md_compute(md, true));
      k_md_kick<<<gb, 256, 0, st>>>(nat, md->dt, md->v.ptr, md->f.ptr, md->minv.ptr, md->sums.ptr,
                                    md->kin_partials.ptr, md->ctrl.ptr, 1);
      ATX_LAUNCHED();
    }
    ATX_CUDA(cudaMemcpyAsync(md->hctrl.ptr, md->ctrl.ptr, sizeof(MdCtrl), cudaMemcpyDeviceToHost, st));
    ATX_CUDA(cudaStreamSynchronize(st));
    ATX_CUDA(cudaGetLastError());
    if (md->pot_kind == ATX_POT_BOP) ATX_PASS(atx_bop_check_overflow((atx_bop *)md->pot));
    if (md->pot_kind == ATX_POT_REBO2) ATX_PASS(atx_rebo2_check_overflow((atx_rebo2 *)md->pot));
    MdCtrl hc = *md->hctrl.ptr;
    int done = hc.steps_done - done_total;
    done_total = hc.steps_done;
    remaining -= done;
    if (hc.stop) {
      // a drift tripped the rebuild rule: rebuild, then finish that step
      ATX_PASS(md_rebuild(md, false));
      if (fused) md->hctrl.ptr->steps_done += 1;   // no kick kernel counts the step the rebuild completes
      ATX_PASS(md_reset_ctrl_after_rebuild(md));
      ATX_PASS(md_compute(md, false));
      if (!fused) {
        k_md_kick<<<gb, 256, 0, st>>>(nat, md->dt, md->v.ptr, md->f.ptr, md->minv.ptr, md->sums.ptr,
                                      md->kin_partials.ptr, md->ctrl.ptr, 1);
        ATX_LAUNCHED();
      }
      done_total += 1;
      remaining -= 1;
    }
  }
  if (fused) {
    // whole-step velocities and the energy sums of the last force evaluation
    k_md_kick<<<gb, 256, 0, st>>>(nat, md->dt, md->v.ptr, md->f.ptr, md->minv.ptr, md->sums.ptr,
                                  md->kin_partials.ptr, md->ctrl.ptr, 0);
    ATX_LAUNCHED();
  }
  ATX_CUDA(cudaMemcpyAsync(md->hctrl.ptr, md->ctrl.ptr, sizeof(MdCtrl), cudaMemcpyDeviceToHost, st));
  ATX_CUDA(cudaEventRecord(md->ev1, st));
  ATX_CUDA(cudaStreamSynchronize(st));
  ATX_CUDA(cudaGetLastError());
  float ms = 0.f;
  ATX_CUDA(cudaEventElapsedTime(&ms, md->ev0, md->ev1));
  md->last_ms = ms;
  if (epot) *epot = md->hctrl.ptr->epot;
  if (ekin) *ekin = md->hctrl.ptr->ekin;
  return 0;
}

extern "C" int atx_md_get_state(atx_md *md, double *r, double *v, double *f) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  atx_ctx *ctx = md->ctx;
  int nat = md->nat;
  ATX_PASS(md->tmp3.reserve(9 * (size_t)nat + 9));
  k_md_gather_state<<<(nat + 255) / 256, 256, 0, ctx->stream>>>(nat, md->id.ptr, md->nl->pos4.ptr,
                                                                md->v.ptr, md->f.ptr, md->tmp3.ptr);
  ATX_LAUNCHED();
  ATX_PASS(md->stage.reserve(9 * (size_t)nat + 16));
  ATX_CUDA(cudaMemcpyAsync(md->stage.ptr, md->tmp3.ptr, sizeof(double) * 9 * nat,
                           cudaMemcpyDeviceToHost, ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(ctx->stream));
  const double *h = md->stage.ptr;
  size_t n3 = 3 * (size_t)nat;
  if (r) memcpy(r, h, sizeof(double) * n3);
  if (v) memcpy(v, h + n3, sizeof(double) * n3);
  if (f) memcpy(f, h + 2 * n3, sizeof(double) * n3);
  return 0;
}

extern "C" int atx_md_get_stats(atx_md *md, long long *nrebuilds, double *last_run_ms) {
  if (md && md->ctx) cudaSetDevice(md->ctx->device);
  if (nrebuilds) *nrebuilds = md->nrebuilds;
  if (last_run_ms) *last_run_ms = md->last_ms;
  return 0;
}
