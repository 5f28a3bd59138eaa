// Tabulated alloy EAM on the device.
//
// Replaces tabulated_alloy_eam_energy_and_forces_kernel
// (src/potentials/eam/tabulated_alloy_eam.f90:423-627) and the spline evaluators it calls
// (src/support/simple_spline.f90:373-434 func, 467-528 dfunc, 536-614 f_and_df).
//
// Two gather passes, no atomics, deterministic:
//   k_eam_density: rho_i = sum_j rho_j(r_ij), clamp, F(rho_i), F'(rho_i)      (the reference's pass 1)
//   k_eam_force:   f_i = sum_j (c_ij + c_ji) dr_ij with
//                    c_ij = -(F'_i rho'_j(r) + (phi' - phi/r)/r)/r   (i's visit of j, :594-595)
//                    c_ji = -(F'_j rho'_i(r) + (phi' - phi/r)/r)/r   (j's visit of i, scattered
//                                                                     with "- df" in the reference)
// which is the reference's scatter form summed per receiving atom.
// A group of LANES threads serves one atom; lanes stride its list entries and reduce by shuffles.
#include <cstdlib>

#include "atx_potential_common.cuh"

struct DSpline {
  int n;
  double x0, dx, cut;
  const double4 *c;  // {y, coeff1, coeff2, coeff3} per interval
  const double4 *d;  // {dcoeff1, dcoeff2, dcoeff3, 0}
};

#define EAM_MAX_DB 10

struct EamDev {
  int ndb;
  double cutoff_sq;
  // common grid of all frho/fphi tables (setfl shares nr, dr); used by the fast kernels
  int r_n;
  double r_x0, r_inv_dx;
  int el2db[32];
  DSpline fF[EAM_MAX_DB];
  DSpline frho[EAM_MAX_DB];
  DSpline fphi[EAM_MAX_DB * EAM_MAX_DB];
  const double4 *pairrec[EAM_MAX_DB * EAM_MAX_DB];  // 64-byte records, see k_eam_force_fast
  // TabulatedEAM (funcfl): the 'phi' table holds Z(r) (scaled by sqrt(Hartree Bohr / 2)) and the
  // pair term is Z**2/r (tabulated_eam.f90:470-476): phi := Z*Z, dphi := 2 Z Z' before the shared
  // (dphi - phi/r)/r
  int zsq;
};

struct atx_eam {
  atx_ctx *ctx = nullptr;
  int ndb = 0;
  double cutoff = 0.0;
  EamDev host{};
  DevBuf<EamDev> dev;
  std::vector<DevBuf<double4> *> tables;
  DevBuf<double> dF, Fe;
  DevBuf<double4> pd4;
  DevBuf<int> flag;
  PotScratch sc;
  bool bound = false;
  bool fast_ok = false;  // all r-tables share one grid that covers the cutoff
  bool force_generic = false;
  bool funcfl = false;
  int fast_lanes = 4, fast_unroll = 2;
  int fast_map = 1;  // 1: consecutive lane -> entry mapping (A/B on a B200, round 2: force 187 vs 192 us, density 119 vs 129 us per step on C2); ATX_EAM_MAP=0 selects the strided mapping
  ~atx_eam() {
    for (auto *t : tables) delete t;
  }
};

// func (no extrapolation): sets *err when x is outside the table
__device__ __forceinline__ double spl_f(const DSpline &s, double x, int *err) {
  double xf;
  int i;
  if (x == s.cut) {
    xf = (double)s.n;
    i = s.n - 1;
  } else {
    xf = (x - s.x0) / s.dx + 1.0;
    i = (int)floor(xf);
  }
  if (i < 1 || i >= s.n) {
    *err = 1;
    return 1.0;
  }
  double B = xf - (double)i;
  double4 c = atx_ld4(&s.c[i - 1]);
  return c.x + B * (c.y + B * (c.z + B * c.w));
}

__device__ __forceinline__ double spl_df(const DSpline &s, double x, int *err) {
  double xf;
  int i;
  if (x == s.cut) {
    xf = (double)s.n;
    i = s.n - 1;
  } else {
    xf = (x - s.x0) / s.dx + 1.0;
    i = (int)floor(xf);
  }
  if (i < 1 || i >= s.n) {
    *err = 1;
    return 1.0;
  }
  double B = xf - (double)i;
  double4 d = atx_ld4(&s.d[i - 1]);
  return d.x + B * (d.y + B * d.z);
}

// f_and_df(extrapolate=.true.)
__device__ __forceinline__ void spl_f_df_x(const DSpline &s, double x, double &f, double &df) {
  double xf = (x - s.x0) / s.dx + 1.0;
  int i = (int)floor(xf);
  if (i < 1) i = 1;
  else if (i >= s.n) i = s.n - 1;
  double B = xf - (double)i;
  double4 c = atx_ld4(&s.c[i - 1]);
  double4 d = atx_ld4(&s.d[i - 1]);
  f = c.x + B * (c.y + B * (c.z + B * c.w));
  df = d.x + B * (d.y + B * d.z);
}

template <int LANES>
__global__ void __launch_bounds__(128)
k_eam_density(int nat, Mat3 A, const EamDev *__restrict__ T, const double4 *__restrict__ pos4,
              const long long *__restrict__ seed, const int2 *__restrict__ list,
              const int *__restrict__ mask, double *__restrict__ dF, double *__restrict__ Fe,
              int *__restrict__ flag, const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int gpb = 128 / LANES;
  int s = blockIdx.x * gpb + threadIdx.x / LANES;
  int lane = threadIdx.x % LANES;
  const bool valid = s < nat;
  double4 pi = valid ? pos4[s] : make_double4(0, 0, 0, 0);
  // outer ghosts (role 0) only lend their position: their list is incomplete, nobody reads their F'
  int dbi = (valid && (!role || role[s] >= 1)) ? T->el2db[(int)pi.w] : -1;
  const bool active = dbi > 0 && (!mask || mask[s] != 0);
  double rho = 0.0;
  int err = 0;
  long long b = active ? seed[s] : 0, e = active ? seed[s + 1] : 0;
  const double cutoff_sq = T->cutoff_sq;
  for (long long a = b + lane; a < e; a += LANES) {
    int2 en = list[a];
    double4 pj = pos4[en.x];
    int dbj = T->el2db[(int)pj.w];
    if (dbj <= 0) continue;
    double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    if (ATX_NONZERO_SHIFT(en.y)) {
      int sx, sy, sz;
      atx_unpack_shift(en.y, sx, sy, sz);
      double ax, ay, az;
      atx_image_vector(A, sx, sy, sz, ax, ay, az);
      dx += ax; dy += ay; dz += az;
    }
    double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < cutoff_sq) rho += spl_f(T->frho[dbj - 1], sqrt(r2), &err);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    rho += __shfl_xor_sync(0xffffffffu, rho, o);
    err |= __shfl_xor_sync(0xffffffffu, err, o);
  }
  if (lane == 0 && valid) {
    double F = 0.0, dFi = 0.0;
    if (active) {
      if (rho < 0.0) rho = 0.0;
      spl_f_df_x(T->fF[dbi - 1], rho, F, dFi);
    }
    dF[s] = dFi;
    Fe[s] = F;
    if (err) atomicOr(flag, 1);
  }
}

template <int LANES, bool PER_AT>
__global__ void __launch_bounds__(128)
k_eam_force(int nat, Mat3 A, const EamDev *__restrict__ T, const double4 *__restrict__ pos4,
            const long long *__restrict__ seed, const int2 *__restrict__ list,
            const int *__restrict__ mask, const double *__restrict__ dF,
            const double *__restrict__ Fe, double *__restrict__ f, double *__restrict__ epa,
            double *__restrict__ wpa, double *__restrict__ partials, int *__restrict__ flag,
            const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * 4];
  const int gpb = 128 / LANES;
  int s = blockIdx.x * gpb + threadIdx.x / LANES;
  int lane = threadIdx.x % LANES;
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;

  const bool valid = s < nat;
  // only owned atoms (role 2) are centres of the force pass: ghost rows receive zero, and the energy /
  // virial of a ghost is counted by the rank (or the periodic image) that owns it
  const bool own = valid && (!role || role[s] >= 2);
  double4 pi = valid ? pos4[s] : make_double4(0, 0, 0, 0);
  int dbi = own ? T->el2db[(int)pi.w] : -1;
  double fx = 0.0, fy = 0.0, fz = 0.0, e = 0.0;
  double wxx = 0, wyy = 0, wzz = 0, wxy = 0, wxz = 0, wyz = 0;    // total virial, i-visits
  double pxx = 0, pyy = 0, pzz = 0, pxy = 0, pxz = 0, pyz = 0;    // per-atom virial
  int err = 0;
  if (dbi > 0) {
    bool act_i = (!mask || mask[s] != 0);
    double dFi = dF[s];
    const double cutoff_sq = T->cutoff_sq;
    const DSpline &rho_i = T->frho[dbi - 1];
    long long b = seed[s], en_ = seed[s + 1];
    for (long long a = b + lane; a < en_; a += LANES) {
      int2 en = list[a];
      double4 pj = pos4[en.x];
      int dbj = T->el2db[(int)pj.w];
      if (dbj <= 0) continue;
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      if (ATX_NONZERO_SHIFT(en.y)) {
        int sx, sy, sz;
        atx_unpack_shift(en.y, sx, sy, sz);
        double ax, ay, az;
        atx_image_vector(A, sx, sy, sz, ax, ay, az);
        dx += ax; dy += ay; dz += az;
      }
      double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 >= cutoff_sq) continue;
      double r = sqrt(r2);
      double rinv = 1.0 / r;
      bool act_j = (!mask || mask[en.x] != 0);
      double phi, dphi;
      spl_f_df_x(T->fphi[(dbi - 1) + T->ndb * (dbj - 1)], r, phi, dphi);
      double pair = (dphi - phi * rinv) * rinv;
      double c = 0.0, cij = 0.0;
      if (act_i) {
        double drho_j = spl_df(T->frho[dbj - 1], r, &err);
        cij = -(dFi * drho_j + pair) * rinv;
        c = cij;
        e += phi * rinv;
      }
      if (act_j) {
        double drho_i = spl_df(rho_i, r, &err);
        c += -(dF[en.x] * drho_i + pair) * rinv;
      }
      fx += c * dx; fy += c * dy; fz += c * dz;
      // wij = -outer(dr, df), df = cij*dr (tabulated_alloy_eam.f90:598-599)
      wxx -= cij * dx * dx; wyy -= cij * dy * dy; wzz -= cij * dz * dz;
      wxy -= cij * dx * dy; wxz -= cij * dx * dz; wyz -= cij * dy * dz;
      if (PER_AT) {
        double h = -0.5 * c;
        pxx += h * dx * dx; pyy += h * dy * dy; pzz += h * dz * dz;
        pxy += h * dx * dy; pxz += h * dx * dz; pyz += h * dy * dz;
      }
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
    err |= __shfl_xor_sync(0xffffffffu, err, o);
    if (PER_AT) {
      pxx += __shfl_xor_sync(0xffffffffu, pxx, o);
      pyy += __shfl_xor_sync(0xffffffffu, pyy, o);
      pzz += __shfl_xor_sync(0xffffffffu, pzz, o);
      pxy += __shfl_xor_sync(0xffffffffu, pxy, o);
      pxz += __shfl_xor_sync(0xffffffffu, pxz, o);
      pyz += __shfl_xor_sync(0xffffffffu, pyz, o);
    }
  }
  if (lane == 0 && valid) {
    f[3 * s] = fx; f[3 * s + 1] = fy; f[3 * s + 2] = fz;
    double ei = own ? e + Fe[s] : 0.0;
    if (epa) epa[s] = ei;
    if (PER_AT) {
      double *w = &wpa[9 * (size_t)s];
      w[0] = pxx; w[1] = pxy; w[2] = pxz;
      w[3] = pxy; w[4] = pyy; w[5] = pyz;
      w[6] = pxz; w[7] = pyz; w[8] = pzz;
    }
    acc[0] = ei;
    if (err) atomicOr(flag, 1);
  }
  // every lane carries its share of the total virial into the block sum
  acc[1] = wxx; acc[2] = wxy; acc[3] = wxz;
  acc[4] = wxy; acc[5] = wyy; acc[6] = wyz;
  acc[7] = wxz; acc[8] = wyz; acc[9] = wzz;
  atx_block_sum<ATX_NSUM, 128>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// Fast kernels (no mask, no per-atom virial): U list entries per lane are processed together so
// that the position gathers and the spline-row loads of U pairs are in flight at the same time
// (the v0 kernels were stalled on dependent loads, profiles/r01_ncu_eam_v0.csv), the three r-tables
// share one interval index, and (r - x0)/dx is evaluated as (r - x0)*(1/dx) -- a <= 1 ulp change of
// the interval coordinate that a C2 spline turns into a relative change far below 1e-10.
// ---------------------------------------------------------------------------

// CONSEC = 0: lane l takes the U neighbouring entries (t*LANES + l)*U + u.  CONSEC = 1: entry
// t*LANES*U + u*LANES + l, so that the lanes of ONE load instruction read neighbouring list entries
// (which point at runs of neighbouring atoms inside a binning cell): fewer distinct 128-byte lines
// per gather in the model of benchmarks/model_gather_wavefronts.py.  Selected by ATX_EAM_MAP=1;
// not measured yet.
template <int LANES, int U, int CONSEC = 0>
__global__ void __launch_bounds__(128)
k_eam_density_fast(int nat, Mat3 A, const EamDev *__restrict__ T, const double4 *__restrict__ pos4,
                   const long long *__restrict__ seed, const int2 *__restrict__ list,
                   double4 *__restrict__ pd4, double *__restrict__ Fe,
                   const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  const int gpb = 128 / LANES;
  const int s = blockIdx.x * gpb + threadIdx.x / LANES;
  const int lane = threadIdx.x % LANES;
  const bool valid = s < nat;
  const double4 pi = valid ? pos4[s] : make_double4(0, 0, 0, 0);
  const int dbi = (valid && (!role || role[s] >= 1)) ? T->el2db[(int)pi.w] : -1;
  const double cutoff_sq = T->cutoff_sq, x0 = T->r_x0, inv_dx = T->r_inv_dx;
  const int nr = T->r_n;
  double rho = 0.0;
  const long long b = dbi > 0 ? seed[s] : 0, e = dbi > 0 ? seed[s + 1] : 0;
  constexpr int ES = CONSEC ? LANES : 1;   // distance between the U entries of one lane
  for (long long a0 = b + (CONSEC ? lane : lane * U); a0 < e; a0 += LANES * U) {
    int2 en[U];
    double4 pj[U];
#pragma unroll
    for (int u = 0; u < U; u++) en[u] = (a0 + u * ES < e) ? list[a0 + u * ES] : make_int2(s, ATX_SHIFT_ZERO);
#pragma unroll
    for (int u = 0; u < U; u++) pj[u] = atx_ld4(&pos4[en[u].x]);
    double B[U], w[U];
    const double4 *row[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      double dx = pi.x - pj[u].x, dy = pi.y - pj[u].y, dz = pi.z - pj[u].z;
      if (ATX_NONZERO_SHIFT(en[u].y)) {
        int sx, sy, sz;
        atx_unpack_shift(en[u].y, sx, sy, sz);
        double ax, ay, az;
        atx_image_vector(A, sx, sy, sz, ax, ay, az);
        dx += ax; dy += ay; dz += az;
      }
      const double r2 = dx * dx + dy * dy + dz * dz;
      const int dbj = T->el2db[ATX_ENTRY_EL(en[u].y)];
      const bool in = (a0 + u * ES < e) && dbj > 0 && r2 < cutoff_sq;
      const double r = sqrt(in ? r2 : 1.0);
      const double xf = (r - x0) * inv_dx + 1.0;
      int i = (int)floor(xf);
      i = i < 1 ? 1 : (i >= nr ? nr - 1 : i);
      B[u] = xf - (double)i;
      w[u] = in ? 1.0 : 0.0;
      row[u] = &T->frho[in ? dbj - 1 : 0].c[i - 1];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double4 c = atx_ld4(row[u]);
      rho += w[u] * (c.x + B[u] * (c.y + B[u] * (c.z + B[u] * c.w)));
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) rho += __shfl_xor_sync(0xffffffffu, rho, o);
  if (lane == 0 && valid) {
    double F = 0.0, dFi = 0.0;
    if (dbi > 0) {
      if (rho < 0.0) rho = 0.0;
      spl_f_df_x(T->fF[dbi - 1], rho, F, dFi);
    }
    pd4[s] = make_double4(pi.x, pi.y, pi.z, dFi);  // position + F'(rho): one gather in the force pass
    Fe[s] = F;
  }
}

// Force pass.  Per pair it gathers ONE 32-byte record {x,y,z,F'_j} and ONE 64-byte table record
// {phi: y c1 c2 c3 | rho_j: c1 c2 c3 | -}; derivative coefficients are formed as k*c_k/dx.
template <int LANES, int U, bool VIRIAL, int CONSEC = 0>
__global__ void __launch_bounds__(128)
k_eam_force_fast(int nat, Mat3 A, const EamDev *__restrict__ T, const double4 *__restrict__ pd4,
                 const double4 *__restrict__ pos4, const long long *__restrict__ seed,
                 const int2 *__restrict__ list, const double *__restrict__ Fe, double *__restrict__ f,
                 double *__restrict__ epa, double *__restrict__ partials,
                 const unsigned char *__restrict__ role, const int *__restrict__ stop) {
  if (stop && *stop) return;
  __shared__ double red[ATX_NSUM * 4];
  const int gpb = 128 / LANES;
  const int s = blockIdx.x * gpb + threadIdx.x / LANES;
  const int lane = threadIdx.x % LANES;
  const bool valid = s < nat;
  const double4 pi = valid ? pd4[s] : make_double4(0, 0, 0, 0);
  const int dbi = (valid && (!role || role[s] >= 2)) ? T->el2db[(int)pos4[s].w] : -1;
  const double cutoff_sq = T->cutoff_sq, x0 = T->r_x0, inv_dx = T->r_inv_dx;
  const int nr = T->r_n, ndb = T->ndb;
  double fx = 0.0, fy = 0.0, fz = 0.0, en_ = 0.0;
  double wxx = 0, wyy = 0, wzz = 0, wxy = 0, wxz = 0, wyz = 0;
  const double dFi = pi.w;
  const bool zsq = T->zsq != 0;
  const long long b = dbi > 0 ? seed[s] : 0, e = dbi > 0 ? seed[s + 1] : 0;
  const int di = dbi > 0 ? dbi - 1 : 0;
  constexpr int ES = CONSEC ? LANES : 1;
  for (long long a0 = b + (CONSEC ? lane : lane * U); a0 < e; a0 += LANES * U) {
    int2 en[U];
    double4 pj[U];
#pragma unroll
    for (int u = 0; u < U; u++) en[u] = (a0 + u * ES < e) ? list[a0 + u * ES] : make_int2(s, ATX_SHIFT_ZERO);
#pragma unroll
    for (int u = 0; u < U; u++) pj[u] = atx_ld4(&pd4[en[u].x]);
    double B[U], w[U], dx[U], dy[U], dz[U], rinv[U];
    int idx[U], dj[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      dx[u] = pi.x - pj[u].x; dy[u] = pi.y - pj[u].y; dz[u] = pi.z - pj[u].z;
      if (ATX_NONZERO_SHIFT(en[u].y)) {
        int sx, sy, sz;
        atx_unpack_shift(en[u].y, sx, sy, sz);
        double ax, ay, az;
        atx_image_vector(A, sx, sy, sz, ax, ay, az);
        dx[u] += ax; dy[u] += ay; dz[u] += az;
      }
      const double r2 = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
      const int dbj = T->el2db[ATX_ENTRY_EL(en[u].y)];
      const bool in = (a0 + u * ES < e) && dbj > 0 && r2 < cutoff_sq;
      const double r = sqrt(in ? r2 : 1.0);
      rinv[u] = 1.0 / r;
      const double xf = (r - x0) * inv_dx + 1.0;
      int i = (int)floor(xf);
      i = i < 1 ? 1 : (i >= nr ? nr - 1 : i);
      B[u] = xf - (double)i;
      idx[u] = i - 1;
      w[u] = in ? 1.0 : 0.0;
      dj[u] = in ? dbj - 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double4 *rec = T->pairrec[di + ndb * dj[u]] + 2 * (size_t)idx[u];
      const double4 c = atx_ld4(rec);       // phi: y c1 c2 c3
      const double4 q = atx_ld4(rec + 1);   // rho_j: c1 c2 c3
      const double Bu = B[u];
      double phi = c.x + Bu * (c.y + Bu * (c.z + Bu * c.w));
      double dphi = (c.y + Bu * (2.0 * c.z + Bu * (3.0 * c.w))) * inv_dx;
      if (zsq) {
        dphi = 2.0 * phi * dphi;
        phi = phi * phi;
      }
      const double drho_j = (q.x + Bu * (2.0 * q.y + Bu * (3.0 * q.z))) * inv_dx;
      double drho_i = drho_j;
      if (dj[u] != di) {
        const double4 ri = atx_ld4(&T->frho[di].c[idx[u]]);
        drho_i = (ri.y + Bu * (2.0 * ri.z + Bu * (3.0 * ri.w))) * inv_dx;
      }
      const double ri_ = rinv[u];
      const double pair = (dphi - phi * ri_) * ri_;
      const double cij = -(dFi * drho_j + pair) * ri_ * w[u];
      const double cji = -(pj[u].w * drho_i + pair) * ri_ * w[u];
      const double cc = cij + cji;
      fx += cc * dx[u]; fy += cc * dy[u]; fz += cc * dz[u];
      en_ += w[u] * phi * ri_;
      if (VIRIAL) {
        wxx -= cij * dx[u] * dx[u]; wyy -= cij * dy[u] * dy[u]; wzz -= cij * dz[u] * dz[u];
        wxy -= cij * dx[u] * dy[u]; wxz -= cij * dx[u] * dz[u]; wyz -= cij * dy[u] * dz[u];
      }
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, o);
    fy += __shfl_xor_sync(0xffffffffu, fy, o);
    fz += __shfl_xor_sync(0xffffffffu, fz, o);
    en_ += __shfl_xor_sync(0xffffffffu, en_, o);
  }
  double acc[ATX_NSUM];
#pragma unroll
  for (int k = 0; k < ATX_NSUM; k++) acc[k] = 0.0;
  if (lane == 0 && valid) {
    f[3 * s] = fx; f[3 * s + 1] = fy; f[3 * s + 2] = fz;
    const double ei = en_ + Fe[s];
    if (epa) epa[s] = dbi > 0 ? ei : 0.0;
    if (dbi > 0) acc[0] = ei;   // ghosts (role < 2) carry an embedding energy that their owner counts
  }
  if (VIRIAL) {
    acc[1] = wxx; acc[2] = wxy; acc[3] = wxz;
    acc[4] = wxy; acc[5] = wyy; acc[6] = wyz;
    acc[7] = wxz; acc[8] = wyz; acc[9] = wzz;
  }
  atx_block_sum<ATX_NSUM, 128>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < ATX_NSUM; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
  }
}

static int upload_spline(atx_eam *pot, const atx_spline &s, DSpline &d) {
  int ni = s.n - 1;
  std::vector<double4> c(ni), dd(ni);
  for (int i = 0; i < ni; i++) {
    c[i] = make_double4(s.y[i], s.coeff1[i], s.coeff2[i], s.coeff3[i]);
    dd[i] = make_double4(s.dcoeff1[i], s.dcoeff2[i], s.dcoeff3[i], 0.0);
  }
  auto *bc = new DevBuf<double4>();
  auto *bd = new DevBuf<double4>();
  pot->tables.push_back(bc);
  pot->tables.push_back(bd);
  ATX_PASS(bc->reserve(ni));
  ATX_PASS(bd->reserve(ni));
  ATX_CUDA(cudaMemcpy(bc->ptr, c.data(), sizeof(double4) * ni, cudaMemcpyHostToDevice));
  ATX_CUDA(cudaMemcpy(bd->ptr, dd.data(), sizeof(double4) * ni, cudaMemcpyHostToDevice));
  d.n = s.n;
  d.x0 = s.x0;
  d.dx = s.dx;
  d.cut = s.x0 + s.dx * (s.n - 1);
  d.c = bc->ptr;
  d.d = bd->ptr;
  return 0;
}

extern "C" int atx_eam_create(atx_ctx *ctx, int ndb, const atx_spline *fF, const atx_spline *frho,
                              const atx_spline *fphi, double cutoff, atx_eam **out) {
  if (ctx) cudaSetDevice(ctx->device);  // entry points do not assume the caller kept the device current
  if (ndb < 1 || ndb > EAM_MAX_DB) {
    atx_set_error("TabulatedAlloyEAM supports 1.." + std::to_string(EAM_MAX_DB) + " elements.");
    return ATX_ERROR_UNSPECIFIED;
  }
  ATX_CUDA(cudaSetDevice(ctx->device));
  atx_eam *pot = new atx_eam();
  pot->ctx = ctx;
  pot->ndb = ndb;
  pot->cutoff = cutoff;
  pot->host.ndb = ndb;
  pot->host.cutoff_sq = cutoff * cutoff;
  for (int i = 0; i < ndb; i++) {
    ATX_PASS(upload_spline(pot, fF[i], pot->host.fF[i]));
    ATX_PASS(upload_spline(pot, frho[i], pot->host.frho[i]));
  }
  for (int j = 0; j < ndb; j++)
    for (int i = 0; i < ndb; i++)
      ATX_PASS(upload_spline(pot, fphi[i + ndb * j], pot->host.fphi[i + ndb * j]));
  for (int k = 0; k < 32; k++) pot->host.el2db[k] = -1;
  {
    const atx_spline &g0 = frho[0];
    bool same = true;
    for (int i = 0; i < ndb; i++) same = same && frho[i].n == g0.n && frho[i].x0 == g0.x0 && frho[i].dx == g0.dx;
    for (int i = 0; i < ndb * ndb; i++) same = same && fphi[i].n == g0.n && fphi[i].x0 == g0.x0 && fphi[i].dx == g0.dx;
    pot->host.r_n = g0.n;
    pot->host.r_x0 = g0.x0;
    pot->host.r_inv_dx = 1.0 / g0.dx;
    // func() without extrapolation must never leave the table: cutoff <= last knot
    pot->fast_ok = same && g0.x0 <= 0.0 && cutoff <= g0.x0 + g0.dx * (g0.n - 1);
    if (pot->fast_ok) {
      const int ni = g0.n - 1;
      std::vector<double4> rec(2 * (size_t)ni);
      for (int j = 0; j < ndb; j++)
        for (int i = 0; i < ndb; i++) {
          const atx_spline &ph = fphi[i + ndb * j];
          const atx_spline &rj = frho[j];
          for (int k = 0; k < ni; k++) {
            rec[2 * k] = make_double4(ph.y[k], ph.coeff1[k], ph.coeff2[k], ph.coeff3[k]);
            rec[2 * k + 1] = make_double4(rj.coeff1[k], rj.coeff2[k], rj.coeff3[k], 0.0);
          }
          auto *bb = new DevBuf<double4>();
          pot->tables.push_back(bb);
          ATX_PASS(bb->reserve(2 * (size_t)ni));
          ATX_CUDA(cudaMemcpy(bb->ptr, rec.data(), sizeof(double4) * 2 * ni, cudaMemcpyHostToDevice));
          pot->host.pairrec[i + ndb * j] = bb->ptr;
        }
    }
  }
  if (const char *v = getenv("ATX_EAM_GENERIC")) pot->force_generic = atoi(v) != 0;
  if (const char *v = getenv("ATX_EAM_LANES")) pot->fast_lanes = atoi(v);
  if (const char *v = getenv("ATX_EAM_UNROLL")) pot->fast_unroll = atoi(v);
  if (const char *v = getenv("ATX_EAM_MAP")) pot->fast_map = atoi(v) != 0;
  ATX_PASS(pot->dev.reserve(1));
  ATX_PASS(pot->flag.reserve(4));
  *out = pot;
  return 0;
}

extern "C" int atx_eam_create_funcfl(atx_ctx *ctx, const atx_spline *fF, const atx_spline *frho,
                                     const atx_spline *fZ, double cutoff, atx_eam **out) {
  ATX_PASS(atx_eam_create(ctx, 1, fF, frho, fZ, cutoff, out));
  atx_eam *pot = *out;
  if (!pot->fast_ok) {
    atx_set_error("TabulatedEAM: the Z and rho tables must share one grid that starts at 0 and covers the cutoff.");
    delete pot;
    *out = nullptr;
    return ATX_ERROR_UNSPECIFIED;
  }
  pot->funcfl = true;
  pot->host.zsq = 1;
  return 0;
}

extern "C" int atx_eam_destroy(atx_eam *pot) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  delete pot;
  return 0;
}

extern "C" int atx_eam_bind_to(atx_eam *pot, atx_particles *p, atx_neighbors *nl, int nel,
                               const int *el2db) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (nel > 31) {
    atx_set_error("Too many particle element ids.");
    return ATX_ERROR_UNSPECIFIED;
  }
  for (int k = 0; k < 32; k++) pot->host.el2db[k] = -1;
  bool any = false;
  for (int k = 0; k < nel; k++) {
    pot->host.el2db[k + 1] = el2db[k];
    any = any || el2db[k] > 0;
  }
  ATX_CUDA(cudaMemcpy(pot->dev.ptr, &pot->host, sizeof(EamDev), cudaMemcpyHostToDevice));
  if (any && nl)   // tabulated_alloy_eam.f90:297-350: every pair of elements the tables cover
    for (int i = 1; i <= nel; i++)
      for (int j = i; j <= nel; j++)
        if (el2db[i - 1] > 0 && el2db[j - 1] > 0)
          ATX_PASS(atx_neighbors_request_interaction_range_pair(nl, pot->cutoff, i, j));
  pot->bound = true;
  return 0;
}

// device-resident evaluation (sorted order); used by library mode and the MD driver
int atx_eam_compute_device(atx_eam *pot, atx_particles *p, atx_neighbors *nl,
                           const int *mask_sorted, const PotOut &o) {
  atx_ctx *ctx = pot->ctx;
  cudaStream_t st = ctx->stream;
  int nat = nl->nat;
  ATX_PASS(pot->dF.reserve(nat + 1));
  ATX_PASS(pot->Fe.reserve(nat + 1));
  ATX_PASS(pot->pd4.reserve(nat + 1));
  // lanes per atom from the mean list length
  double mean = nat > 0 ? (double)nl->npairs / nat : 0.0;
  int lanes = mean > 48 ? 16 : (mean > 20 ? 8 : 4);
  int gpb = 128 / lanes;
  int nblocks = (nat + gpb - 1) / gpb;
  if (nblocks < 1) nblocks = 1;
  ATX_PASS(pot->sc.partials.reserve((size_t)nblocks * ATX_NSUM));
  if (!o.stop) ATX_CUDA(cudaMemsetAsync(pot->flag.ptr, 0, sizeof(int), st));
  if (pot->fast_ok && !mask_sorted && !o.wpa && (!pot->force_generic || pot->funcfl)) {
    const int L = pot->fast_lanes, gp = 128 / L;
    const int nb = nat > 0 ? (nat + gp - 1) / gp : 1;
    ATX_PASS(pot->sc.partials.reserve((size_t)nb * ATX_NSUM));
    const bool vir = o.want_virial;
#define EAM_FAST_M(LL, UU, MM)                                                                         \
  do {                                                                                            \
    {                                                                                             \
      ProfScope ps_(ctx, "eam_density");                                                          \
      k_eam_density_fast<LL, UU, MM><<<nb, 128, 0, st>>>(nat, p->Abox, pot->dev.ptr, nl->pos4.ptr,    \
                                                     nl->seed.ptr, nl->list.ptr, pot->pd4.ptr,    \
                                                     pot->Fe.ptr, o.role, o.stop);                \
    }                                                                                             \
    ATX_LAUNCHED();                                                                               \
    ProfScope ps2_(ctx, "eam_force");                                                             \
    if (vir)                                                                                      \
      k_eam_force_fast<LL, UU, true, MM><<<nb, 128, 0, st>>>(nat, p->Abox, pot->dev.ptr, pot->pd4.ptr, \
          nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, pot->Fe.ptr, o.f, o.epa, pot->sc.partials.ptr, \
          o.role, o.stop);                                                                        \
    else                                                                                          \
      k_eam_force_fast<LL, UU, false, MM><<<nb, 128, 0, st>>>(nat, p->Abox, pot->dev.ptr, pot->pd4.ptr, \
          nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, pot->Fe.ptr, o.f, o.epa, pot->sc.partials.ptr, \
          o.role, o.stop);                                                                        \
    ATX_LAUNCHED();                                                                               \
  } while (0)
#define EAM_FAST(LL, UU) EAM_FAST_M(LL, UU, 0)
    const int U = pot->fast_unroll;
    if (pot->fast_map && L == 4 && U == 2) EAM_FAST_M(4, 2, 1);
    else if (pot->fast_map && L == 4 && U == 4) EAM_FAST_M(4, 4, 1);
    else if (pot->fast_map && L == 8 && U == 2) EAM_FAST_M(8, 2, 1);
    else if (pot->fast_map && L == 8 && U == 4) EAM_FAST_M(8, 4, 1);
    else if (pot->fast_map && L == 16 && U == 2) EAM_FAST_M(16, 2, 1);
    else if (pot->fast_map) EAM_FAST_M(16, 1, 1);
    else if (L == 4 && U == 4) EAM_FAST(4, 4);
    else if (L == 4 && U == 2) EAM_FAST(4, 2);
    else if (L == 8 && U == 4) EAM_FAST(8, 4);
    else if (L == 8 && U == 2) EAM_FAST(8, 2);
    else if (L == 2 && U == 4) EAM_FAST(2, 4);
    else if (L == 16 && U == 2) EAM_FAST(16, 2);
    else if (L == 16 && U == 1) EAM_FAST(16, 1);
    else if (L == 1 && U == 4) EAM_FAST(1, 4);
    else EAM_FAST(8, 2);
#undef EAM_FAST
#undef EAM_FAST_M
    if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nb, o.sums, o.stop));
    return 0;
  }
#define EAM_LAUNCH(L)                                                                             \
  do {                                                                                            \
    {                                                                                             \
    ProfScope ps_(ctx, "eam_density");                                                            \
    k_eam_density<L><<<nblocks, 128, 0, st>>>(nat, p->Abox, pot->dev.ptr, nl->pos4.ptr,           \
                                              nl->seed.ptr, nl->list.ptr, mask_sorted,            \
                                              pot->dF.ptr, pot->Fe.ptr, pot->flag.ptr, o.role,    \
                                              o.stop);                                            \
    }                                                                                             \
    ATX_LAUNCHED();                                                                               \
    ProfScope ps2_(ctx, "eam_force");                                                             \
    if (o.wpa)                                                                                    \
      k_eam_force<L, true><<<nblocks, 128, 0, st>>>(                                              \
          nat, p->Abox, pot->dev.ptr, nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, mask_sorted,      \
          pot->dF.ptr, pot->Fe.ptr, o.f, o.epa, o.wpa, pot->sc.partials.ptr, pot->flag.ptr,      \
          o.role, o.stop);                                                                        \
    else                                                                                          \
      k_eam_force<L, false><<<nblocks, 128, 0, st>>>(                                             \
          nat, p->Abox, pot->dev.ptr, nl->pos4.ptr, nl->seed.ptr, nl->list.ptr, mask_sorted,      \
          pot->dF.ptr, pot->Fe.ptr, o.f, o.epa, o.wpa, pot->sc.partials.ptr, pot->flag.ptr,      \
          o.role, o.stop);                                                                        \
    ATX_LAUNCHED();                                                                               \
  } while (0)
  if (lanes == 16) EAM_LAUNCH(16);
  else if (lanes == 8) EAM_LAUNCH(8);
  else EAM_LAUNCH(4);
#undef EAM_LAUNCH
  if (o.want_sums) ATX_PASS(atx_reduce_partials(ctx, pot->sc.partials.ptr, nblocks, o.sums, o.stop));
  return 0;
}

int atx_eam_check_flag(atx_eam *pot) {
  int h = 0;
  ATX_CUDA(cudaMemcpyAsync(&h, pot->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, pot->ctx->stream));
  ATX_CUDA(cudaStreamSynchronize(pot->ctx->stream));
  if (h) {
    atx_set_error("simple_spline: x outside of the defined interval (density or pair table).");
    return ATX_ERROR_UNSPECIFIED;
  }
  return 0;
}

extern "C" int atx_eam_energy_and_forces(atx_eam *pot, atx_particles *p, atx_neighbors *nl,
                                         const int *mask, double *epot, double *f, double *wpot,
                                         double *epot_per_at, double *wpot_per_at) {
  if (pot && pot->ctx) cudaSetDevice(pot->ctx->device);
  if (!pot->bound) {
    atx_set_error("TabulatedAlloyEAM: bind_to has not been called.");
    return ATX_ERROR_UNSPECIFIED;
  }
  if (pot->funcfl && (mask || wpot_per_at)) {
    // tabulated_eam_energy_and_forces has no mask and never fills wpot_per_at
    atx_set_error("TabulatedEAM does not support masks or per-atom virials.");
    return ATX_ERROR_UNSPECIFIED;
  }
  ATX_PASS(atx_neighbors_update(nl, p));
  PotOut o;
  ATX_PASS(atx_prepare_out(pot->ctx, nl, pot->sc, epot_per_at != nullptr, wpot_per_at != nullptr, o));
  if (nl->external) o.role = nl->role_ext.ptr;   // masks / per-atom virials: the generic kernels honour roles too
  const int *mask_sorted = nullptr;
  ATX_PASS(atx_prepare_mask(pot->ctx, nl, pot->sc, mask, &mask_sorted));
  ATX_PASS(atx_eam_compute_device(pot, p, nl, mask_sorted, o));
  ATX_PASS(atx_eam_check_flag(pot));
  return atx_finish_to_host(pot->ctx, nl, pot->sc, o, epot, f, wpot, epot_per_at, wpot_per_at);
}

extern "C" int atx_eam_set_store_outputs(atx_eam *pot, int on) {
  if (!pot) return ATX_ERROR_UNSPECIFIED;
  pot->sc.store_outputs = on != 0;
  return 0;
}
