// Host-side init helpers (no device code).
//
// The reference builds these once per bind_to/init on the host in Fortran:
//   simple_spline_init           src/support/simple_spline.f90:127-195
//   gaussn (LAPACK dgesv)        src/support/f_linearalgebra.f90:599-637
//   table2d_init / table3d_init  src/special/table2d.f90:84-226, table3d.f90:85-284
//   rebo2_db_make_cc_g_spline    src/potentials/bop/rebo2/rebo2_db.f90:405-524
// A Fortran host keeps its own code and hands the results to atx_*_create; hosts without the
// Fortran layer (the Python mirror in this repo) call these.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/atomistica_b200.h"

void atx_set_error(const std::string &msg);

extern "C" int atx_host_spline_init(int n, double x0, double dx, const double *y_in, double *y,
                                    double *d2y, double *coeff1, double *coeff2, double *coeff3,
                                    double *dcoeff1, double *dcoeff2, double *dcoeff3) {
  if (n < 2) {
    atx_set_error("simple_spline_init: need at least two points.");
    return ATX_ERROR_UNSPECIFIED;
  }
  (void)x0;
  const double sig = 0.5;
  std::vector<double> u(n, 0.0);
  for (int i = 0; i < n; i++) y[i] = y_in[i];
  d2y[0] = 0.0;  // natural spline
  u[0] = 0.0;
  for (int i = 1; i < n - 1; i++) {
    double p = sig * d2y[i - 1] + 2;
    d2y[i] = (sig - 1) / p;
    u[i] = (6.0 * ((y[i + 1] - y[i]) / dx - (y[i] - y[i - 1]) / dx) / (2 * dx) - sig * u[i - 1]) / p;
  }
  const double qn = 0.0, un = 0.0;
  d2y[n - 1] = (un - qn * u[n - 2]) / (qn * d2y[n - 2] + 1.0);
  for (int k = n - 2; k >= 0; k--) d2y[k] = d2y[k] * d2y[k + 1] + u[k];
  const double dx2 = dx * dx;
  for (int k = 0; k < n - 1; k++) {
    coeff1[k] = y[k + 1] - y[k] - (2 * d2y[k] + d2y[k + 1]) * dx2 / 6;
    coeff2[k] = d2y[k] * dx2 / 2;
    coeff3[k] = (d2y[k + 1] - d2y[k]) * dx2 / 6;
    dcoeff1[k] = coeff1[k] / dx;
    dcoeff2[k] = 2 * coeff2[k] / dx;
    dcoeff3[k] = 3 * coeff3[k] / dx;
  }
  return 0;
}

// LU with partial pivoting (what dgesv does), column-major, in place
extern "C" int atx_host_gaussn(int n, double *A, int m, double *B) {
  std::vector<int> piv(n);
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[k + (size_t)n * k]);
    for (int i = k + 1; i < n; i++) {
      double v = std::fabs(A[i + (size_t)n * k]);
      if (v > best) { best = v; p = i; }
    }
    if (best == 0.0) {
      atx_set_error("gaussn: singular matrix (dgesv info = " + std::to_string(k + 1) + ")");
      return ATX_ERROR_UNSPECIFIED;
    }
    piv[k] = p;
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(A[k + (size_t)n * j], A[p + (size_t)n * j]);
      for (int j = 0; j < m; j++) std::swap(B[k + (size_t)n * j], B[p + (size_t)n * j]);
    }
    double inv = 1.0 / A[k + (size_t)n * k];
    for (int i = k + 1; i < n; i++) A[i + (size_t)n * k] *= inv;
    for (int j = k + 1; j < n; j++) {
      double akj = A[k + (size_t)n * j];
      if (akj != 0.0)
        for (int i = k + 1; i < n; i++) A[i + (size_t)n * j] -= A[i + (size_t)n * k] * akj;
    }
    for (int j = 0; j < m; j++) {
      double bkj = B[k + (size_t)n * j];
      if (bkj != 0.0)
        for (int i = k + 1; i < n; i++) B[i + (size_t)n * j] -= A[i + (size_t)n * k] * bkj;
    }
  }
  for (int j = 0; j < m; j++)
    for (int k = n - 1; k >= 0; k--) {
      double x = B[k + (size_t)n * j] / A[k + (size_t)n * k];
      B[k + (size_t)n * j] = x;
      if (x != 0.0)
        for (int i = 0; i < k; i++) B[i + (size_t)n * j] -= A[i + (size_t)n * k] * x;
    }
  return 0;
}

static double ipowi(int b, int e) {
  double v = 1.0;
  for (int i = 0; i < e; i++) v *= b;
  return v;
}

extern "C" int atx_host_table2d_init(int nx, int ny, const double *values, const double *dvdx,
                                     const double *dvdy, double *coeff) {
  const int npara = 16, ncorn = 4;
  static const int ix1[4] = {0, 1, 1, 0}, ix2[4] = {0, 0, 1, 1};
  std::vector<double> A(npara * npara, 0.0);
  auto a = [&](int r, int c) -> double & { return A[r + npara * c]; };
  for (int ic = 0; ic < ncorn; ic++)
    for (int p1 = 0; p1 < 4; p1++)
      for (int p2 = 0; p2 < 4; p2++) {
        int p1m = p1 > 0 ? p1 - 1 : 0, p2m = p2 > 0 ? p2 - 1 : 0;
        int col = 4 * p1 + p2, n1 = ix1[ic], n2 = ix2[ic];
        a(ic, col) = ipowi(n1, p1) * ipowi(n2, p2);
        a(ic + 4, col) = p1 * ipowi(n1, p1m) * ipowi(n2, p2);
        a(ic + 8, col) = ipowi(n1, p1) * p2 * ipowi(n2, p2m);
        a(ic + 12, col) = p1 * ipowi(n1, p1m) * p2 * ipowi(n2, p2m);
      }
  int nboxs = nx * ny;
  std::vector<double> B((size_t)npara * nboxs, 0.0);
  auto v2 = [&](const double *t, int i, int j) { return t[i + (nx + 1) * j]; };
  for (int nh = 0; nh < nx; nh++)
    for (int nc = 0; nc < ny; nc++) {
      int col = ny * nh + nc;
      for (int ic = 0; ic < ncorn; ic++) {
        int n1 = ix1[ic] + nh, n2 = ix2[ic] + nc;
        B[ic + (size_t)npara * col] = v2(values, n1, n2);
        if (dvdx) B[ic + ncorn + (size_t)npara * col] = v2(dvdx, n1, n2);
        if (dvdy) B[ic + 2 * ncorn + (size_t)npara * col] = v2(dvdy, n1, n2);
      }
    }
  int err = atx_host_gaussn(npara, A.data(), nboxs, B.data());
  if (err) return err;
  for (int ibox = 0; ibox < nboxs; ibox++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        coeff[ibox + (size_t)nboxs * (i + 4 * j)] = B[(4 * i + j) + (size_t)npara * ibox];
  return 0;
}

extern "C" int atx_host_table3d_init(int nx, int ny, int nz, const double *values,
                                     const double *dvdx, const double *dvdy, const double *dvdz,
                                     double *coeff) {
  const int npara = 64, ncorn = 8;
  static const int ix1[8] = {0, 1, 1, 0, 0, 1, 1, 0}, ix2[8] = {0, 0, 1, 1, 0, 0, 1, 1},
                   ix3[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  std::vector<double> A(npara * npara, 0.0);
  auto a = [&](int r, int c) -> double & { return A[r + npara * c]; };
  for (int ic = 0; ic < ncorn; ic++)
    for (int p1 = 0; p1 < 4; p1++)
      for (int p2 = 0; p2 < 4; p2++)
        for (int p3 = 0; p3 < 4; p3++) {
          int p1m = p1 > 0 ? p1 - 1 : 0, p2m = p2 > 0 ? p2 - 1 : 0, p3m = p3 > 0 ? p3 - 1 : 0;
          int col = 16 * p1 + 4 * p2 + p3, n1 = ix1[ic], n2 = ix2[ic], n3 = ix3[ic];
          double e1 = ipowi(n1, p1), e2 = ipowi(n2, p2), e3 = ipowi(n3, p3);
          double d1 = p1 * ipowi(n1, p1m), d2 = p2 * ipowi(n2, p2m), d3 = p3 * ipowi(n3, p3m);
          a(ic, col) = e1 * e2 * e3;
          a(ic + ncorn, col) = d1 * e2 * e3;
          a(ic + 2 * ncorn, col) = e1 * d2 * e3;
          a(ic + 3 * ncorn, col) = e1 * e2 * d3;
          a(ic + 4 * ncorn, col) = d1 * d2 * e3;
          a(ic + 5 * ncorn, col) = d1 * e2 * d3;
          a(ic + 6 * ncorn, col) = e1 * d2 * d3;
          a(ic + 7 * ncorn, col) = d1 * d2 * d3;
        }
  int nboxs = nx * ny * nz;
  std::vector<double> B((size_t)npara * nboxs, 0.0);
  auto v3 = [&](const double *t, int i, int j, int k) {
    return t[i + (nx + 1) * (j + (size_t)(ny + 1) * k)];
  };
  for (int ni = 0; ni < nx; ni++)
    for (int nj = 0; nj < ny; nj++)
      for (int nc = 0; nc < nz; nc++) {
        int col = nx * (ny * nc + nj) + ni;
        for (int ic = 0; ic < ncorn; ic++) {
          int n1 = ix1[ic] + ni, n2 = ix2[ic] + nj, n3 = ix3[ic] + nc;
          B[ic + (size_t)npara * col] = v3(values, n1, n2, n3);
          if (dvdx) B[ic + ncorn + (size_t)npara * col] = v3(dvdx, n1, n2, n3);
          if (dvdy) B[ic + 2 * ncorn + (size_t)npara * col] = v3(dvdy, n1, n2, n3);
          if (dvdz) B[ic + 3 * ncorn + (size_t)npara * col] = v3(dvdz, n1, n2, n3);
        }
      }
  int err = atx_host_gaussn(npara, A.data(), nboxs, B.data());
  if (err) return err;
  for (int ibox = 0; ibox < nboxs; ibox++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++)
          coeff[ibox + (size_t)nboxs * (i + 4 * (j + 4 * k))] =
              B[(16 * i + 4 * j + k) + (size_t)npara * ibox];
  return 0;
}

extern "C" int atx_host_rebo2_g_spline(const double *th, const double *g1, const double *dg1,
                                       const double *d2g1, const double *g2, double *g1c,
                                       double *g2c) {
  double A[36], As[36], B[6];
  auto a = [&](double *M, int r, int c) -> double & { return M[r + 6 * c]; };
  // third interval
  std::memset(A, 0, sizeof(A));
  for (int i = 3; i <= 6; i++) {
    double z = th[i - 1];
    for (int j = 1; j <= 6; j++) a(A, i - 3, j - 1) = std::pow(z, j - 1);
  }
  double z = th[2];
  a(A, 4, 1) = 1.0;
  a(A, 5, 2) = 2.0;
  for (int j = 3; j <= 6; j++) {
    a(A, 4, j - 1) = (j - 1) * std::pow(z, j - 2);
    if (j >= 4) a(A, 5, j - 1) = (j - 2) * (j - 1) * std::pow(z, j - 3);
  }
  std::memcpy(As, A, sizeof(A));
  for (int i = 0; i < 4; i++) B[i] = g1[2 + i];
  B[4] = dg1[2];
  B[5] = d2g1[2];
  int err = atx_host_gaussn(6, A, 1, B);
  if (err) return err;
  for (int i = 0; i < 6; i++) g1c[i + 6 * 2] = B[i];
  std::memcpy(A, As, sizeof(A));
  for (int i = 0; i < 4; i++) B[i] = g2[2 + i];
  B[4] = dg1[2];
  B[5] = d2g1[2];
  err = atx_host_gaussn(6, A, 1, B);
  if (err) return err;
  for (int i = 0; i < 6; i++) g2c[i + 6 * 2] = B[i];
  // first and second interval
  for (int k = 0; k <= 1; k++) {
    std::memset(A, 0, sizeof(A));
    for (int i = 0; i <= 1; i++) {
      double zz = th[k] * (1 - i) + th[1 + k] * i;
      a(A, 3 * i, 0) = 1.0;
      a(A, 3 * i + 1, 1) = 1.0;
      a(A, 3 * i + 2, 2) = 2.0;
      for (int j = 2; j <= 6; j++) {
        a(A, 3 * i, j - 1) = std::pow(zz, j - 1);
        if (j >= 3) a(A, 3 * i + 1, j - 1) = (j - 1) * std::pow(zz, j - 2);
        if (j >= 4) a(A, 3 * i + 2, j - 1) = (j - 2) * (j - 1) * std::pow(zz, j - 3);
      }
    }
    B[0] = g1[k]; B[1] = dg1[k]; B[2] = d2g1[k];
    B[3] = g1[1 + k]; B[4] = dg1[1 + k]; B[5] = d2g1[1 + k];
    err = atx_host_gaussn(6, A, 1, B);
    if (err) return err;
    for (int i = 0; i < 6; i++) {
      g1c[i + 6 * k] = B[i];
      g2c[i + 6 * k] = B[i];
    }
  }
  return 0;
}
