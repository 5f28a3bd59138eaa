/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see atomistica_oracle.h).
 *
 * Cell-list neighbour build, restated from
 *   src/python/f90/python_neighbors.f90:765-871 (binning_init)
 *   src/python/f90/python_neighbors.f90:904-959 (binning_update)
 *   src/python/f90/python_neighbors.f90:570-754 (fill_neighbor_list)
 * Compile with -ffp-contract=off: the reference (gfortran, x86-64 baseline)
 * rounds every multiply and add separately.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "atomistica_oracle.h"

#define M3(M, i, j) (M)[(j) * 3 + (i)] /* Fortran M(i+1,j+1) */

static double dot3(const double *a, const double *b) {
  double s = 0.0;
  for (int i = 0; i < 3; i++) s += a[i] * b[i];
  return s;
}

static void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* python_neighbors.f90:765-871 */
int orc_binning_init(const double *Abox, const double *Bbox, double cutoff, double bin_size,
                     orc_binning_t *b) {
  double cell_size[9], box_size[3];
  for (int x = 0; x < 3; x++) box_size[x] = sqrt(dot3(&Abox[3 * x], &Abox[3 * x]));
  for (int x = 0; x < 3; x++) {
    b->n_cells[x] = (int)(box_size[x] / bin_size);
    if (b->n_cells[x] < 3) b->n_cells[x] = 3;
  }
  for (int x = 0; x < 3; x++)
    for (int i = 0; i < 3; i++) {
      M3(cell_size, i, x) = M3(Abox, i, x) / b->n_cells[x];
      M3(b->rec_cell_size, x, i) = M3(Bbox, x, i) * b->n_cells[x];
    }
  double nx[3], ny[3], nz[3];
  cross3(&cell_size[3], &cell_size[6], nx);
  cross3(&cell_size[6], &cell_size[0], ny);
  cross3(&cell_size[0], &cell_size[3], nz);
  double cv = dot3(&cell_size[0], nx);
  double nxx = dot3(nx, nx), nyy = dot3(ny, ny), nzz = dot3(nz, nz);
  for (int i = 0; i < 3; i++) {
    nx[i] = cv * nx[i] / nxx;
    ny[i] = cv * ny[i] / nyy;
    nz[i] = cv * nz[i] / nzz;
  }
  b->dx = (int)(cutoff / sqrt(dot3(nx, nx))) + 1;
  b->dy = (int)(cutoff / sqrt(dot3(ny, ny))) + 1;
  b->dz = (int)(cutoff / sqrt(dot3(nz, nz))) + 1;
  return 0;
}

/* floor(matmul(rec_cell_size, r)) -- 0-based cell, lower_with_border = 0 in the Python host */
static void cell_of(const double *rec, const double *r, int *c) {
  for (int k = 0; k < 3; k++) {
    double s = 0.0;
    for (int j = 0; j < 3; j++) s += M3(rec, k, j) * r[j];
    c[k] = (int)floor(s);
  }
}

static int orc_nl_threads = 1;

/* OpenMP threads of the list build (bench.py's CPU legs).  The reference's fill_neighbor_list is
 * parallel too (python_neighbors.f90:610-628, one chunk of the list per thread).  Here the
 * threaded path counts per atom, prefix-sums and fills, which yields exactly the serial layout;
 * 1 (default) runs the serial loop below. */
void orc_nl_set_threads(int n) { orc_nl_threads = n > 0 ? n : 1; }

/* neighbours of atom i in the reference's traversal order; fill = 0 only counts, fill = 1 writes
 * them to slots pos, pos+1, ... (0-based) */
static long nl_atom(int i, const double *r, const double *Abox, const int *pbc, const orc_binning_t *b,
                    const int *bin_seed, const int *next, double cutoff_sq, int fill, long pos,
                    int *neighbors, int *dc) {
  const int *n = b->n_cells;
  long cnt = 0;
  int celli[3], shift[3] = {0, 0, 0};
  cell_of(b->rec_cell_size, &r[3 * i], celli);
  for (int k = 0; k < 3; k++) {
    if (pbc[k]) {
      while (celli[k] < 0) { celli[k] += n[k]; shift[k] += 1; }
      while (celli[k] >= n[k]) { celli[k] -= n[k]; shift[k] -= 1; }
    } else {
      if (celli[k] < 0) celli[k] = 0;
      if (celli[k] >= n[k]) celli[k] = n[k] - 1;
    }
  }
  for (int x = -b->dx; x <= b->dx; x++)
    for (int y = -b->dy; y <= b->dy; y++)
      for (int z = -b->dz; z <= b->dz; z++) {
        int cc[3] = {celli[0] + x, celli[1] + y, celli[2] + z};
        int shift1[3] = {shift[0], shift[1], shift[2]};
        int exists = 1;
        for (int k = 0; k < 3; k++) {
          if (pbc[k]) {
            while (cc[k] < 0) { cc[k] += n[k]; shift1[k] += 1; }
            while (cc[k] >= n[k]) { cc[k] -= n[k]; shift1[k] -= 1; }
          }
          if (cc[k] < 0 || cc[k] >= n[k]) exists = 0;
        }
        if (!exists) continue;
        int j = bin_seed[cc[0] + (long)n[0] * (cc[1] + (long)n[1] * cc[2])];
        while (j != -1) {
          int cellj[3], shift2[3] = {shift1[0], shift1[1], shift1[2]};
          cell_of(b->rec_cell_size, &r[3 * j], cellj);
          for (int k = 0; k < 3; k++) {
            if (pbc[k]) {
              while (cellj[k] < 0) { cellj[k] += n[k]; shift2[k] -= 1; }
              while (cellj[k] >= n[k]) { cellj[k] -= n[k]; shift2[k] += 1; }
            }
          }
          if (i != j || shift2[0] != 0 || shift2[1] != 0 || shift2[2] != 0) {
            double d[3];
            for (int k = 0; k < 3; k++) {
              double s = 0.0;
              for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)shift2[c];
              d[k] = r[3 * i + k] - r[3 * j + k] + s;
            }
            if (dot3(d, d) < cutoff_sq) {
              if (fill) {
                neighbors[pos + cnt] = j + 1;
                dc[3 * (pos + cnt) + 0] = shift2[0];
                dc[3 * (pos + cnt) + 1] = shift2[1];
                dc[3 * (pos + cnt) + 2] = shift2[2];
              }
              cnt++;
            }
          }
          j = next[j];
        }
      }
  return cnt;
}

long orc_nl_build(int nat, const double *r, const double *Abox, const double *Bbox, const int *pbc,
                  double cutoff, long capacity, intptr_t *seed, intptr_t *last, int *neighbors,
                  int *dc) {
  orc_binning_t b;
  orc_binning_init(Abox, Bbox, cutoff, cutoff, &b);
  const int *n = b.n_cells;
  long ncell = (long)n[0] * n[1] * n[2];
  int *bin_seed = (int *)malloc(sizeof(int) * ncell);
  int *bin_last = (int *)malloc(sizeof(int) * ncell);
  int *next = (int *)malloc(sizeof(int) * (nat > 0 ? nat : 1));
  for (long c = 0; c < ncell; c++) bin_seed[c] = bin_last[c] = -1;
  for (int i = 0; i < nat; i++) next[i] = -1;

  /* binning_update: append atom i (ascending) to its cell's linked list */
  for (int i = 0; i < nat; i++) {
    int c[3];
    cell_of(b.rec_cell_size, &r[3 * i], c);
    for (int k = 0; k < 3; k++) {
      if (pbc[k]) {
        while (c[k] < 0) c[k] += n[k];
        while (c[k] >= n[k]) c[k] -= n[k];
      } else {
        if (c[k] < 0) c[k] = 0;
        if (c[k] >= n[k]) c[k] = n[k] - 1;
      }
    }
    long ci = c[0] + (long)n[0] * (c[1] + (long)n[1] * c[2]);
    if (bin_seed[ci] == -1) {
      bin_seed[ci] = i;
      bin_last[ci] = i;
    } else {
      next[bin_last[ci]] = i;
      bin_last[ci] = i;
    }
  }

  double cutoff_sq = cutoff * cutoff;
  if (orc_nl_threads > 1) {
    long *cnt = (long *)malloc(sizeof(long) * (nat > 0 ? nat : 1));
#pragma omp parallel for schedule(dynamic, 64) num_threads(orc_nl_threads)
    for (int i = 0; i < nat; i++)
      cnt[i] = nl_atom(i, r, Abox, pbc, &b, bin_seed, next, cutoff_sq, 0, 0, NULL, NULL);
    long cur1 = 1, total = 0;
    for (int i = 0; i < nat; i++) {
      seed[i] = cur1;
      last[i] = cur1 + cnt[i] - 1;
      cur1 += cnt[i] + 1; /* + terminator slot */
      total += cnt[i];
    }
    long result = total;
    if (cur1 - 1 > capacity) {
      result = -1; /* same condition as the serial "cur >= capacity" check */
    } else {
      seed[nat] = cur1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(orc_nl_threads)
      for (int i = 0; i < nat; i++) {
        nl_atom(i, r, Abox, pbc, &b, bin_seed, next, cutoff_sq, 1, seed[i] - 1, neighbors, dc);
        neighbors[last[i]] = 0; /* terminator slot (0-based index last[i]) */
      }
    }
    free(cnt); free(bin_seed); free(bin_last); free(next);
    return result;
  }
  long cur = 1, nn = 0; /* 1-based slot as in the reference */
  int overflow = 0;
  for (int i = 0; i < nat && !overflow; i++) {
    int celli[3], shift[3] = {0, 0, 0};
    cell_of(b.rec_cell_size, &r[3 * i], celli);
    for (int k = 0; k < 3; k++) {
      if (pbc[k]) {
        while (celli[k] < 0) { celli[k] += n[k]; shift[k] += 1; }
        while (celli[k] >= n[k]) { celli[k] -= n[k]; shift[k] -= 1; }
      } else {
        if (celli[k] < 0) celli[k] = 0;
        if (celli[k] >= n[k]) celli[k] = n[k] - 1;
      }
    }
    seed[i] = cur;
    for (int x = -b.dx; x <= b.dx && !overflow; x++)
      for (int y = -b.dy; y <= b.dy && !overflow; y++)
        for (int z = -b.dz; z <= b.dz && !overflow; z++) {
          int cc[3] = {celli[0] + x, celli[1] + y, celli[2] + z};
          int shift1[3] = {shift[0], shift[1], shift[2]};
          int exists = 1;
          for (int k = 0; k < 3; k++) {
            if (pbc[k]) {
              while (cc[k] < 0) { cc[k] += n[k]; shift1[k] += 1; }
              while (cc[k] >= n[k]) { cc[k] -= n[k]; shift1[k] -= 1; }
            }
            if (cc[k] < 0 || cc[k] >= n[k]) exists = 0;
          }
          if (!exists) continue;
          int j = bin_seed[cc[0] + (long)n[0] * (cc[1] + (long)n[1] * cc[2])];
          while (j != -1) {
            int cellj[3], shift2[3] = {shift1[0], shift1[1], shift1[2]};
            cell_of(b.rec_cell_size, &r[3 * j], cellj);
            for (int k = 0; k < 3; k++) {
              if (pbc[k]) {
                while (cellj[k] < 0) { cellj[k] += n[k]; shift2[k] -= 1; }
                while (cellj[k] >= n[k]) { cellj[k] -= n[k]; shift2[k] += 1; }
              }
            }
            if (i != j || shift2[0] != 0 || shift2[1] != 0 || shift2[2] != 0) {
              double d[3], as[3];
              for (int k = 0; k < 3; k++) {
                /* matmul(Abox, shift2) */
                double s = 0.0;
                for (int c = 0; c < 3; c++) s += M3(Abox, k, c) * (double)shift2[c];
                as[k] = s;
              }
              for (int k = 0; k < 3; k++) d[k] = r[3 * i + k] - r[3 * j + k] + as[k];
              double d2 = dot3(d, d);
              if (d2 < cutoff_sq) {
                if (cur >= capacity) { overflow = 1; break; }
                neighbors[cur - 1] = j + 1;
                dc[3 * (cur - 1) + 0] = shift2[0];
                dc[3 * (cur - 1) + 1] = shift2[1];
                dc[3 * (cur - 1) + 2] = shift2[2];
                cur++;
                nn++;
              }
            }
            j = next[j];
          }
        }
    last[i] = cur - 1;
    if (!overflow) {
      neighbors[cur - 1] = 0; /* terminator slot */
      cur++;
    }
  }
  if (!overflow) seed[nat] = cur;
  free(bin_seed);
  free(bin_last);
  free(next);
  return overflow ? -1 : nn;
}
