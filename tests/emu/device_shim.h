// Host stand-ins for the few CUDA device facilities that the per-atom headers of
// atomistica_b200/csrc use, so that a host compiler can build the SAME source and the test-suite can
// run it serially against the CPU oracle (tests/test_emu_rebo2_scr.py).  Test infrastructure only;
// nothing here is part of the product.  Mat3 / shift packing restate atx_internal.cuh:120-226.
#pragma once
#include <cmath>
#include <cstddef>

#define __device__
#define __forceinline__ inline
#define __ldg(p) (*(p))
#define __align__(n) alignas(n)

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
using std::sqrt; using std::exp; using std::pow;

struct Mat3 { double m[9]; };
#define ATX_SHIFT_BIAS 128
#define ATX_SHIFT_ZERO (ATX_SHIFT_BIAS | (ATX_SHIFT_BIAS << 8) | (ATX_SHIFT_BIAS << 16))
#define ATX_SHIFT_MASK 0xFFFFFF
#define ATX_NONZERO_SHIFT(p) (((p) & ATX_SHIFT_MASK) != ATX_SHIFT_ZERO)
static inline int atx_pack_shift(int sx, int sy, int sz) {
  return (sx + ATX_SHIFT_BIAS) | ((sy + ATX_SHIFT_BIAS) << 8) | ((sz + ATX_SHIFT_BIAS) << 16);
}
static inline void atx_unpack_shift(int p, int &sx, int &sy, int &sz) {
  sx = (p & 255) - ATX_SHIFT_BIAS;
  sy = ((p >> 8) & 255) - ATX_SHIFT_BIAS;
  sz = ((p >> 16) & 255) - ATX_SHIFT_BIAS;
}
// built with -ffp-contract=off: same association order, no FMA
static inline void atx_image_vector(const Mat3 &A, int sx, int sy, int sz, double &ax, double &ay, double &az) {
  double s0 = (double)sx, s1 = (double)sy, s2 = (double)sz;
  ax = (A.m[0] * s0 + A.m[3] * s1) + A.m[6] * s2;
  ay = (A.m[1] * s0 + A.m[4] * s1) + A.m[7] * s2;
  az = (A.m[2] * s0 + A.m[5] * s1) + A.m[8] * s2;
}
// one host thread runs the atoms one after the other: plain read-modify-write
#define ATX_NSUM 10
#define RBS_ADD(p, v) (*(p) += (v))
#define RBS_OR(p, v) (*(p) |= (v))
