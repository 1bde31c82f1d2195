// vecops.cuh -- the vector half of the CG iteration: dot, calc_xr, calc_p
// (reference: CSR/CPUContext.cpp:82-90, 92-105, 107-113).
//
// Pure HBM streams.  Element arithmetic is written with explicit round-to-nearest
// multiply and add (no FMA contraction) so x, r and p match the CPU target bit
// for bit given the same scalars; only the order of the reductions differs
// (fixed tree here, sequential there).
//
// Each kernel exists in two flavours selected by DEV: scalars passed by value
// from the host (the CGContext API: alpha, beta come back from cg.cpp), or read
// from device slots (the device-resident loop, abft_cuda_cg_solve).
#pragma once
#include "common.cuh"

namespace abft {

struct ScalarArg {
  double value;          // host flavour
  const double *num;     // device flavour: value = *num / *den
  const double *den;
};

template <bool DEV>
__device__ __forceinline__ double scalar_of(const ScalarArg &s) {
  if (DEV) return __ddiv_rn(*s.num, *s.den);
  return s.value;
}

// grid-stride over pairs of doubles (128-bit accesses) when the pointers allow it;
// the odd tail, or everything when a pointer is only 8-byte aligned, goes scalar
__global__ void __launch_bounds__(kBlock) dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int N,
                                                      int vec_ok, ReduceOut red, const int *__restrict__ active) {
  __shared__ double s_warp[kBlock / 32];
  if (active && *active == 0) return;
  const int n2 = vec_ok ? (N >> 1) : 0;  // vec_ok: every pointer is 16-byte aligned
  const double2 *a2 = reinterpret_cast<const double2 *>(a);
  const double2 *b2 = reinterpret_cast<const double2 *>(b);
  double acc = 0.0;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n2; i += gridDim.x * kBlock) {
    const double2 u = a2[i], v = b2[i];
    acc = __dadd_rn(acc, __dmul_rn(u.x, v.x));
    acc = __dadd_rn(acc, __dmul_rn(u.y, v.y));
  }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock)
    acc = __dadd_rn(acc, __dmul_rn(a[i], b[i]));
  const double mine = block_sum(acc, s_warp);
  grid_sum_finalize(mine, red, s_warp);
}

// x += alpha p ; r -= alpha w ; sum r^2     (one pass: 32 B read + 16 B written per row)
template <bool DEV>
__global__ void __launch_bounds__(kBlock)
calc_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
               const double *__restrict__ w, ScalarArg alpha_arg, int N, int vec_ok, ReduceOut red,
               const int *__restrict__ active) {
  __shared__ double s_warp[kBlock / 32];
  if (active && *active == 0) {
    // converged: this iteration is a no-op, and so is the next one
    if (red.active && blockIdx.x == 0 && threadIdx.x == 0) *red.active = 0;
    return;
  }
  const double alpha = scalar_of<DEV>(alpha_arg);
  const int n2 = vec_ok ? (N >> 1) : 0;
  double2 *x2 = reinterpret_cast<double2 *>(x);
  double2 *r2 = reinterpret_cast<double2 *>(r);
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  const double2 *w2 = reinterpret_cast<const double2 *>(w);
  double acc = 0.0;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n2; i += gridDim.x * kBlock) {
    double2 xv = x2[i], rv = r2[i];
    const double2 pv = p2[i], wv = w2[i];
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    rv.x = __dsub_rn(rv.x, __dmul_rn(alpha, wv.x));
    rv.y = __dsub_rn(rv.y, __dmul_rn(alpha, wv.y));
    x2[i] = xv;
    r2[i] = rv;
    acc = __dadd_rn(acc, __dmul_rn(rv.x, rv.x));
    acc = __dadd_rn(acc, __dmul_rn(rv.y, rv.y));
  }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) {
    const double xn = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
    const double rn = __dsub_rn(r[i], __dmul_rn(alpha, w[i]));
    x[i] = xn;
    r[i] = rn;
    acc = __dadd_rn(acc, __dmul_rn(rn, rn));
  }
  const double mine = block_sum(acc, s_warp);
  grid_sum_finalize(mine, red, s_warp);
}

// p = r + beta p      (16 B read + 8 B written per row)
template <bool DEV>
__global__ void __launch_bounds__(kBlock)
calc_p_kernel(double *__restrict__ p, const double *__restrict__ r, ScalarArg beta_arg, int N, int vec_ok,
              const int *__restrict__ active) {
  if (active && *active == 0) return;
  const double beta = scalar_of<DEV>(beta_arg);
  const int n2 = vec_ok ? (N >> 1) : 0;
  double2 *p2 = reinterpret_cast<double2 *>(p);
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n2; i += gridDim.x * kBlock) {
    double2 pv = p2[i];
    const double2 rv = r2[i];
    pv.x = __dadd_rn(rv.x, __dmul_rn(beta, pv.x));
    pv.y = __dadd_rn(rv.y, __dmul_rn(beta, pv.y));
    p2[i] = pv;
  }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock)
    p[i] = __dadd_rn(r[i], __dmul_rn(beta, p[i]));
}

}  // namespace abft
