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
#include "ecc.cuh"

namespace abft {

struct ScalarArg {
  double value;          // host flavour
  const double *num;     // device flavour: value = *num / *den
  const double *den;
};

// VSED (mode secdedv, ecc.cuh): every element read is parity-checked (an odd one is reported: detect-only, fatal) and
// every element written carries its parity in the lowest mantissa bit.  `which` names the vector in the report.
template <bool VSED>
__device__ __forceinline__ void vchk(double v, int index, int which, const LogRef &L) {
  if (VSED && (vec_parity_bits(v) & 1u)) log_event(L, ABFT_EV_VECTOR_SED, (uint32_t)index, (uint32_t)which);
}
template <bool VSED>
__device__ __forceinline__ double venc(double v) { return VSED ? vec_encode(v) : v; }

template <bool DEV>
__device__ __forceinline__ double scalar_of(const ScalarArg &s) {
  if (DEV) return __ddiv_rn(*s.num, *s.den);
  return s.value;
}

// grid-stride over pairs of doubles (128-bit accesses) when the pointers allow it;
// the odd tail, or everything when a pointer is only 8-byte aligned, goes scalar
template <bool VSED>
__global__ void __launch_bounds__(kBlock) dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int N,
                                                      int vec_ok, ReduceOut red, const int *__restrict__ active, LogRef L) {
  __shared__ double s_warp[kBlock / 32];
  if (active && *active == 0) return;
  const int n2 = vec_ok ? (N >> 1) : 0;  // vec_ok: every pointer is 16-byte aligned
  const double2 *a2 = reinterpret_cast<const double2 *>(a);
  const double2 *b2 = reinterpret_cast<const double2 *>(b);
  double acc = 0.0;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n2; i += gridDim.x * kBlock) {
    const double2 u = a2[i], v = b2[i];
    vchk<VSED>(u.x, 2 * i, VEC_DOT_A, L); vchk<VSED>(u.y, 2 * i + 1, VEC_DOT_A, L);
    vchk<VSED>(v.x, 2 * i, VEC_DOT_B, L); vchk<VSED>(v.y, 2 * i + 1, VEC_DOT_B, L);
    acc = __dadd_rn(acc, __dmul_rn(u.x, v.x));
    acc = __dadd_rn(acc, __dmul_rn(u.y, v.y));
  }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) {
    vchk<VSED>(a[i], i, VEC_DOT_A, L); vchk<VSED>(b[i], i, VEC_DOT_B, L);
    acc = __dadd_rn(acc, __dmul_rn(a[i], b[i]));
  }
  const double mine = block_sum(acc, s_warp);
  grid_sum_finalize(mine, red, s_warp);
}

// x += alpha p ; r -= alpha w ; sum r^2     (one pass: 32 B read + 16 B written per row)
template <bool DEV, bool VSED>
__global__ void __launch_bounds__(kBlock)
calc_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
               const double *__restrict__ w, ScalarArg alpha_arg, int N, int vec_ok, ReduceOut red,
               const int *__restrict__ active, LogRef L) {
  __shared__ double s_warp[kBlock / 32];
  if (active && *active == 0) {
    // converged: this iteration is a no-op, and so is the next one
    if (red.active && blockIdx.x == 0 && threadIdx.x == 0) *red.active = 0;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) phase_stamp(red.phase, PH_XR_BEGIN);
  const double alpha = scalar_of<DEV>(alpha_arg);
  const int n2 = vec_ok ? (N >> 1) : 0;
  double2 *x2 = reinterpret_cast<double2 *>(x);
  double2 *r2 = reinterpret_cast<double2 *>(r);
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  const double2 *w2 = reinterpret_cast<const double2 *>(w);
  double acc = 0.0;
  // Swept from the END of the vectors towards the front: spmv has just written w (and gathered p) front to
  // back, so their tails are what the 126 MB L2 still holds; calc_p then runs front to back again and finds
  // the head of r this kernel wrote last.  (Fixed per kernel, so results do not depend on call history.)
  const int stride = gridDim.x * kBlock;
  const int first = blockIdx.x * kBlock + threadIdx.x;
  if (first < n2)
    for (int i = first + ((n2 - 1 - first) / stride) * stride; i >= first; i -= stride) {
      // x is touched by no other kernel of the iteration, w is dead after this read: stream them through
      // L2 (evict-first) so that r and p, which calc_p and the next spmv want, stay
      double2 xv = __ldcs(x2 + i), rv = r2[i];
      const double2 pv = p2[i], wv = __ldcs(w2 + i);
      if (VSED) {
        vchk<VSED>(xv.x, 2 * i, VEC_X, L); vchk<VSED>(xv.y, 2 * i + 1, VEC_X, L);
        vchk<VSED>(rv.x, 2 * i, VEC_R, L); vchk<VSED>(rv.y, 2 * i + 1, VEC_R, L);
        vchk<VSED>(pv.x, 2 * i, VEC_P, L); vchk<VSED>(pv.y, 2 * i + 1, VEC_P, L);
        vchk<VSED>(wv.x, 2 * i, VEC_W, L); vchk<VSED>(wv.y, 2 * i + 1, VEC_W, L);
      }
      xv.x = venc<VSED>(__dadd_rn(xv.x, __dmul_rn(alpha, pv.x)));
      xv.y = venc<VSED>(__dadd_rn(xv.y, __dmul_rn(alpha, pv.y)));
      rv.x = venc<VSED>(__dsub_rn(rv.x, __dmul_rn(alpha, wv.x)));
      rv.y = venc<VSED>(__dsub_rn(rv.y, __dmul_rn(alpha, wv.y)));
      __stcs(x2 + i, xv);
      r2[i] = rv;
      acc = __dadd_rn(acc, __dmul_rn(rv.x, rv.x));
      acc = __dadd_rn(acc, __dmul_rn(rv.y, rv.y));
    }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) {
    if (VSED) {
      vchk<VSED>(x[i], i, VEC_X, L); vchk<VSED>(r[i], i, VEC_R, L); vchk<VSED>(p[i], i, VEC_P, L); vchk<VSED>(w[i], i, VEC_W, L);
    }
    const double xn = venc<VSED>(__dadd_rn(x[i], __dmul_rn(alpha, p[i])));
    const double rn = venc<VSED>(__dsub_rn(r[i], __dmul_rn(alpha, w[i])));
    x[i] = xn;
    r[i] = rn;
    acc = __dadd_rn(acc, __dmul_rn(rn, rn));
  }
  const double mine = block_sum(acc, s_warp);
  grid_sum_finalize(mine, red, s_warp);
}

// p = r + beta p      (16 B read + 8 B written per row)
template <bool DEV, bool VSED>
__global__ void __launch_bounds__(kBlock)
calc_p_kernel(double *__restrict__ p, const double *__restrict__ r, ScalarArg beta_arg, int N, int vec_ok,
              const int *__restrict__ active, LogRef L) {
  if (active && *active == 0) return;
  const double beta = scalar_of<DEV>(beta_arg);
  const int n2 = vec_ok ? (N >> 1) : 0;
  double2 *p2 = reinterpret_cast<double2 *>(p);
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n2; i += gridDim.x * kBlock) {
    double2 pv = p2[i];
    const double2 rv = r2[i];
    if (VSED) {
      vchk<VSED>(pv.x, 2 * i, VEC_P, L); vchk<VSED>(pv.y, 2 * i + 1, VEC_P, L);
      vchk<VSED>(rv.x, 2 * i, VEC_R, L); vchk<VSED>(rv.y, 2 * i + 1, VEC_R, L);
    }
    pv.x = venc<VSED>(__dadd_rn(rv.x, __dmul_rn(beta, pv.x)));
    pv.y = venc<VSED>(__dadd_rn(rv.y, __dmul_rn(beta, pv.y)));
    p2[i] = pv;
  }
  for (int i = 2 * n2 + blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) {
    if (VSED) { vchk<VSED>(p[i], i, VEC_P, L); vchk<VSED>(r[i], i, VEC_R, L); }
    p[i] = venc<VSED>(__dadd_rn(r[i], __dmul_rn(beta, p[i])));
  }
}

// mode secdedv, host -> device: give every element of a freshly uploaded vector its parity bit (unmap_vector) ...
__global__ void __launch_bounds__(kBlock) vec_encode_kernel(double *__restrict__ v, int N) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) v[i] = vec_encode(v[i]);
}
// ... and device -> host: check a vector that is about to be mapped (map_vector)
__global__ void __launch_bounds__(kBlock) vec_check_kernel(const double *__restrict__ v, int N, LogRef L) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < N; i += gridDim.x * kBlock) vchk<true>(v[i], i, VEC_MAPPED, L);
}

// ---- multi-GPU: p = r + beta p fused with the PUSH of the freshly computed entries into the
// peers' copies of p over NVLink (plain stores to peer memory), then one flag per peer.
// Which rows go where is the exchange plan built at setup from the partition's column
// ranges: a banded matrix pushes a thin halo, an irregular one its whole slice (all-gather).
struct PushPlan {
  double *peer_vec[kMaxRanks];  // peers' full-length vectors, indexed by GLOBAL row
  CommWindow *win[kMaxRanks];   // win[rank] is local
  int lo[kMaxRanks], hi[kMaxRanks];  // global rows [lo, hi) of MY slice that peer s reads
  int world, rank;
  int row0;            // global row of local element 0
  unsigned int epoch;  // published in the peers' pflag[rank] once every store has landed
  unsigned long long *phase;  // this iteration's row of phase stamps (null: off)
};

template <bool DEV, bool COMPUTE, bool VSED>
__global__ void __launch_bounds__(kBlock)
calc_p_push_kernel(double *__restrict__ p, const double *__restrict__ r, ScalarArg beta_arg, int N, PushPlan plan,
                   const int *__restrict__ active, LogRef L) {
  if (active && *active == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) phase_stamp(plan.phase, PH_P_BEGIN);
  const double beta = COMPUTE ? scalar_of<DEV>(beta_arg) : 0.0;
  constexpr int kPer = 8;
  constexpr int kChunk = kBlock * kPer;
  for (long long base = (long long)blockIdx.x * kChunk; base < N; base += (long long)gridDim.x * kChunk) {
    double v[kPer];
#pragma unroll
    for (int u = 0; u < kPer; u++) {
      const long long i = base + u * kBlock + threadIdx.x;
      v[u] = 0.0;
      if (i < N) {
        if (COMPUTE) {
          if (VSED) { vchk<VSED>(p[i], (int)i, VEC_P, L); vchk<VSED>(r[i], (int)i, VEC_R, L); }
          v[u] = venc<VSED>(__dadd_rn(r[i], __dmul_rn(beta, p[i])));
          p[i] = v[u];
        } else {
          v[u] = p[i];
        }
      }
    }
    const long long g0 = plan.row0 + base;
    const long long g1 = min(g0 + (long long)kChunk, (long long)plan.row0 + N);
    for (int s = 0; s < plan.world; s++) {
      const long long lo = max((long long)plan.lo[s], g0), hi = min((long long)plan.hi[s], g1);
      if (s == plan.rank || lo >= hi) continue;  // uniform over the CTA
      double *dst = plan.peer_vec[s];
#pragma unroll
      for (int u = 0; u < kPer; u++) {
        const long long g = g0 + u * kBlock + threadIdx.x;
        if (g >= lo && g < hi) dst[g] = v[u];
      }
    }
  }
  // completion: every thread fences its peer stores, the CTA that finishes last raises the flags
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    CommWindow *me = plan.win[plan.rank];
    const unsigned int t = atomicAdd(&me->push_counter, 1u);
    if (t == gridDim.x - 1) {
      me->push_counter = 0u;
      __threadfence_system();
      for (int s = 0; s < plan.world; s++)
        if (s != plan.rank && plan.hi[s] > plan.lo[s])
          *reinterpret_cast<volatile unsigned int *>(&plan.win[s]->pflag[plan.rank]) = plan.epoch;
      phase_stamp(plan.phase, PH_P_END);
    }
  }
}

}  // namespace abft
