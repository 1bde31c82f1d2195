// cuda_runtime.h -- HOST SHIM for tests/helpers/coo_emul.cpp (test infrastructure).
// Lets g++ compile the device code of csrc/common.cuh + csrc/coo.cuh as plain C++ so that the kernels'
// LOGIC can be executed thread by thread on the CPU (no GPU in the build container) and compared with the
// pinned oracle.  It says nothing about races, memory ordering or performance.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

using std::max;
using std::min;

static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __brev(uint32_t x) {
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31u;
  return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31u;
  return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
static inline double __dadd_rn(double a, double b) { return a + b; }   // built with -ffp-contract=off
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline long long __double_as_longlong(double v) { long long b; memcpy(&b, &v, 8); return b; }
static inline double __longlong_as_double(long long b) { double v; memcpy(&v, &b, 8); return v; }
static inline int __double2loint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)b; }
static inline int __double2hiint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double v; memcpy(&v, &b, 8); return v;
}
template <class T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
