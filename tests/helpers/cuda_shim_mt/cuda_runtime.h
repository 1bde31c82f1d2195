// cuda_runtime.h -- MULTI-THREADED HOST SHIM for tests/helpers/csr_emul.cpp (test infrastructure).
// Lets g++ compile the device code of csrc/{common,ecc,spmv,encode,vecops}.cuh as plain C++ and run a CTA as 256
// real threads: __syncthreads and the full-mask warp collectives are barriers, shared memory is function-local
// statics (one CTA runs at a time), atomics are atomics, and the mbarrier / TMA bulk-copy helpers of spmv.cuh are
// replaced (under ABFT_HOST_EMU) by a small state machine + memcpy.  This executes the kernels' LOGIC -- tile walk,
// checks, repairs, fault log, reductions' shape -- against the pinned oracle without a GPU.  It says nothing about
// memory ordering on the device, launch configuration or performance.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define ABFT_HOST_EMU 1
#define __device__
#define __global__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct alignas(16) double2 { double x, y; };
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;

using std::max;
using std::min;

static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __brev(uint32_t x) {
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
// PRMT: result byte k = byte (selector nibble k) of the 8-byte pool {x (bytes 0..3), y (bytes 4..7)} (no sign modes used)
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t pool = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) r |= (uint32_t)((pool >> (8 * ((s >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31u;
  return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline double __dadd_rn(double a, double b) { return a + b; }   // built with -ffp-contract=off
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline long long __double_as_longlong(double v) { long long b; memcpy(&b, &v, 8); return b; }
static inline double __longlong_as_double(long long b) { double v; memcpy(&v, &b, 8); return v; }
static inline int __double2loint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)b; }
static inline int __double2hiint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double v; memcpy(&v, &b, 8); return v;
}
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, const T &v) { *p = v; }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace emu {

// ---- one CTA = `n` threads with a CTA barrier and one barrier + exchange row per warp ------------------------
struct Cta {
  explicit Cta(int n) : cta(n) {
    for (int w = 0; w < (n + 31) / 32; w++) warp.emplace_back(new std::barrier<>(std::min(32, n - 32 * w)));
    xchg.assign((size_t)((n + 31) / 32) * 32 * 2, 0);  // two buffers, see warp_exchange
  }
  std::barrier<> cta;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<uint64_t> xchg;
};
static Cta *g_cta = nullptr;
static thread_local unsigned t_phase = 0;  // collectives this thread has made in the running CTA (see warp_exchange)

// run `grid` CTAs one after the other, each as `block` real threads (a persistent pool: thread creation would
// otherwise dominate the run time of the tests)
struct Pool {
  std::mutex m;
  std::condition_variable go, done;
  std::vector<std::thread> threads;
  const std::function<void()> *job = nullptr;
  unsigned generation = 0;
  int block = 0, cta_index = 0, remaining = 0;

  void worker(int t) {
    unsigned seen = 0;
    for (;;) {
      const std::function<void()> *f;
      int b;
      {
        std::unique_lock<std::mutex> lk(m);
        go.wait(lk, [&] { return generation != seen; });
        seen = generation;
        if (t >= block) continue;  // this CTA is smaller than the pool
        f = job;
        b = cta_index;
      }
      blockIdx.x = (unsigned)b;
      threadIdx.x = (unsigned)t;
      t_phase = 0;
      (*f)();
      std::lock_guard<std::mutex> lk(m);
      if (--remaining == 0) done.notify_all();
    }
  }
  void run_cta(int b, int n, const std::function<void()> &f) {
    while ((int)threads.size() < n) {
      const int t = (int)threads.size();
      threads.emplace_back([this, t] { worker(t); });
      threads.back().detach();
    }
    std::unique_lock<std::mutex> lk(m);
    job = &f;
    block = n;
    cta_index = b;
    remaining = n;
    generation++;
    go.notify_all();
    done.wait(lk, [&] { return remaining == 0; });
  }
};
static Pool *g_pool = new Pool();  // leaked on purpose: its detached threads outlive static destruction

template <class F>
static void launch(int grid, int block, F f) {
  gridDim.x = (unsigned)grid;
  blockDim.x = (unsigned)block;
  const std::function<void()> job = f;
  for (int b = 0; b < grid; b++) {
    Cta cta(block);
    g_cta = &cta;
    g_pool->run_cta(b, block, job);
    g_cta = nullptr;
  }
}

// full-mask warp exchange: everyone publishes, one barrier, everyone reads.  Two buffers used alternately (each
// thread counts its own collectives, and all lanes of a warp make the same ones): a slot is only written again two
// collectives later, i.e. after a barrier that every reader of the old value has passed.
template <class T>
static inline T warp_exchange(T mine, int src_lane_of_me) {
  static_assert(sizeof(T) <= 8, "exchange slot is 8 bytes");
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  uint64_t *buf = g_cta->xchg.data() + (size_t)(t_phase++ & 1u) * g_cta->xchg.size() / 2 + (size_t)warp * 32;
  uint64_t raw = 0;
  memcpy(&raw, &mine, sizeof(T));
  buf[lane] = raw;
  g_cta->warp[warp]->arrive_and_wait();
  T out = mine;
  if (src_lane_of_me >= 0 && src_lane_of_me < 32) memcpy(&out, &buf[src_lane_of_me], sizeof(T));
  return out;
}
static inline uint32_t warp_ballot(bool pred) {
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  uint64_t *buf = g_cta->xchg.data() + (size_t)(t_phase++ & 1u) * g_cta->xchg.size() / 2 + (size_t)warp * 32;
  buf[lane] = pred ? 1u : 0u;
  g_cta->warp[warp]->arrive_and_wait();
  uint32_t m = 0;
  for (int l = 0; l < 32; l++) m |= (uint32_t)(buf[l] & 1u) << l;
  return m;
}

// ---- mbarrier + bulk copy (the contracts of the PTX helpers in spmv.cuh) ----------------------------------------
struct MBar {  // lives in the kernel's uint64_t
  int32_t tx;       // bytes still expected
  int16_t pending;  // arrivals still expected in this phase
  uint8_t phase;    // parity of the phase in progress
  uint8_t count;    // arrivals per phase
};
static_assert(sizeof(MBar) == 8, "fits the mbarrier word");
static std::mutex g_mbar_mutex;
static std::condition_variable g_mbar_cv;
static inline void mbar_settle(MBar *m) {
  if (m->pending == 0 && m->tx == 0) {
    m->phase ^= 1u;
    m->pending = m->count;
    g_mbar_cv.notify_all();
  }
}
static inline void mbar_init(uint64_t *bar, uint32_t count) {
  std::lock_guard<std::mutex> g(g_mbar_mutex);
  MBar m{0, (int16_t)count, 0, (uint8_t)count};
  memcpy(bar, &m, 8);
}
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(g_mbar_mutex);
  MBar *m = reinterpret_cast<MBar *>(bar);
  m->tx += (int32_t)bytes;
  m->pending -= 1;
  mbar_settle(m);
}
static inline void mbar_arrive(uint64_t *bar) {
  std::lock_guard<std::mutex> g(g_mbar_mutex);
  MBar *m = reinterpret_cast<MBar *>(bar);
  m->pending -= 1;
  mbar_settle(m);
}
static inline void mbar_wait(uint64_t *bar, uint32_t parity) {
  std::unique_lock<std::mutex> lk(g_mbar_mutex);
  // the phase with this parity is complete once the barrier has moved on to the other one
  g_mbar_cv.wait(lk, [&] { return reinterpret_cast<MBar *>(bar)->phase != (uint8_t)parity; });
}
static inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> g(g_mbar_mutex);
  MBar *m = reinterpret_cast<MBar *>(bar);
  m->tx -= (int32_t)bytes;
  mbar_settle(m);
}

// dynamic shared memory of the running CTA
static unsigned char *g_dyn_smem = nullptr;

}  // namespace emu

static inline void __syncthreads() { emu::g_cta->cta.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_cta->warp[(int)threadIdx.x >> 5]->arrive_and_wait(); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int delta) {
  const int lane = (int)threadIdx.x & 31;
  return emu::warp_exchange(v, lane + delta < 32 ? lane + delta : lane);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_exchange(v, src & 31); }
static inline uint32_t __ballot_sync(unsigned, bool pred) { return emu::warp_ballot(pred); }
static inline bool __any_sync(unsigned, bool pred) { return emu::warp_ballot(pred) != 0; }
static inline bool __all_sync(unsigned, bool pred) { return emu::warp_ballot(pred) == 0xFFFFFFFFu; }
