// common.cuh -- shared device-side pieces: event log, deterministic reductions,
// cache-hinted loads.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abft_cuda.h"

namespace abft {

constexpr int kBlock = 256;            // threads per CTA of every kernel here
constexpr int kLogCap = 2048;          // device-side fault records kept between polls
constexpr int kNumSlots = 8;           // scalar result slots
constexpr int kMaxPartials = 1 << 20;  // per-CTA partial sums of one reduction

// ---- device fault log ------------------------------------------------------
struct DevLog {
  unsigned int count;      // records appended (may exceed kLogCap; extra ones dropped)
  unsigned int pad;
  abft_event ev[kLogCap];
};

struct LogRef {
  DevLog *log;                   // device memory
  volatile unsigned int *flag;   // mapped pinned host word: non-zero once anything was logged
};

__device__ __forceinline__ void log_event(const LogRef &L, int kind, uint32_t index, uint32_t bit) {
  unsigned int slot = atomicAdd(&L.log->count, 1u);
  if (slot < (unsigned)kLogCap) {
    L.log->ev[slot].kind = kind;
    L.log->ev[slot].index = index;
    L.log->ev[slot].bit = bit;
  }
  *L.flag = 1u;
}

// ---- loads -------------------------------------------------------------------
// matrix stream: read once per spmv, keep it from displacing the vectors in L2
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }

// ---- deterministic sum reductions ----------------------------------------------
// Fixed shape: shuffle tree inside the warp, then warp 0 adds the warp sums in a
// shuffle tree.  No floating-point atomics anywhere, so a given launch geometry
// gives bitwise reproducible results.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// all threads of the CTA must call; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *s_warp /* [kBlock/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < kBlock / 32) ? s_warp[lane] : 0.0;
#pragma unroll
    for (int o = (kBlock / 64); o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;
}

// ---- multi-GPU exchange window (one per rank, device memory, mapped into every peer) -------
// Ranks PUSH into each other's windows / vectors with plain stores over NVLink (peer memory via
// CUDA IPC or same-process peer access) and only ever READ their own memory.
constexpr int kMaxRanks = 8;

struct CommWindow {
  double slots[2][kMaxRanks];      // [parity of the reduction sequence number][source rank]
  unsigned int rflag[kMaxRanks];   // last reduction sequence number published by each source rank
  unsigned int pflag[kMaxRanks];   // last vector-push epoch completed by each source rank
  unsigned int push_counter;       // local: CTAs of the running push kernel that are done
  unsigned int pad[15];
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
  return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
  return *reinterpret_cast<const volatile double *>(p);
}

// Cross-rank part of a reduction: all-reduce (sum, fixed rank order => every rank gets the same
// bits) of one double per rank through the peers' windows.  world == 0 or 1: local only.
struct PeerReduce {
  CommWindow *win[kMaxRanks];  // win[rank] is local
  int world;
  int rank;
  unsigned int seq;            // this reduction's sequence number (same on every rank)
};

// called by ONE thread; returns the global sum
__device__ __forceinline__ double peer_allreduce(double mine, const PeerReduce &pr) {
  const int par = (int)(pr.seq & 1u);
  for (int s = 0; s < pr.world; s++) pr.win[s]->slots[par][pr.rank] = mine;
  __threadfence_system();
  for (int s = 0; s < pr.world; s++) *reinterpret_cast<volatile unsigned int *>(&pr.win[s]->rflag[pr.rank]) = pr.seq;
  const CommWindow *me = pr.win[pr.rank];
  double total = 0.0;
  for (int s = 0; s < pr.world; s++) {
    // a rank can run at most one reduction ahead, and then it writes the other parity
    while ((int)(ld_volatile_u32(&me->rflag[s]) - pr.seq) < 0) {
    }
  }
  __threadfence_system();
  for (int s = 0; s < pr.world; s++) total += ld_volatile_f64(&me->slots[par][s]);
  return total;
}

// Where a reduction's final value goes and what the device-resident CG loop
// derives from it (all optional).
struct ReduceOut {
  double *partials;        // [>= gridDim.x]
  unsigned int *counter;   // zero before launch; reset by the last CTA
  double *dev;             // device scalar slot
  volatile double *host;   // mapped pinned copy of the slot
  // device-resident loop control (null when driven from the host)
  int *active;             // set to (value > threshold && *iters < max_itrs) by the finalizer
  int *iters;
  double threshold;
  int max_itrs;
  int bump_iters;          // 1: this reduction closes an iteration (calc_xr)
  PeerReduce peers;        // world > 1: the value published is the sum over all ranks
};

// Every CTA contributes `mine` (valid in thread 0); the CTA that arrives last
// adds the partials in index order with a fixed tree and publishes the value.
// n_given != 0: the CTAs have already stored o.partials[0 .. n_given) themselves (spmv's per-tile partials) and
// `mine` is ignored.  Returns true in every thread of the CTA that did the final sum.
__device__ __forceinline__ bool grid_sum_finalize(double mine, const ReduceOut &o, double *s_warp, unsigned int n_given = 0) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    if (n_given == 0) o.partials[blockIdx.x] = mine;
    __threadfence();
    unsigned int t = atomicAdd(o.counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc = 0.0;
  const unsigned int n_partials = n_given ? n_given : gridDim.x;
  for (unsigned int i = threadIdx.x; i < n_partials; i += kBlock) acc += __ldcg(o.partials + i);
  double total = block_sum(acc, s_warp);
  if (threadIdx.x == 0) {
    if (o.peers.world > 1) total = peer_allreduce(total, o.peers);
    *o.dev = total;
    if (o.host) *o.host = total;
    if (o.active) {
      int it = *o.iters + (o.bump_iters ? 1 : 0);
      if (o.bump_iters) *o.iters = it;
      *o.active = (total > o.threshold && it < o.max_itrs) ? 1 : 0;
    }
    *o.counter = 0u;
    __threadfence_system();
  }
  return true;
}

}  // namespace abft
