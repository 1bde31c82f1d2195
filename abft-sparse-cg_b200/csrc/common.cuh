// common.cuh -- shared device-side pieces: event log, deterministic reductions,
// cache-hinted loads.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abft_cuda.h"

namespace abft {

constexpr int kBlock = 256;            // threads per CTA of every kernel here
constexpr int kLogCap = 16384;         // device-side fault records kept between polls (more: ABFT_EV_LOG_OVERFLOW)
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
  __threadfence_system();  // (rare path) the host may learn of the kernel's end by polling a word, not by a stream sync
}

// ---- phase stamps (off unless abft_cuda_phases_enable was called: the pointer is null) --------------------------------
// Where a partitioned CG iteration spends its time besides the streams: one 64-bit globaltimer value (ns, one clock for
// all GPUs of a box) per event, written by a single thread.  A row of kPhaseSlots stamps per iteration:
enum {
  PH_SPMV_BEGIN = 0,      // spmv: CTA 0 starts
  PH_SPMV_WAIT_END = 1,   // spmv: CTA 0 has seen every peer's push of p (== PH_SPMV_BEGIN + wait)
  PH_SPMV_REDUCE = 2,     // spmv: the last CTA enters the cross-rank all-reduce of p.w (its own rows are done)
  PH_SPMV_END = 3,        // spmv: all ranks' partials are in, p.w published
  PH_XR_BEGIN = 4,        // calc_xr: CTA 0 starts (after the host round trip on the plugin path)
  PH_XR_REDUCE = 5,       // calc_xr: the last CTA enters the all-reduce of r.r
  PH_XR_END = 6,          // calc_xr: r.r published
  PH_P_BEGIN = 7,         // calc_p (+ push): CTA 0 starts
  PH_P_END = 8,           // calc_p: every store to the peers has landed, flags raised
  kPhaseSlots = 12
};
__device__ __forceinline__ void phase_stamp(unsigned long long *row, int slot) {
#ifndef ABFT_HOST_EMU
  if (row) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    row[slot] = t;
  }
#else
  (void)row; (void)slot;
#endif
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

// all threads of the CTA (BLOCK of them) must call; result valid in thread 0
template <int BLOCK = kBlock>
__device__ __forceinline__ double block_sum(double v, double *s_warp /* [BLOCK/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < BLOCK / 32) ? s_warp[lane] : 0.0;
#pragma unroll
    for (int o = (BLOCK / 64); o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;
}

// ---- multi-GPU exchange window (one per rank, device memory, mapped into every peer) -------
// Ranks PUSH into each other's windows / vectors with plain stores over NVLink (peer memory via
// CUDA IPC or same-process peer access) and only ever READ their own memory.
constexpr int kMaxRanks = 8;

struct CommWindow {
  double slots[2][kMaxRanks];      // [parity of the reduction sequence number][source rank]; kSlotEmpty when consumed
  unsigned int pflag[kMaxRanks];   // last vector-push epoch completed by each source rank
  unsigned int push_counter;       // local: CTAs of the running push kernel that are done
  unsigned int pad[23];
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
  return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const double *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Cross-rank part of a reduction: all-reduce (sum, fixed rank order => every rank gets the same
// bits) of one double per rank through the peers' windows.  world == 0 or 1: local only.
struct PeerReduce {
  CommWindow *win[kMaxRanks];  // win[rank] is local
  int world;
  int rank;
  unsigned int seq;            // this reduction's sequence number (same on every rank)
};

// A slot that holds no contribution: a NaN bit pattern no computation produces (a partial that happens to be this
// very NaN is sent as another NaN).  The data word is its own arrival flag: one 8-byte store per peer, no fence and
// no second round of flag stores on the critical path (round 1 stored data, fenced system-wide, then stored flags).
constexpr unsigned long long kSlotEmpty = 0x7FF8ABF7C0DE5EEDull;

// Called by the first `world` lanes of ONE warp together (lane s talks to rank s); every calling lane returns the
// global sum.  Rank r stores its partial into slot [seq & 1][r] of every rank's window (its own included), each lane
// spins on its own slot of the local window, the partials are added in rank order, and the lane empties the slot it
// has read.  A rank can run at most one reduction ahead of the slowest rank (it needs everybody's partial to finish),
// and then it writes the other parity; the slot it will write two reductions later was emptied before its owner
// sent the partial of the reduction in between (program order + the kernel boundary between two reductions).
__device__ __forceinline__ double peer_allreduce(double mine, const PeerReduce &pr, int lane) {
  const int par = (int)(pr.seq & 1u);
  unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
  if (bits == kSlotEmpty) bits ^= 1ull;
  const unsigned int peers = (1u << pr.world) - 1u;
  *reinterpret_cast<volatile unsigned long long *>(&pr.win[lane]->slots[par][pr.rank]) = bits;
  double *slot = &pr.win[pr.rank]->slots[par][lane];
  unsigned long long got;
  while ((got = ld_volatile_u64(slot)) == kSlotEmpty) {
  }
  *reinterpret_cast<volatile unsigned long long *>(slot) = kSlotEmpty;
  const double v = __longlong_as_double((long long)got);
  double total = 0.0;
  for (int s = 0; s < pr.world; s++) total += __shfl_sync(peers, v, s);
  return total;
}

// Where a reduction's final value goes and what the device-resident CG loop
// derives from it (all optional).
struct ReduceOut {
  double *partials;        // [>= gridDim.x]
  unsigned int *counter;   // zero before launch; reset by the last CTA
  double *dev;             // device scalar slot
  volatile double *host;   // mapped pinned copy of the slot, polled by the host: kSlotEmpty until the value lands (null: not wanted)
  // device-resident loop control (null when driven from the host)
  int *active;             // set to (value > threshold && *iters < max_itrs) by the finalizer
  int *iters;
  double threshold;
  int max_itrs;
  int bump_iters;          // 1: this reduction closes an iteration (calc_xr)
  PeerReduce peers;        // world > 1: the value published is the sum over all ranks
  unsigned long long *phase;  // this iteration's row of phase stamps (null: off)
  int phase_slot;             // PH_SPMV_REDUCE or PH_XR_REDUCE: the stamp taken on entering the all-reduce (+1 on leaving)
};

// Every CTA contributes `mine` (valid in thread 0); the CTA that arrives last
// adds the partials in index order with a fixed tree and publishes the value.
// n_given != 0: the CTAs have already stored o.partials[0 .. n_given) themselves (spmv's per-tile partials) and
// `mine` is ignored.  Returns true in every thread of the CTA that did the final sum.
template <int BLOCK = kBlock>
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
  for (unsigned int i = threadIdx.x; i < n_partials; i += BLOCK) acc += __ldcg(o.partials + i);
  double total = block_sum<BLOCK>(acc, s_warp);
  if (threadIdx.x == 0) phase_stamp(o.phase, o.phase_slot);
  if (o.peers.world > 1) {
    // the first `world` lanes of warp 0 exchange the partials with the peers (lane s <-> rank s)
    if (threadIdx.x < 32) {
      total = __shfl_sync(0xffffffffu, total, 0);
      if ((int)threadIdx.x < o.peers.world) total = peer_allreduce(total, o.peers, (int)threadIdx.x);
    }
  }
  if (threadIdx.x == 0) {
    phase_stamp(o.phase, o.phase_slot + 1);
    *o.dev = total;
    if (o.host) {
      // The host polls this very word (it holds kSlotEmpty until now): one 8-byte store over PCIe, no flag and no
      // system-wide fence on the way -- round 2's phase stamps put ~7 us between two dependent kernels of which two
      // fences that waited for posted PCIe writes were about half.
      unsigned long long bits = (unsigned long long)__double_as_longlong(total);
      if (bits == kSlotEmpty) bits ^= 1ull;
      *reinterpret_cast<volatile unsigned long long *>(const_cast<double *>(o.host)) = bits;
    }
    if (o.active) {
      int it = *o.iters + (o.bump_iters ? 1 : 0);
      if (o.bump_iters) *o.iters = it;
      *o.active = (total > o.threshold && it < o.max_itrs) ? 1 : 0;
    }
    *o.counter = 0u;
  }
  return true;
}

}  // namespace abft
