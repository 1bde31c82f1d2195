// coo_stream.cuh -- the grouped-layout COO spmv (coo.cuh) with the element stream staged through shared memory by
// cp.async.bulk, the way the CSR kernel stages its three arrays (spmv.cuh).
//
// coo_spmv_kernel<MODE, true> lets thread c read its group straight from global memory: consecutive 16-byte elements
// per thread, i.e. a warp's load touches 32 different 112-byte runs (7-point matrix) and lives off the L1 for the
// rest of them.  Here a tile is 256 consecutive output entries (groups) = one contiguous run of `elements` plus 257
// entries of `colptr`; one elected thread streams both into a two-stage ring (mbarrier completion, L2 evict-first:
// the stream is read once per spmv and must not displace vec[]), thread c then walks its group with 128-bit
// shared-memory loads (stride 7 x 16 B: conflict-free).  Persistent grid, tiles dealt round-robin.  Everything
// observable is coo_spmv_kernel's: same checks, same repairs written back to global memory, same events, same
// stray list, same order of additions.  dot(vec, result) is reduced in the same pass (the CSR kernel's fused dot): the
// abft_cuda_dot(p, w) that follows abft_cuda_coo_spmv(A, p, w) in the CG loop launches nothing.
//
// A tile whose groups hold more than kCooTileCap elements (a dense column) is not staged: its threads read global memory
// like coo_spmv_kernel.
#pragma once
#include "coo.cuh"
#include "spmv.cuh"   // mbarrier / bulk-copy helpers

namespace abft {

constexpr int kCooTileCols = kBlock;     // output entries per tile, one per thread
constexpr int kCooTileCap = 2048;        // elements a stage holds (32 KB); 256 groups of a 7-point matrix: 1792
constexpr int kCooStages = 2;
constexpr int kCooBatch = 8;             // elements of a group in flight at a time
constexpr uint32_t kCooColptrPad = 8;    // colptr is allocated with this many extra entries (the window is rounded up to 16 B)

struct CooStage {
  uint4 el[kCooTileCap + kCooBatch];   // (the fast walk reads up to kCooBatch - 1 slots past a group's end)
  uint32_t cptr[kCooTileCols + 8];
};
constexpr size_t kCooDynSmem = sizeof(CooStage) * kCooStages;

// LEN slots of a group from the staged stream, added to `sum` in element order: all loads and gathers, then the
// checks, which only OR their verdict into `bad` -- a check of the mode trips (raw popcounts, bit 0), the element's
// column is not this group's, its row index is out of range (any bit) -- then the products.  No branch, no call.
// FULL: all LEN slots are the group's.  Otherwise only the first `count`: the slots past the group's end hold the
// next group's elements (or, past the tile's end, whatever the stage held before) -- loaded and gathered (the row
// index clamped), never counted.  A group that comes out `bad` is walked again, one element at a time.
template <int MODE, int LEN, bool FULL>
__device__ __forceinline__ double coo_walk_from(double sum, const uint4 *el, uint32_t count, uint32_t c, uint32_t N,
                                                const double *__restrict__ vec, uint32_t &bad) {
  uint4 e[LEN];
  double xv[LEN];
  uint32_t pops = 0, bits = 0;
#pragma unroll
  for (int j = 0; j < LEN; j++) {
    e[j] = el[j];
    const uint32_t row = min(e[j].y, N - 1u);
    xv[j] = vec[row];
    const uint32_t off = (coo_out_index<MODE>(e[j]) ^ c) | (row ^ e[j].y);   // non-zero: not this group's / row out of range
    bits |= (FULL || (uint32_t)j < count) ? off : 0u;
  }
#pragma unroll
  for (int j = 0; j < LEN; j++) {
    const uint32_t p = coo_suspect<MODE>(e[j]);
    pops |= (FULL || (uint32_t)j < count) ? p : 0u;
  }
  bad |= (pops & 1u) | (bits != 0u ? 1u : 0u);
  if (FULL) {
#pragma unroll
    for (int j = 0; j < LEN; j++) sum = __dadd_rn(sum, __dmul_rn(coo_value(e[j]), xv[j]));
  } else {
#pragma unroll
    for (int j = 0; j < LEN; j++) sum = add_if(sum, __dmul_rn(coo_value(e[j]), xv[j]), (uint32_t)j < count);
  }
  return sum;
}
template <int MODE, int LEN, bool FULL>
__device__ __forceinline__ double coo_walk(const uint4 *el, uint32_t count, uint32_t c, uint32_t N, const double *__restrict__ vec,
                                           uint32_t &bad) {
  return coo_walk_from<MODE, LEN, FULL>(0.0, el, count, c, N, vec, bad);
}
// (groups longer than a batch: the next kCooBatch slots, continuing `sum`)
template <int MODE>
__device__ __forceinline__ double coo_walk_more(double sum, const uint4 *el, uint32_t count, uint32_t c, uint32_t N,
                                                const double *__restrict__ vec, uint32_t &bad) {
  return coo_walk_from<MODE, kCooBatch, false>(sum, el, count, c, N, vec, bad);
}

template <int MODE>
__global__ void __launch_bounds__(kBlock, 3)
coo_spmv_stream_kernel(CooView A, const double *__restrict__ vec, double *__restrict__ result, LogRef L, ReduceOut red) {
#ifndef ABFT_HOST_EMU
  extern __shared__ __align__(128) unsigned char dyn_smem[];
#else
  unsigned char *const dyn_smem = emu::g_dyn_smem;
#endif
  CooStage *stage = reinterpret_cast<CooStage *>(dyn_smem);
  __shared__ __align__(8) uint64_t s_full[kCooStages];
  __shared__ uint32_t s_kb[kCooStages];       // first element of the staged tile
  __shared__ uint32_t s_staged[kCooStages];   // 0: the tile's elements were too many to stage
  __shared__ double s_warp[kBlock / 32];

  const int tid = threadIdx.x;
  const uint32_t G = gridDim.x;
  const uint32_t n_tiles = (A.N + kCooTileCols - 1) / kCooTileCols;
  const uint32_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + G - 1) / G : 0u;

  if (tid == 0) {
    for (int s = 0; s < kCooStages; s++) mbar_init(&s_full[s], 1);
    fence_mbarrier_init();
  }
  __syncthreads();
  const uint64_t stream_policy = l2_evict_first_policy();

  // producer (thread 0): the element range of tile `tile`, and its stream into stage s
  auto bounds = [&](uint32_t tile, uint32_t &kb, uint32_t &ke) {
    const uint32_t c0 = tile * kCooTileCols;
    kb = A.colptr[c0];
    ke = A.colptr[min(c0 + (uint32_t)kCooTileCols, A.N)];
  };
  auto issue = [&](int s, uint32_t tile, uint32_t kb, uint32_t ke) {
    const uint32_t c0 = tile * kCooTileCols;
    const uint32_t ncols = min((uint32_t)kCooTileCols, A.N - c0);
    const uint32_t cbytes = ((ncols + 1 + 3) & ~3u) * 4u;
    const uint32_t n = ke - kb;
    const bool staged = n <= (uint32_t)kCooTileCap;
    s_kb[s] = kb;
    s_staged[s] = staged ? 1u : 0u;
    fence_proxy_async();
    mbar_expect_tx(&s_full[s], cbytes + (staged ? n * 16u : 0u));
    bulk_g2s(stage[s].cptr, A.colptr + c0, cbytes, &s_full[s], stream_policy);
    if (staged && n) bulk_g2s(stage[s].el, A.elements + kb, n * 16u, &s_full[s], stream_policy);
  };

  uint32_t nkb = 0, nke = 0;   // thread 0: bounds of the tile to stream next, fetched one tile ahead
  if (tid == 0) {
    for (uint32_t i = 0; i < (uint32_t)kCooStages && i < n_my; i++) {
      bounds(blockIdx.x + i * G, nkb, nke);
      issue((int)i, blockIdx.x + i * G, nkb, nke);
    }
    if ((uint32_t)kCooStages < n_my) bounds(blockIdx.x + kCooStages * G, nkb, nke);
  }
  __syncthreads();

  double dot_part = 0.0;   // this thread's share of dot(vec, result)
  for (uint32_t it = 0; it < n_my; it++) {
    const int st = (int)(it % kCooStages);
    const uint32_t tile = blockIdx.x + it * G;
    mbar_wait(&s_full[st], (it / kCooStages) & 1u);
    const CooStage &B = stage[st];
    const uint32_t c = tile * kCooTileCols + tid;
    const bool has = c < A.N;                   // (the last tile can have fewer groups than threads)
    const uint32_t k0 = s_kb[st];
    const bool staged = s_staged[st] != 0u;
    const uint32_t kb = has ? B.cptr[tid] : k0, ke = has ? B.cptr[tid + 1] : k0;
    const uint32_t len = ke - kb;
    const double xc = has ? vec[c] : 0.0;       // dot()'s operand: in flight behind the walk
    double sum = 0.0;
    // Fast walk (staged tiles; CONSTRAINTS needs each element's successor in the caller's order and takes the one-by-one
    // walk).  Which one is a warp-uniform vote, so the walk itself has no branch: every group of the warp exactly 7
    // (5) elements long -- the interior of a 3-D (2-D) Poisson matrix -- walks 7 (5) slots with nothing predicated;
    // every group at most 8 long walks 8 slots, the ones past its end not counted; anything else 8 at a time.
    uint32_t bad = (MODE == ABFT_MODE_CONSTRAINTS || !staged) ? 1u : 0u;
    if (!bad) {
      const uint4 *el = B.el + (kb - k0);
      if (__all_sync(0xffffffffu, !has || len == 7u)) sum = coo_walk<MODE, 7, true>(el, 7u, c, A.N, vec, bad);
      else if (__all_sync(0xffffffffu, !has || len == 5u)) sum = coo_walk<MODE, 5, true>(el, 5u, c, A.N, vec, bad);
      else if (__all_sync(0xffffffffu, len <= (uint32_t)kCooBatch)) sum = coo_walk<MODE, kCooBatch, false>(el, len, c, A.N, vec, bad);
      else
        for (uint32_t k = 0; k < len; k += kCooBatch)
          sum = coo_walk_more<MODE>(sum, el + k, len - k, c, A.N, vec, bad);
    }
    if (has) {
      if (bad) {
        // something unusual in this group (or no fast walk for this tile / mode): one element at a time, with the
        // mode's repairs and reports
        if (staged) {
          sum = 0.0;
          for (uint32_t k = kb; k < ke; k++) coo_accumulate<MODE, true>(B.el[k - k0], k, k, c, A, vec, L, sum);
        } else {
          sum = coo_group_from_global<MODE>(A, kb, ke, c, vec, L);
        }
      }
      result[c] = sum;
      dot_part = __dadd_rn(dot_part, __dmul_rn(sum, xc));
    }
    __syncthreads();  // everyone is done with stage `st`: refill it
    if (tid == 0 && it + kCooStages < n_my) {
      issue(st, blockIdx.x + (it + kCooStages) * G, nkb, nke);
      if (it + kCooStages + 1 < n_my) bounds(blockIdx.x + (it + kCooStages + 1) * G, nkb, nke);
    }
  }
  // fused dot(vec, result): CTA partials, then the grid-wide deterministic sum into red.dev.  red.host is null: the
  // stray fix-up kernel that follows may still change entries of result[]; it corrects the sum and hands it to the host.
  const double mine = block_sum<kBlock>(dot_part, s_warp);
  grid_sum_finalize<kBlock>(mine, red, s_warp);
}

}  // namespace abft
