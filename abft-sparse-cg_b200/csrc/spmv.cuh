// spmv.cuh -- CSR SpMV with the per-element ABFT check/correct fused into the
// column/value stream (reference: CSR/CPUContext.cpp:115-133 NONE, 162-207
// CONSTRAINTS, 214-245 SED, 252-289 SEC7, 297-345 SEC8, 353-411 SECDED; CRC32C is
// this repository's own mode).
//
// B200 design (DESIGN.md "spmv kernel"):
//  * Work unit = "tile": a run of whole rows with <= kTileNnz non-zeros and
//    <= BLOCK rows (one per thread of the CTA), built once in create_matrix.
//  * Persistent CTAs (grid = SMs x resident CTAs), tile t of CTA b is b + t*grid so
//    the whole chip sweeps the matrix front together and vec[] stays L2-resident.
//  * The tile's cols / values / row pointers are contiguous in memory: one elected
//    thread streams them into shared memory with TMA bulk copies (cp.async.bulk +
//    mbarrier complete_tx), kStages deep, so HBM latency is hidden behind the ECC
//    arithmetic of the previous tile instead of behind occupancy.
//  * Regular tiles (row lengths within 2x of the mean; every stencil matrix):
//    thread r walks row r in shared memory -- check, gather vec[col] (lanes = rows,
//    so a banded matrix gathers coalesced), multiply, add -- in ELEMENT ORDER with
//    separate multiply and add, i.e. exactly the reference's
//    `tmp += values[i] * vec[col]`: result[] is bit-identical to the CPU target.
//  * Irregular tiles (power-law rows): element-parallel check + product into shared
//    memory, one CTA barrier, then a warp-local row phase: a lane adds its row's products
//    in element order (same bits as the CPU sum) for rows of up to 32 elements, the warp
//    takes its lanes' longer rows one at a time (lane-strided sums + shuffle tree).
//  * A row longer than a tile is a tile of its own, streamed from global memory in
//    chunks with a tree reduction (not bit-identical to the sequential sum; < 1e-10).
//  * Matrices with irregular tiles run the DYN variant: tiles are drawn from a device
//    counter (their cost varies 5x), dot() partials are kept per tile so the result does
//    not depend on the schedule, and tiles are smaller so that the L1 keeps 92 KB for the
//    gathers in flight (profiles/r1_powerlaw_investigation.md).
//  * The owning lane corrects a faulty element in registers, writes it back to
//    global memory and logs the event; dot(vec, result) is reduced in the epilogue.
// CONSTRAINTS element events carry their global row in `bit` (the host orders them
// against ROW_* events with it and clears it before reporting).
#pragma once
#include "common.cuh"
#include "ecc.cuh"
#include "rowsecc.cuh"

namespace abft {

// Threads per CTA = rows per tile (one row per thread on regular tiles); a tile holds up to 8 elements per row.
// Measured on 256^3 secded (profiles/r2_ab_session2_block_sizes.jsonl, r2_ab_session3_warp_vs_cta.jsonl): 256 threads /
// 2048-element tiles 0.33 ms; 128 threads 0.35; 64 threads 0.39; 512 threads 0.40; a private two-stage TMA ring per
// warp (32-row tiles, no CTA barrier at all) 0.40 -- the per-tile bookkeeping (~100 instructions per warp per tile,
// a quarter of everything executed) and the stream set-up grow faster than the barrier wait shrinks.
constexpr int kSpmvBlock = kBlock;
constexpr int kTileNnz = 8 * kSpmvBlock;
// Matrices with irregular tiles (dynamic schedule) are gather-bound, and a gather in flight holds an L1 line: with
// 4 CTAs x 52 KB of stage buffers the SM's 256 KB are all shared memory and the ~28 KB of L1 left cap the gathers in
// flight (measured: same kernel, carveout forced down to 164 KB / 3 CTAs: 0.41 -> 0.32 ms).  Their tiles are smaller,
// so that 4 CTAs need 160 KB and the L1 keeps 92 KB.  Measured on the power-law workload (sec7, 4 CTAs per SM):
// 2048 -> 0.40 ms (L1 28 KB), 1536 -> 0.295 (92 KB), 1024 -> 0.335 (124 KB), 768 -> 0.378 (156 KB).
constexpr int kSpmvBlockDyn = kBlock;
constexpr int kTileNnzDyn = 1536;
constexpr int kStages = 2;
constexpr uint32_t kSeqRowMax = 32;  // irregular tiles: rows up to this long are summed sequentially by one thread
constexpr int kLongBatch = 4;  // rows longer than a tile: elements each thread streams at a time
// irregular tiles: gathers each thread keeps in flight in the element phase (measured on the power-law workload:
// 1 -> 0.50 ms, 2 -> 0.52, 4 -> 0.53, 8 -> 0.60: the gather rate is the bound, not the loads in flight)
constexpr int kElemBatch = 1;
constexpr uint32_t kOverrun = 8;  // elements staged past a tile's end (>= longest unrolled walk - 1)

constexpr uint32_t kTileNone = 0xFFFFFFFFu;   // info of the "no more tiles" marker
constexpr uint32_t kTileRegular = 1u << 16;  // info flag: thread-per-row path
constexpr uint32_t kTileLongRow = 1u << 17;  // info flag: single row > kTileNnz, streamed
constexpr int kTileMaxLenShift = 18;         // info bits 18..29: longest row of the tile

struct Tile {
  uint32_t row_begin;  // local row
  uint32_t nnz_begin;  // local element indices
  uint32_t nnz_end;
  uint32_t info;       // bits 0..15 number of rows, 16..17 flags, 18..29 longest row
};

struct MatView {
  const uint32_t *rowptr;  // [n_rows+1 (+pad)] local element indices
  uint32_t *cols;          // [nnz (+pad)] column word incl. check bits
  double *values;          // [nnz (+pad)]
  const Tile *tiles;
  unsigned int *tile_counter;  // dynamic schedule (matrices with irregular tiles): next tile to hand out; null: tile b, b+grid, ...
  uint32_t n_tiles;
  uint32_t n_rows, nnz;
  uint32_t n_cols;      // global N: bound for the gather
  uint32_t row_offset;  // global index of local row 0
  uint32_t nnz_offset;  // global index of local element 0 (for reported indices)
  uint32_t rows_ecc;    // rowptr[] carries the SECDED pair code of rowsecc.cuh (check bits in bits 28..31)
};

// one pipeline stage of staged tile data (windows start at 16-byte aligned indices)
// (+16: the windows are rounded out to 16 bytes, and the unrolled row walk may read up to 6 slots
// past a row's end)
template <int TN, int ROWS>
struct __align__(16) StageBufT {
  double vals[TN + 16];
  uint32_t cols[TN + 16];
  uint32_t rptr[ROWS + 8];
};
template <bool DYN>
struct SpmvCfg {
  static constexpr int block = DYN ? kSpmvBlockDyn : kSpmvBlock;   // threads per CTA = rows per tile
  static constexpr int tile_nnz = DYN ? kTileNnzDyn : kTileNnz;
  using Stage = StageBufT<tile_nnz, block>;
  static constexpr size_t dyn_smem = sizeof(Stage) * kStages;
};

// ---- PTX: mbarrier + TMA bulk copy (global -> shared) -------------------------------
#ifndef ABFT_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// The matrix stream is read exactly once per spmv: tag it evict-first in L2 so that it does not push
// out vec[], which every row block re-reads (7-point stencil: each entry is gathered by 7 rows spread
// over ~11 MB of stream).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fence_mbarrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// sum += prod if ok: a predicated add (not a select between sum and sum + prod: that costs two FSEL per element)
__device__ __forceinline__ double add_if(double sum, double prod, bool ok) {
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p add.rn.f64 %0, %0, %1;\n}" : "+d"(sum) : "d"(prod), "r"((uint32_t)ok));
  return sum;
}
#else
// Host emulation (tests/helpers/cuda_shim_mt/, test infrastructure): the same contracts in plain C++, so that the
// kernel below can run as 256 CPU threads against the oracle.  Never compiled by nvcc.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { emu::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { emu::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
__device__ __forceinline__ uint64_t l2_evict_first_policy() { return 0; }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t) {
  emu::bulk_g2s(dst, src, bytes, bar);
}
__device__ __forceinline__ void fence_mbarrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ double add_if(double sum, double prod, bool ok) { return ok ? __dadd_rn(sum, prod) : sum; }
#endif

__device__ __forceinline__ void split_value(double v, uint32_t &w0, uint32_t &w1) {
  w0 = (uint32_t)__double2loint(v);
  w1 = (uint32_t)__double2hiint(v);
}

// ---- the per-element check ------------------------------------------------------------
// Slow path (a fault was seen): everything by value / in registers, nothing on the stack.
// Returns {w0, w1, col, changed}.  SEC7 / SEC8 / SECDED undo the flip the syndrome points
// at (CSR/CPUContext.cpp:272-278, 320-334, 376-390); a multi-bit fault can yield bit >= 96,
// where the reference writes past its 12-byte element (UB): the bit is reported and the
// element left alone.
template <int MODE>
__device__ __noinline__ uint4 handle_fault(uint32_t w0, uint32_t w1, uint32_t col, uint32_t gindex, DevLog *log,
                                           volatile unsigned int *flag) {
  const LogRef L{log, flag};
  uint32_t changed = 0;
  if ((MODE & 7) == ABFT_MODE_SED) {
    log_event(L, ABFT_EV_SED_DETECTED, gindex, 0);
  } else {
    const uint32_t odd = overall_parity(w0, w1, col);
    const uint32_t s = syndrome7(w0, w1, col);
    const bool correct = ((MODE & 7) == ABFT_MODE_SEC7) ? (s != 0) : (odd != 0);
    if (correct) {
      if (s) {
        const uint32_t bit = flipped_bit(s);
        if (bit < 32) w0 ^= 1u << bit;
        else if (bit < 64) w1 ^= 1u << (bit - 32);
        else if (bit < 96) col ^= 1u << (bit - 64);
        log_event(L, ABFT_EV_CORRECTED_BIT, gindex, bit);
      } else {
        col ^= 1u << 24;
        log_event(L, ABFT_EV_CORRECTED_PARITY, gindex, 0);
      }
      changed = 1;
    } else if ((MODE & 7) == ABFT_MODE_SECDED && s) {
      log_event(L, ABFT_EV_DOUBLE_BIT, gindex, 0);
    }
  }
  return make_uint4(w0, w1, col, changed);
}

// bit 0 set iff the element fails the mode's fast check
template <int MODE>
__device__ __forceinline__ uint32_t element_suspect(uint32_t w0, uint32_t w1, uint32_t col) {
  if ((MODE & 7) == ABFT_MODE_SED || (MODE & 7) == ABFT_MODE_SEC8) return overall_parity(w0, w1, col);
  if ((MODE & 7) == ABFT_MODE_SEC7) return syndrome7_any(w0, w1, col);
  if ((MODE & 7) == ABFT_MODE_SECDED) return (overall_parity(w0, w1, col) | syndrome7_any(w0, w1, col)) & 1u;
  return 0;
}

// acc | (verdict bits of one element), without the final `& 1`: callers OR these over a whole row and test
// bit 0 once.  Written as a flat chain of 3-input ORs (one LOP3 each).
template <int MODE>
__device__ __forceinline__ uint32_t or_suspect_bits(uint32_t acc, uint32_t w0, uint32_t w1, uint32_t col) {
  if ((MODE & 7) == ABFT_MODE_SED || (MODE & 7) == ABFT_MODE_SEC8) return acc | __popc(w0 ^ w1 ^ col);
  if ((MODE & 7) == ABFT_MODE_SEC7 || (MODE & 7) == ABFT_MODE_SECDED) {
    acc = acc | __popc(masked<1>(w0, w1, col)) | __popc(masked<2>(w0, w1, col));
    acc = acc | __popc(masked<3>(w0, w1, col)) | __popc(masked<4>(w0, w1, col));
    acc = acc | __popc(masked<5>(w0, w1, col)) | __popc(masked<6>(w0, w1, col));
    if ((MODE & 7) == ABFT_MODE_SECDED) acc = acc | __popc(masked<7>(w0, w1, col)) | __popc(w0 ^ w1 ^ col);
    else acc = acc | __popc(masked<7>(w0, w1, col));
  }
  return acc;
}

// MODE, throughout this file: the ABFT mode in bits 0..2; bit 3 (kVecSed) adds the vector protection of mode secdedv
// (ecc.cuh: parity in every vector element's lowest mantissa bit) to what the kernel does for the base mode.
constexpr int kVecSed = 8;
template <int MODE>
struct ModeTraits {
  static constexpr bool per_element_ecc = (MODE & 7) >= ABFT_MODE_SED && (MODE & 7) <= ABFT_MODE_SECDED;
  static constexpr bool masks_column = (MODE & 7) >= ABFT_MODE_SED;  // check bits in column bits 24..31
  static constexpr bool vec_sed = (MODE & kVecSed) != 0;
};

// check (and repair) one element, gather vec[col], return the rounded product
template <int MODE>
__device__ __forceinline__ double element_product(uint32_t cw, double v, uint32_t li, const MatView &A,
                                                  const double *__restrict__ x, const LogRef &L) {
  if (ModeTraits<MODE>::per_element_ecc) {
    uint32_t w0, w1;
    split_value(v, w0, w1);
    if (element_suspect<MODE>(w0, w1, cw)) {
      const uint4 f = handle_fault<MODE>(w0, w1, cw, A.nnz_offset + li, L.log, L.flag);
      if (f.w) {
        cw = f.z;
        v = __hiloint2double((int)f.y, (int)f.x);
        A.cols[li] = cw;
        A.values[li] = v;
      }
    }
  }
  const uint32_t col = ModeTraits<MODE>::masks_column ? (cw & 0x00FFFFFFu) : cw;
  // a column no check caught can point outside vec[] (the reference reads out of bounds: UB): clamp it
  const double xv = __ldg(x + min(col, A.n_cols - 1u));
  if (ModeTraits<MODE>::vec_sed && (vec_parity_bits(xv) & 1u)) log_event(L, ABFT_EV_VECTOR_SED, min(col, A.n_cols - 1u), VEC_SPMV_OPERAND);
  return __dmul_rn(v, xv);
}

// U elements at once (element-parallel paths): all checks and gathers first (U independent loads in flight),
// then the products.  The checks only OR their verdict into one word; a batch that trips goes through
// element_product one element at a time (rare).  Slots with ok[k] == false must still hold a valid element
// (callers repeat their first one): they are checked and gathered, never repaired, and return garbage.
template <int MODE, int U>
__device__ __forceinline__ void batch_products(const uint32_t (&li)[U], const uint32_t (&cw)[U], const double (&v)[U],
                                               const bool (&ok)[U], const MatView &A, const double *__restrict__ x,
                                               const LogRef &L, double (&out)[U]) {
  double xv[U];
  uint32_t suspect = 0;
#pragma unroll
  for (int k = 0; k < U; k++) {
    if (ModeTraits<MODE>::per_element_ecc) {
      uint32_t w0, w1;
      split_value(v[k], w0, w1);
      suspect = or_suspect_bits<MODE>(suspect, w0, w1, cw[k]);
    }
    uint32_t col = ModeTraits<MODE>::masks_column ? (cw[k] & 0x00FFFFFFu) : cw[k];
    xv[k] = __ldg(x + min(col, A.n_cols - 1u));
    if (ModeTraits<MODE>::vec_sed) suspect |= vec_parity_bits(xv[k]);
  }
  if (!(suspect & 1u)) {
#pragma unroll
    for (int k = 0; k < U; k++) out[k] = __dmul_rn(v[k], xv[k]);
  } else {
    for (int k = 0; k < U; k++)
      if (ok[k]) out[k] = element_product<MODE>(cw[k], v[k], li[k], A, x, L);
  }
}

// Fast walk over up to LEN consecutive elements of one row, fully unrolled: all shared-memory
// loads first, then all checks and gathers (LEN independent loads in flight), then the products
// are added to `sum` in element order.  No per-element branch: the checks only OR their verdict
// into `suspect`; the caller re-walks the row on the careful path if bit 0 ends up set (rare).
// Slots past the row's end (k >= count) are still loaded and checked -- they hold the next row's
// elements (the stage is filled kOverrun elements past the tile's end for exactly this; the buffers start
// zeroed and the matrix arrays end in zero padding), i.e. valid code words; at worst a faulty neighbour
// raises a false `suspect` and sends this row down the (exact) careful path -- but they are never added.
// That keeps the loop body free of selects and zero-fills.
// `full`: a warp-uniform vote of the caller that every row of the warp fills the chunk (interior rows of a stencil:
// 3 of 4 warps of a 256^3 tile); the adds then need no predicate (the predicated form costs an ISETP and two FSEL per
// element: 280 instead of 320 instructions per 7-element SECDED row, measured 0.332 -> 0.325 ms on 256^3).
template <int MODE, int LEN>
__device__ __forceinline__ double walk_chunk(double sum, uint32_t first, uint32_t count, const uint32_t *s_cols,
                                             const double *s_vals, uint32_t n_cols, const double *__restrict__ x,
                                             uint32_t &suspect, bool full = false) {
  uint32_t c[LEN + 1];
  double v[LEN], xv[LEN];
  bool ok[LEN];
#pragma unroll
  for (int k = 0; k < LEN; k++) {
    c[k] = s_cols[first + k];
    v[k] = s_vals[first + k];
  }
  if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) c[LEN] = s_cols[first + LEN];
  // every gather is issued before the first check (ncu had 7 % of the stall samples on the first DMUL, waiting for the
  // gather issued last; ptxas places the loads mid-way through the checks either way: measured neutral, 0.328 ms)
#pragma unroll
  for (int k = 0; k < LEN; k++) {
    const uint32_t col = ModeTraits<MODE>::masks_column ? (c[k] & 0x00FFFFFFu) : c[k];
    // A column no check caught (mode none, or a multi-bit fault the code cannot see) can point outside
    // vec[]: the reference reads out of bounds there (UB); here, and in the oracle, the column is clamped
    // to the last entry.
    xv[k] = __ldg(x + min(col, n_cols - 1u));
    ok[k] = ((uint32_t)k < count);
  }
#pragma unroll
  for (int k = 0; k < LEN; k++) {
    if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) {
      // CSR/CPUContext.cpp:188-200: col < N, and strictly increasing inside the row
      suspect |= (c[k] >= n_cols) ? 1u : 0u;
      suspect |= ((uint32_t)(k + 1) < count && c[k + 1] <= c[k]) ? 1u : 0u;
    }
    if (ModeTraits<MODE>::per_element_ecc) {
      uint32_t w0, w1;
      split_value(v[k], w0, w1);
      suspect = or_suspect_bits<MODE>(suspect, w0, w1, c[k]);
    }
    if (ModeTraits<MODE>::vec_sed) suspect |= vec_parity_bits(xv[k]);  // (the one-by-one path names the element)
  }
  if (full) {
#pragma unroll
    for (int k = 0; k < LEN; k++) sum = __dadd_rn(sum, __dmul_rn(v[k], xv[k]));
    return sum;
  }
#pragma unroll
  for (int k = 0; k < LEN; k++) sum = add_if(sum, __dmul_rn(v[k], xv[k]), ok[k]);
  return sum;
}

// CONSTRAINTS, warp-collective: true iff no element of the row violates `col < N` or the strict
// column order (CSR/CPUContext.cpp:188-200).  cols points at the row's first staged column word.
__device__ __forceinline__ bool constraints_fast_ok_row(const uint32_t *cols, uint32_t len, uint32_t n_cols, int lane) {
  bool bad = false;
  for (uint32_t i = lane; i < len; i += 32) {
    const uint32_t c = cols[i];
    bad |= (c >= n_cols);
    if (i + 1 < len) bad |= (cols[i + 1] <= c);
  }
  return !__any_sync(0xffffffffu, bad);
}

// ---- CRC32C row check ---------------------------------------------------------------------------------------------
// create_matrix (encode.cuh) walks a row with the classic byte tables; the spmv check is organised differently,
// because 12 byte-table lookups per element, hit by 32 lanes at random, are ~3-way bank conflicted and that alone
// (41 shared-memory wavefronts per 32 elements) held round 1's kernel at 0.53 of the HBM roofline:
//  * The raw state after one 32-bit word is F(state ^ word), F = "advance over four zero bytes", a LINEAR map.  Over
//    one element {w0, w1, col}:  state' = F3(state ^ w0) ^ F2(w1) ^ F(col)  (F2 = F.F, F3 = F.F.F).  Only the first
//    term is on the row's serial chain: 7 dependent steps per 7-point row instead of 21.
//  * A linear map of a 32-bit word is the XOR of table entries of its 5-BIT chunks (7 chunks; col has 24 bits: 5).
//    A 32-entry table laid out one entry per bank is conflict-free whatever the lanes ask for (same entry = same
//    address = broadcast).  19 conflict-free lookups per element, 21 tables x 128 B = 2.7 KB per CTA.
__device__ uint32_t g_crc5_tab[kCrcTabWords];  // [(m-1)*7 + j][c], filled by the host at library init
__device__ uint32_t g_crc_tab[1024];           // slicing-by-4 byte tables (create_matrix), filled by the host at library init

// CRC-32C of one 32-bit word appended to a running (raw) state; slicing-by-4 tables tab[s*256 + b]
__device__ __forceinline__ uint32_t crc_word(uint32_t crc, uint32_t w, const uint32_t *tab) {
  const uint32_t x = crc ^ w;
  return tab[768 + (x & 0xFF)] ^ tab[512 + ((x >> 8) & 0xFF)] ^ tab[256 + ((x >> 16) & 0xFF)] ^ tab[x >> 24];
}

// ---- the same check with the tables in REGISTERS (stencil instantiations) ------------------------------------------
// Shared-memory lookups, conflict-free or not, cost the data pipe one wavefront each and the ALU pipe a shift and a
// mask to form the address; that pipe pair is what bounds the kernel above.  A 32-entry table is also one register
// per lane -- lane c holds entry c -- and a lookup one SHFL.IDX whose lane operand is `x >> 5j` as it stands (the
// instruction reads the low five bits only): no mask, no address, and for chunk 0 no shift.  The chunks are cut from
// the whole 88-bit element {state ^ w0, w1, col} (ecc.cuh, crc_elem_entry): 18 lookups instead of 7 + 7 + 5, two
// of them through a funnel shift across a word boundary.  The whole warp must walk in step.  256^3: 0.459 -> 0.405 ms
// (profiles/r2_ab_session11_crc_register_tables.jsonl).  SHFL shares the data pipe with LDS (tools/microbench.cu:
// together they issue at the rate of either), so the kernel stays bound by that pipe (82 % busy) next to the ALU pipe
// (62 %); the last, three-bit chunk computed without a lookup (multiply, three PRMT sign masks, three LOP3) was slower.
__device__ uint32_t g_crc_elem_tab[kCrcElemChunks * 32];  // [chunk][c], filled by the host at library init
struct CrcRegs {
  uint32_t t[kCrcElemChunks];
};
__device__ __forceinline__ void crc_regs_load(CrcRegs &cr, int lane) {
#pragma unroll
  for (int i = 0; i < kCrcElemChunks; i++) cr.t[i] = g_crc_elem_tab[i * 32 + lane];
}
#define ABFT_CRC_LOOKUP(i, idx) __shfl_sync(0xffffffffu, cr.t[i], (int)(idx))
// the 11 lookups of an element that do not depend on the running state (bits 35..87: w1 from bit 3 on, and col)
__device__ __forceinline__ uint32_t crc_elem_free(uint32_t w1, uint32_t colw, uint32_t col, const CrcRegs &cr) {
  uint32_t r = ABFT_CRC_LOOKUP(7, w1 >> 3) ^ ABFT_CRC_LOOKUP(8, w1 >> 8) ^ ABFT_CRC_LOOKUP(9, w1 >> 13);
  r ^= ABFT_CRC_LOOKUP(10, w1 >> 18) ^ ABFT_CRC_LOOKUP(11, w1 >> 23) ^ ABFT_CRC_LOOKUP(12, __funnelshift_r(w1, colw, 28));
  // (colw = the stored column word, check byte and all: its bits above 23 never reach the low five of these four)
  r ^= ABFT_CRC_LOOKUP(13, colw >> 1) ^ ABFT_CRC_LOOKUP(14, colw >> 6) ^ ABFT_CRC_LOOKUP(15, colw >> 11);
  return r ^ ABFT_CRC_LOOKUP(16, colw >> 16) ^ ABFT_CRC_LOOKUP(17, col >> 21);
}
// ... and the 7 that do (bits 0..34: x = state ^ w0, and the low three bits of w1)
__device__ __forceinline__ uint32_t crc_elem_chain(uint32_t x, uint32_t w1, const CrcRegs &cr) {
  const uint32_t a = ABFT_CRC_LOOKUP(0, x) ^ ABFT_CRC_LOOKUP(1, x >> 5) ^ ABFT_CRC_LOOKUP(2, x >> 10);
  const uint32_t b = ABFT_CRC_LOOKUP(3, x >> 15) ^ ABFT_CRC_LOOKUP(4, x >> 20) ^ ABFT_CRC_LOOKUP(5, x >> 25);
  return a ^ b ^ ABFT_CRC_LOOKUP(6, __funnelshift_r(x, w1, 30));
}

// F^M(x) from the 5-bit chunk tables in shared memory (tab = all 21 tables); NCH chunks of x can be non-zero
template <int M, int NCH>
__device__ __forceinline__ uint32_t crc_map(uint32_t x, const uint32_t *tab) {
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) {
    // byte offset of entry chunk_j(x) inside its table: the chunk, times 4
    // (tried: the shift as a multiply-high by a power of two ptxas cannot see, i.e. IMAD.HI on the idle FMA pipe instead
    // of SHF on the ALU pipe that bounds this kernel -- ptxas folds the table base into the 64-bit addend and pays for
    // it with two more moves per lookup: 0.57 ms against 0.47, profiles/r2_ab_session2_block_sizes.jsonl)
    const uint32_t off = (j == 0) ? ((x << 2) & 0x7Cu) : ((x >> (5 * j - 2)) & 0x7Cu);
    r ^= *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(tab + ((M - 1) * kCrcChunks + j) * 32) + off);
  }
  return r;
}

// the raw state after one more element (col without its top byte)
__device__ __forceinline__ uint32_t crc_element(uint32_t state, uint32_t w0, uint32_t w1, uint32_t col, const uint32_t *tab) {
  return crc_map<3, 7>(state ^ w0, tab) ^ crc_map<2, 7>(w1, tab) ^ crc_map<1, 5>(col, tab);
}

// Unrolled CRC32C row walk for regular tiles, walk_chunk's shape: all shared-memory loads, then per element the
// gather and the state-independent two thirds of the CRC step, with the serial third (F3) chained through; the
// products are added in element order.  For rows of 4..LEN elements (the CRC sits in the top bytes of the first
// four column words); anything else takes the one-by-one loop.  WARP: every lane of the warp is here (lanes without
// a row included), tables in registers (cr); else tables in shared memory (crc_tab).
template <int LEN, bool WARP = false>
__device__ __forceinline__ double walk_crc(uint32_t count, const uint32_t *s_cols, const double *s_vals, uint32_t n_cols,
                                           const double *__restrict__ x, const uint32_t *crc_tab, bool full, bool &bad,
                                           const CrcRegs *cr = nullptr) {
  uint32_t c[LEN];
  double v[LEN], xv[LEN];
#pragma unroll
  for (int k = 0; k < LEN; k++) {
    c[k] = s_cols[k];
    v[k] = s_vals[k];
  }
  uint32_t crc = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < LEN; k++) {
    const uint32_t col = c[k] & 0x00FFFFFFu;
    xv[k] = __ldg(x + min(col, n_cols - 1u));
    uint32_t w0, w1;
    split_value(v[k], w0, w1);
    const uint32_t next = WARP ? (crc_elem_chain(crc ^ w0, w1, *cr) ^ crc_elem_free(w1, c[k], col, *cr))
                               : crc_element(crc, w0, w1, col, crc_tab);
    crc = (full || (uint32_t)k < count) ? next : crc;        // slots past the row's end hold the next row's elements
  }
  // (count >= 4: the first four column words are the row's own)
  const uint32_t stored = __byte_perm(__byte_perm(c[0], c[1], 0x7373), __byte_perm(c[2], c[3], 0x7373), 0x5410);
  bad = (crc ^ 0xFFFFFFFFu) != stored;
  double sum = 0.0;
  if (full) {
#pragma unroll
    for (int k = 0; k < LEN; k++) sum = __dadd_rn(sum, __dmul_rn(v[k], xv[k]));
    return sum;
  }
#pragma unroll
  for (int k = 0; k < LEN; k++) sum = add_if(sum, __dmul_rn(v[k], xv[k]), (uint32_t)k < count);
  return sum;
}

// multi-GPU: the peers whose pushes into vec[] must have landed before it is read
struct WaitSpec {
  const CommWindow *me;  // local window
  unsigned int epoch;    // wait until pflag[s] >= epoch ...
  unsigned int mask;     // ... for every rank s in this mask (0: single GPU, no wait)
  unsigned long long *phase;  // this iteration's row of phase stamps (null: off)
};

// Protected row pointers (rowsecc.cuh): check every pair of a staged row-pointer window before anyone looks at it,
// repair single flips (in the stage and in global memory), strip the check bits; after the next CTA barrier the stage
// holds plain pointers and the rest of the kernel is unchanged.  The window starts and ends on multiples of four
// pointers, so pairs are never split; words past rows[n_rows] are zero padding (a valid code word).
template <int BLOCK>
__device__ __forceinline__ void rows_prepass(uint32_t *rp, const Tile &t, const MatView &A, const LogRef &L, int tid) {
  const uint32_t wb = t.row_begin & ~3u;
  const uint32_t n_words = ((t.row_begin + (t.info & 0xFFFFu) + 1 + 3) & ~3u) - wb;
  // the UPPER half of the CTA's threads takes the pairs: thread 0's warp already runs the tile refill on top of its
  // rows and is the one the per-tile barrier waits for
  const int u = BLOCK >= 64 ? tid - BLOCK / 2 : tid;
  if (u < 0) return;
  for (uint32_t k = 2 * (uint32_t)u; k + 1 < n_words; k += (BLOCK >= 64 ? BLOCK : 2 * BLOCK)) {
    uint32_t w0 = rp[k], w1 = rp[k + 1], bit = 0;
    const int verdict = rows_check_pair(w0, w1, &bit);
    const uint32_t pair = (A.row_offset + wb + k) / 2;   // (partitions start on even rows when protected)
    if (verdict == 1) {
      log_event(L, ABFT_EV_ROWPTR_CORRECTED, pair, bit);
      uint32_t *g = const_cast<uint32_t *>(A.rowptr) + wb + k;
      g[0] = w0;
      g[1] = w1;
    } else if (verdict == 2) {
      log_event(L, ABFT_EV_ROWPTR_DOUBLE, pair, 0);
    }
    rp[k] = w0 & kRowsDataMask;
    rp[k + 1] = w1 & kRowsDataMask;
  }
}

// CONSTRAINTS, one row, exactly the reference's checks in the reference's order (CSR/CPUContext.cpp:173-182 row checks,
// :184-203 element checks); returns the row sum.  s_cols / s_vals: the tile's staged windows (starting at element
// indices cb / vb); staged_vals_ok = false: the staged values have been overwritten (irregular tiles: by products).
__device__ __forceinline__ double constraints_careful_row(uint32_t lrow, uint32_t rs, uint32_t re, const Tile &t,
                                                          const uint32_t *s_cols, const double *s_vals, uint32_t cb,
                                                          uint32_t vb, bool staged_vals_ok, const MatView &A,
                                                          const double *__restrict__ x, const LogRef &L) {
  double sum = 0.0;
  if (re > A.nnz) log_event(L, ABFT_EV_ROW_SIZE, A.row_offset + lrow, 0);
  else if (re < rs) log_event(L, ABFT_EV_ROW_ORDER, A.row_offset + lrow, 0);
  else
    for (uint32_t i = rs; i < re; i++) {
      // elements normally sit in the staged window; a corrupted row pointer can send the row
      // outside it, then go to global memory (i < nnz was checked above)
      const bool in = (i >= t.nnz_begin) && (i < t.nnz_end);
      const uint32_t col = in ? s_cols[i - cb] : A.cols[i];
      if (col >= A.n_cols) { log_event(L, ABFT_EV_COL_SIZE, A.nnz_offset + i, A.row_offset + lrow); break; }
      if (i < re - 1) {
        const bool in2 = (i + 1 >= t.nnz_begin) && (i + 1 < t.nnz_end);
        const uint32_t nxt = in2 ? s_cols[i + 1 - cb] : A.cols[i + 1];
        if (nxt <= col) { log_event(L, ABFT_EV_COL_ORDER, A.nnz_offset + i, A.row_offset + lrow); break; }
      }
      const double val = (in && staged_vals_ok) ? s_vals[i - vb] : A.values[i];
      sum = __dadd_rn(sum, __dmul_rn(val, __ldg(x + col)));
    }
  return sum;
}

// One row of a regular tile per lane, straight from the staged stream (WARP-COLLECTIVE: every lane of the warp calls it,
// lanes without a row with has_row = false).  `k` = the lane's row number inside the tile.  Writes y[row] and adds the
// row's share of dot(vec, result) to dot_part.
// FLEN != 0: the matrix-wide longest row, fixed at compile time (the stencil instantiations of spmv_kernel): one unrolled
// walk, no dispatch on the tile's own longest row.
template <int MODE, int FLEN = 0>
__device__ __forceinline__ void regular_row(const Tile &t, uint32_t k, bool has_row, const uint32_t *s_rptr,
                                            const uint32_t *s_cols, const double *s_vals, const MatView &A,
                                            const double *__restrict__ x, double *__restrict__ y, const LogRef &L,
                                            const uint32_t *s_crc, double &dot_part, const CrcRegs &cr) {
  const uint32_t rb = t.row_begin & ~3u, cb = t.nnz_begin & ~3u, vb = t.nnz_begin & ~1u;
  const uint32_t row = t.row_begin + k;
  const uint32_t max_len = FLEN ? (uint32_t)FLEN : ((t.info >> kTileMaxLenShift) & 0xFFFu);  // uniform over the tile
  const uint32_t need = max_len <= 4 ? 4u : max_len;               // the unrolled walk's length
  uint32_t rs = 0, re = 0;
  double xrow = 0.0;
  if (has_row) {
    rs = s_rptr[row - rb];
    re = s_rptr[row + 1 - rb];
    xrow = __ldg(x + A.row_offset + row);   // dot()'s operand: in flight behind the walk, not after it
    if ((MODE & 7) != ABFT_MODE_CONSTRAINTS) {
      // these modes trust rows[] like the reference; clamp so a bad pointer cannot fault
      rs = min(max(rs, t.nnz_begin), t.nnz_end);
      re = min(max(re, rs), t.nnz_end);
    }
  }
  const uint32_t len = re - rs;
  // (CONSTRAINTS: only if its row pointers keep the row inside the staged window; anything else,
  // i.e. a corrupted rows[], takes the careful path that mirrors the reference's checks one by one)
  const bool in_window = (MODE & 7) != ABFT_MODE_CONSTRAINTS || (rs >= t.nnz_begin && re <= t.nnz_end && rs <= re);
  // does every row of this warp fill the unrolled walk?  Lanes without a row do not veto.
  const bool full_rows = max_len <= 7 && __all_sync(0xffffffffu, !has_row || (in_window && len == need));
  double sum = 0.0;
  bool careful = true;
  if ((MODE & 7) == ABFT_MODE_CRC32C && FLEN >= 4 && FLEN <= 7) {
    // register tables: the whole warp walks in step (SHFL), so only if every one of its rows takes the unrolled
    // walk; lanes without a row walk the tile's first slots and drop the result
    if (__all_sync(0xffffffffu, !has_row || (len >= 4 && len <= need))) {
      const uint32_t *sc0 = has_row ? s_cols + (rs - cb) : s_cols;
      const double *sv0 = has_row ? s_vals + (rs - vb) : s_vals;
      bool bad = false;
      sum = walk_crc<(FLEN >= 4 && FLEN <= 7) ? FLEN : 4, true>(len, sc0, sv0, A.n_cols, x, s_crc, full_rows, bad, &cr);
      if (!has_row) return;
      if (bad) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + row, 0);
      y[row] = sum;
      dot_part = __dadd_rn(dot_part, __dmul_rn(sum, xrow));
      return;
    }
  }
  if (!has_row) return;
  const uint32_t *sc = s_cols + (rs - cb);
  const double *sv = s_vals + (rs - vb);
  if ((MODE & 7) != ABFT_MODE_CRC32C && in_window) {
    uint32_t suspect = 0;
    // max_len is what create_matrix saw.  A corrupted rows[] can make this row longer while keeping it inside
    // the window; the reference walks the whole [start, end) range (CSR/CPUContext.cpp:125,184), so such a
    // row goes down the one-by-one path instead of an unrolled walk that would stop after max_len elements.
    if (max_len <= 7 && len > need) suspect = 1u;
    else if (FLEN) sum = walk_chunk<MODE, (FLEN >= 4 && FLEN <= 7) ? FLEN : 4>(sum, 0, len, sc, sv, A.n_cols, x, suspect, full_rows);
    else if (max_len <= 4) sum = walk_chunk<MODE, 4>(sum, 0, len, sc, sv, A.n_cols, x, suspect, full_rows);
    else if (max_len == 5) sum = walk_chunk<MODE, 5>(sum, 0, len, sc, sv, A.n_cols, x, suspect, full_rows);
    else if (max_len == 6) sum = walk_chunk<MODE, 6>(sum, 0, len, sc, sv, A.n_cols, x, suspect, full_rows);
    else if (max_len == 7) sum = walk_chunk<MODE, 7>(sum, 0, len, sc, sv, A.n_cols, x, suspect, full_rows);
    else
      // longer rows: chunks of 4 up to this row's own length (never more than 3 slots past its end,
      // which the over-staged window covers; the trip count may differ between lanes)
      for (uint32_t base = 0; base < len; base += 4)
        sum = walk_chunk<MODE, 4>(sum, base, len - base, sc, sv, A.n_cols, x, suspect);
    careful = (suspect & 1u) != 0;  // a check tripped somewhere in this row: walk it again, carefully
  }
  if ((MODE & 7) == ABFT_MODE_CRC32C && max_len <= 7 && len >= 4 && len <= need) {
    bool bad = false;
    if (FLEN) sum = walk_crc<(FLEN >= 4 && FLEN <= 7) ? FLEN : 4>(len, sc, sv, A.n_cols, x, s_crc, full_rows, bad);
    else if (max_len <= 4) sum = walk_crc<4>(len, sc, sv, A.n_cols, x, s_crc, full_rows, bad);
    else if (max_len == 5) sum = walk_crc<5>(len, sc, sv, A.n_cols, x, s_crc, full_rows, bad);
    else if (max_len == 6) sum = walk_crc<6>(len, sc, sv, A.n_cols, x, s_crc, full_rows, bad);
    else sum = walk_crc<7>(len, sc, sv, A.n_cols, x, s_crc, full_rows, bad);
    if (bad) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + row, 0);
    careful = false;
  }
  if (careful) {
    sum = 0.0;
    if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) {
      sum = constraints_careful_row(row, rs, re, t, s_cols, s_vals, cb, vb, true, A, x, L);
    } else if ((MODE & 7) == ABFT_MODE_CRC32C) {
      uint32_t crc = 0xFFFFFFFFu, stored = 0;
      bool bad = false;
      for (uint32_t i = rs; i < re; i++) {
        const uint32_t cw = s_cols[i - cb];
        const double val = s_vals[i - vb];
        uint32_t w0, w1;
        split_value(val, w0, w1);
        if (len >= 4) {
          if (i - rs < 4) stored |= (cw >> 24) << (8 * (i - rs));
          crc = crc_element(crc, w0, w1, cw & 0x00FFFFFFu, s_crc);
        } else {
          // too short to hold 32 CRC bits: SECDED check byte per element, detect-only
          if ((overall_parity(w0, w1, cw) | syndrome7_any(w0, w1, cw)) & 1u) bad = true;
        }
        sum = __dadd_rn(sum, element_product<MODE>(cw, val, i, A, x, L));
      }
      if (len >= 4 && (crc ^ 0xFFFFFFFFu) != stored) bad = true;
      if (bad) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + row, 0);
    } else {
      // element by element, repairing / reporting as the mode demands (the owning thread
      // writes a corrected element back)
      for (uint32_t i = rs; i < re; i++)
        sum = __dadd_rn(sum, element_product<MODE>(s_cols[i - cb], s_vals[i - vb], i, A, x, L));
    }
  }
  if (ModeTraits<MODE>::vec_sed) {
    if (vec_parity_bits(xrow) & 1u) log_event(L, ABFT_EV_VECTOR_SED, A.row_offset + row, VEC_SPMV_OPERAND);
    sum = vec_encode(sum);   // what is stored is what dot() and calc_xr will read
  }
  y[row] = sum;
  dot_part = __dadd_rn(dot_part, __dmul_rn(sum, xrow));
}

// DYN: the dynamic tile schedule (A.tile_counter), compiled as its own kernel so that the static one carries none of it
// FLEN != 0 (static schedule only): a stencil matrix -- every tile regular, no row longer than FLEN (5 or 7) -- gets an
// instantiation without the long-row and irregular-tile paths and with the row walk's length fixed: ~20 fewer
// instructions per warp and tile of the ~100 that are not the walk itself (this kernel's time follows its instruction
// count, see the notes at the refill below).
template <int MODE, bool DYN, int FLEN = 0>
// (crc32c with its tables in registers: 80 registers per thread, 3 CTAs per SM; everything else 64 and 4)
__global__ void __launch_bounds__(SpmvCfg<DYN>::block, ((MODE & 7) == ABFT_MODE_CRC32C && FLEN) ? 3 : 4)
spmv_kernel(MatView A, const double *__restrict__ x, double *__restrict__ y, LogRef L, ReduceOut red,
            const int *__restrict__ active, WaitSpec wait) {
#ifndef ABFT_HOST_EMU
  extern __shared__ __align__(128) unsigned char dyn_smem[];
#else
  unsigned char *const dyn_smem = emu::g_dyn_smem;
#endif
  constexpr int BLOCK = SpmvCfg<DYN>::block;
  constexpr int kWarps = BLOCK / 32;
  using StageBuf = typename SpmvCfg<DYN>::Stage;
  constexpr size_t kSpmvDynSmem = SpmvCfg<DYN>::dyn_smem;
  StageBuf *stage = reinterpret_cast<StageBuf *>(dyn_smem);
  __shared__ __align__(8) uint64_t s_full[kStages];
  __shared__ Tile s_tile[kStages];
  __shared__ uint32_t s_tidx[kStages];              // index of the staged tile (dynamic schedule: its partial's slot)
  __shared__ double s_wdot[2][kWarps];              // dynamic schedule: the warps' dot() partials of a tile
  __shared__ uint32_t s_crc[((MODE & 7) == ABFT_MODE_CRC32C) ? kCrcTabWords : 1];
  __shared__ double s_warp[kWarps];

  if (active && *active == 0) return;  // device-resident loop already converged

  const int tid = threadIdx.x;
  const uint32_t G = gridDim.x;
  CrcRegs crc_regs;
  if ((MODE & 7) == ABFT_MODE_CRC32C && FLEN) crc_regs_load(crc_regs, tid & 31);
  // Tile source.  Static: CTA b takes tiles b, b+grid, ... (the whole chip sweeps the matrix as one front; each
  // thread keeps its own dot() partial).  Dynamic (matrices with irregular tiles, whose cost per tile varies by
  // 5x): CTAs draw tiles from a counter; a tile's dot() contribution is then reduced per tile into
  // red.partials[tile], so the value published at the end is the same whichever CTA ran which tile.
  constexpr bool dynamic = DYN;
  // static: this CTA's share of the tiles; dynamic: unknown, the loop ends at the "no more tiles" marker
  const uint32_t n_my = (!DYN && A.n_tiles > blockIdx.x) ? (A.n_tiles - blockIdx.x + G - 1) / G : 0;
  Tile ahead{};            // thread 0: descriptor and index of the next tile to issue
  uint32_t ahead_idx = blockIdx.x - G;  // (static: tiles b, b+G, ...; wraps on purpose, the first fetch adds G)
  auto fetch_ahead = [&]() {
    if (DYN) ahead_idx = atomicAdd(A.tile_counter, 1u);
    else ahead_idx += G;
    if (ahead_idx < A.n_tiles) ahead = A.tiles[ahead_idx];  // consumed an iteration later: the latency overlaps a tile's work
    else ahead.info = kTileNone;
  };

  {
    // the row walk may look a few slots past the data TMA delivers: start from all-zero buffers
    // (a zero element is a valid code word in every mode)
    uint4 *z = reinterpret_cast<uint4 *>(dyn_smem);
    for (uint32_t i = tid; i < kSpmvDynSmem / 16; i += BLOCK) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&s_full[s], 1);
    fence_mbarrier_init();
  }
  __syncthreads();
  if (tid == 0) fence_proxy_async();
  if ((MODE & 7) == ABFT_MODE_CRC32C)
    for (int i = tid; i < kCrcTabWords; i += BLOCK) s_crc[i] = g_crc5_tab[i];
  __syncthreads();
  // (Tried in round 2 and removed: launching this kernel as a programmatic dependent of calc_p, so that the set-up above
  // overlaps calc_p's tail.  The stamps of tools/dist_phases.py showed the gap calc_p -> spmv shrink from 6.0 to 2.8 us
  // at 2 GPUs -- and calc_p and this kernel's own rows grow by 4.7 and 3.7 us, the early CTAs competing with the
  // predecessor's last ones: 3568 it/s with, 3605 without, profiles/r2_pdl_ab_n1_n2.json.  Round 1 saw the same.)

  const uint64_t stream_policy = l2_evict_first_policy();
  // producer (thread 0): stream tile `t` into stage `s`
  auto issue = [&](int s, const Tile &t) {
    StageBuf &B = stage[s];
    const uint32_t nrows = t.info & 0xFFFFu;
    const uint32_t rb = t.row_begin & ~3u;
    const uint32_t rbytes = ((t.row_begin + nrows + 1 + 3) & ~3u) * 4u - rb * 4u;
    uint32_t cbytes = 0, vbytes = 0;
    const uint32_t cb = t.nnz_begin & ~3u, vb = t.nnz_begin & ~1u;
    if (!(t.info & kTileLongRow)) {
      // kOverrun real elements past the tile's end come along: the unrolled row walk may look that far
      // past a row's end, and must find matched (column, value) pairs there, not stale halves
      cbytes = ((t.nnz_end + kOverrun + 3) & ~3u) * 4u - cb * 4u;
      vbytes = ((t.nnz_end + kOverrun + 1) & ~1u) * 8u - vb * 8u;
    }
    fence_proxy_async();
    mbar_expect_tx(&s_full[s], rbytes + cbytes + vbytes);
    bulk_g2s(B.rptr, A.rowptr + rb, rbytes, &s_full[s], stream_policy);
    if (cbytes) {
      bulk_g2s(B.cols, A.cols + cb, cbytes, &s_full[s], stream_policy);
      bulk_g2s(B.vals, A.values + vb, vbytes, &s_full[s], stream_policy);
    }
  };

  // (Tried in round 2 and removed: splitting the refill over the eight warps -- every warp's lane 0 streaming one part
  // of the tile and arriving on the stage's mbarrier -- so that no single warp reaches the per-tile CTA barrier ~90
  // instructions late.  It did what it was meant to, ncu: barrier stall 1.84 -> 0.69 warps per issue cycle, but the
  // part bookkeeping cost every warp ~75 instructions per tile: 219 M -> 259 M warp instructions and 0.33 -> 0.37 ms.
  // This kernel's time follows its instruction count, profiles/r2_ab_session4_split_refill.jsonl.)
  if (tid == 0) {
    fetch_ahead();
    s_tile[0] = ahead;
    s_tidx[0] = ahead_idx;
    if (ahead.info != kTileNone) issue(0, ahead);
    else mbar_arrive(&s_full[0]);  // nothing to stream: complete the phase so the waiters get to see the marker
    fetch_ahead();
  }
  if (tid == 0 && blockIdx.x == 0) phase_stamp(wait.phase, PH_SPMV_BEGIN);
  if (wait.mask) {
    // the matrix stream of the first tile is already in flight; vec[] is only touched after this
    if (tid == 0) {
      for (int s2 = 0; s2 < kMaxRanks; s2++)
        if ((wait.mask >> s2) & 1u)
          while ((int)(ld_volatile_u32(&wait.me->pflag[s2]) - wait.epoch) < 0) {
          }
      __threadfence_system();
    }
  }
  if (tid == 0 && blockIdx.x == 0) phase_stamp(wait.phase, PH_SPMV_WAIT_END);
  __syncthreads();
  if (A.rows_ecc) {
    // the first tile's row pointers; those of tile it + 1 are checked at the end of iteration it, behind the walk
    mbar_wait(&s_full[0], 0u);
    const Tile t0 = s_tile[0];
    if (t0.info != kTileNone) rows_prepass<BLOCK>(stage[0].rptr, t0, A, L, tid);
    __syncthreads();
  }

  double dot_part = 0.0;

  uint32_t it = 0;
  for (; DYN || it < n_my; it++) {
    const int st = (int)(it % kStages);
    // thread 0: bank the previous tile's dot() partial (dynamic schedule), refill the stage that tile released
    if (tid == 0) {
      if (dynamic && it > 0) {
        // the warps' sums, added in warp order
        const double *wd = s_wdot[(it - 1) & 1u];
        double tp = wd[0];
#pragma unroll
        for (int k = 1; k < kWarps; k++) tp = __dadd_rn(tp, wd[k]);
        red.partials[s_tidx[(it - 1) % kStages]] = tp;
      }
      const int js = (int)((it + kStages - 1) % kStages);
      s_tile[js] = ahead;
      s_tidx[js] = ahead_idx;
      if (ahead.info != kTileNone) {
        issue(js, ahead);
        fetch_ahead();
      } else {
        mbar_arrive(&s_full[js]);  // nothing to stream: complete the phase so the waiters get to see the marker
      }
    }
    mbar_wait(&s_full[st], (it / kStages) & 1u);
    const Tile t = s_tile[st];  // (written by thread 0 before it armed the barrier)
    if (DYN && t.info == kTileNone) break;
    const bool irregular_tile = FLEN == 0 && !(t.info & (kTileLongRow | kTileRegular)) && (MODE & 7) != ABFT_MODE_CRC32C;
    const StageBuf &B = stage[st];
    if (dynamic) dot_part = 0.0;  // reduced per tile below
    const uint32_t nrows = t.info & 0xFFFFu;
    const uint32_t n = t.nnz_end - t.nnz_begin;
    const uint32_t rb = t.row_begin & ~3u, cb = t.nnz_begin & ~3u, vb = t.nnz_begin & ~1u;
    const bool has_row = (uint32_t)tid < nrows;

    if (FLEN == 0 && (t.info & kTileLongRow)) {
      // ---------------- one long row: stream it from global memory in chunks ----------------
      double acc = 0.0;
      bool flagged = false;
      // CONSTRAINTS looks at the row's own pointers (staged with the tile), like the reference (CSR/CPUContext.cpp:
      // 173-182): the tile's range is what they were at create_matrix.  If they have moved, thread 0 does the
      // reference's row checks and, when they pass, its walk over [start, end) from global memory -- one by one,
      // in its order (rare: a corrupted rows[]).  Uniform over the CTA (every thread reads the same two words).
      bool moved = false;
      if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) {
        const uint32_t prs = B.rptr[t.row_begin - rb], pre = B.rptr[t.row_begin + 1 - rb];
        moved = prs != t.nnz_begin || pre != t.nnz_end;
        if (moved && tid == 0) {
          const uint32_t grow = A.row_offset + t.row_begin;
          if (pre > A.nnz) log_event(L, ABFT_EV_ROW_SIZE, grow, 0);
          else if (pre < prs) log_event(L, ABFT_EV_ROW_ORDER, grow, 0);
          else
            for (uint32_t i = prs; i < pre; i++) {
              const uint32_t col = A.cols[i];
              if (col >= A.n_cols) { log_event(L, ABFT_EV_COL_SIZE, A.nnz_offset + i, grow); break; }
              if (i < pre - 1 && A.cols[i + 1] <= col) { log_event(L, ABFT_EV_COL_ORDER, A.nnz_offset + i, grow); break; }
              acc = __dadd_rn(acc, __dmul_rn(A.values[i], __ldg(x + col)));
            }
        }
      }
      for (uint32_t base = tid; !moved && base < n; base += kLongBatch * BLOCK) {
        uint32_t li[kLongBatch], cw[kLongBatch];
        double val[kLongBatch], pr[kLongBatch];
        bool ok[kLongBatch];
#pragma unroll
        for (int k = 0; k < kLongBatch; k++) {
          ok[k] = (k == 0) || (base + k * BLOCK < n);
          li[k] = t.nnz_begin + (ok[k] ? base + k * BLOCK : base);
          cw[k] = ld_stream(A.cols + li[k]);
          val[k] = ld_stream(A.values + li[k]);
        }
        if ((MODE & 7) == ABFT_MODE_CONSTRAINTS && !flagged) {
#pragma unroll
          for (int k = 0; k < kLongBatch; k++)
            if (ok[k] && !flagged) {
              if (cw[k] >= A.n_cols) { log_event(L, ABFT_EV_COL_SIZE, A.nnz_offset + li[k], A.row_offset + t.row_begin); flagged = true; }
              else if (li[k] + 1 < t.nnz_end && A.cols[li[k] + 1] <= cw[k]) { log_event(L, ABFT_EV_COL_ORDER, A.nnz_offset + li[k], A.row_offset + t.row_begin); flagged = true; }
            }
        }
        batch_products<MODE, kLongBatch>(li, cw, val, ok, A, x, L, pr);
#pragma unroll
        for (int k = 0; k < kLongBatch; k++)
          if (ok[k]) acc = __dadd_rn(acc, pr[k]);
      }
      if ((MODE & 7) == ABFT_MODE_CRC32C && tid == 0) {
        // long rows are rare: one thread walks the row for the CRC
        uint32_t crc = 0xFFFFFFFFu, stored = 0;
        for (uint32_t e = 0; e < n; e++) {
          const uint32_t i = t.nnz_begin + e;
          const uint32_t cw = A.cols[i];
          uint32_t w0, w1;
          split_value(A.values[i], w0, w1);
          if (e < 4) stored |= (cw >> 24) << (8 * e);
          crc = crc_element(crc, w0, w1, cw & 0x00FFFFFFu, s_crc);
        }
        if ((crc ^ 0xFFFFFFFFu) != stored) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + t.row_begin, 0);
      }
      double total = block_sum<BLOCK>(acc, s_warp);
      if (tid == 0) {
        const double xr = x[A.row_offset + t.row_begin];
        if (ModeTraits<MODE>::vec_sed) {
          if (vec_parity_bits(xr) & 1u) log_event(L, ABFT_EV_VECTOR_SED, A.row_offset + t.row_begin, VEC_SPMV_OPERAND);
          total = vec_encode(total);
        }
        y[t.row_begin] = total;
        dot_part = __dadd_rn(dot_part, __dmul_rn(total, xr));
      }
    } else {
      const bool irregular = irregular_tile;
      auto constraints_careful = [&](uint32_t grow, uint32_t rs, uint32_t re) -> double {
        return constraints_careful_row(grow, rs, re, t, B.cols, B.vals, cb, vb, !irregular, A, x, L);
      };
      if (irregular) {
        // ---------------- irregular tile ----------------
        // The tile's rows are dealt to the warps in equal runs of consecutive rows (warp w owns rows
        // [w*c, (w+1)*c), c = ceil(nrows / warps) <= 32, one per lane), so every warp has its share whatever the tile's
        // row count, and a warp finishes its own rows -- short and long -- without a CTA-wide barrier.  The dot()
        // operand is fetched now, ahead of the element phase.
        const int lane = tid & 31, warp = tid >> 5;
        const uint32_t per_warp = (nrows + kWarps - 1) / kWarps;
        const uint32_t q = (uint32_t)warp * per_warp + lane;
        const bool has_q = (uint32_t)lane < per_warp && q < nrows;
        const uint32_t qrow = t.row_begin + q;
        const double xq = has_q ? x[A.row_offset + qrow] : 0.0;

        // element phase: product -> shared memory (in place).  kElemBatch elements per thread at a time: all
        // shared-memory loads, then all checks and gathers, then the stores.  The checks only OR their verdict
        // into `suspect`; a batch that trips goes through element_product one element at a time (rare).
        double *prod = const_cast<double *>(B.vals);
        for (uint32_t e0 = tid; e0 < n; e0 += kElemBatch * BLOCK) {
          uint32_t li[kElemBatch], cw[kElemBatch];
          double v[kElemBatch], pr[kElemBatch];
          bool ok[kElemBatch];
#pragma unroll
          for (int k = 0; k < kElemBatch; k++) {
            // slots past the tile's end repeat this thread's first element (loaded, checked, never stored)
            ok[k] = (k == 0) || (e0 + k * BLOCK < n);
            li[k] = t.nnz_begin + (ok[k] ? e0 + k * BLOCK : e0);
            cw[k] = B.cols[li[k] - cb];
            v[k] = B.vals[li[k] - vb];
          }
          batch_products<MODE, kElemBatch>(li, cw, v, ok, A, x, L, pr);
#pragma unroll
          for (int k = 0; k < kElemBatch; k++)
            if (ok[k]) prod[li[k] - vb] = pr[k];
        }
        __syncthreads();

        // row phase: a thread adds its row's products in element order (the reference's order: same bits) if
        // the row has at most kSeqRowMax elements ...
        uint32_t rs = 0, re = 0;
        if (has_q) {
          rs = B.rptr[qrow - rb];
          re = B.rptr[qrow + 1 - rb];
          if ((MODE & 7) != ABFT_MODE_CONSTRAINTS) {
            // these modes trust rows[] like the reference; clamp so a bad pointer cannot fault
            rs = min(max(rs, t.nnz_begin), t.nnz_end);
            re = min(max(re, rs), t.nnz_end);
          }
        }
        // (CONSTRAINTS: a corrupted rows[] can send the row outside the staged window; then the careful path
        // that mirrors the reference's checks one by one)
        const bool in_window = rs >= t.nnz_begin && re <= t.nnz_end && rs <= re;
        const bool is_long = has_q && in_window && (re - rs > kSeqRowMax);
        double sum = 0.0;
        if (has_q && !is_long) {
          if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) {
            bool bad = !in_window;
            if (!bad)
              for (uint32_t i = rs; i < re; i++) {
                const uint32_t col = B.cols[i - cb];
                bad |= (col >= A.n_cols);
                if (i + 1 < re) bad |= (B.cols[i + 1 - cb] <= col);
                sum = __dadd_rn(sum, B.vals[i - vb]);
              }
            if (bad) sum = constraints_careful(qrow, rs, re);
          } else {
#pragma unroll 4
            for (uint32_t i = rs; i < re; i++) sum = __dadd_rn(sum, B.vals[i - vb]);
          }
        }
        // ... and the warp then takes its lanes' longer rows one at a time: lane-strided partial sums + a fixed
        // shuffle tree (deterministic; not the sequential order, so not bit-identical to the CPU sum)
        uint32_t longs = __ballot_sync(0xffffffffu, is_long);
        while (longs) {
          const int src = __ffs(longs) - 1;
          longs &= longs - 1;
          const uint32_t ls = __shfl_sync(0xffffffffu, rs, src), le = __shfl_sync(0xffffffffu, re, src);
          double part = 0.0;
          bool bad = false;
          for (uint32_t i = ls + lane; i < le; i += 32) {
            if ((MODE & 7) == ABFT_MODE_CONSTRAINTS) {
              // CSR/CPUContext.cpp:188-200; a violation sends the whole row down the careful path
              const uint32_t c = B.cols[i - cb];
              bad |= (c >= A.n_cols);
              if (i + 1 < le) bad |= (B.cols[i + 1 - cb] <= c);
            }
            part = __dadd_rn(part, B.vals[i - vb]);
          }
          part = warp_sum(part);
          part = __shfl_sync(0xffffffffu, part, 0);
          const bool any_bad = ((MODE & 7) == ABFT_MODE_CONSTRAINTS) && __any_sync(0xffffffffu, bad);
          if (lane == src) sum = any_bad ? constraints_careful(qrow, rs, re) : part;
        }
        if (has_q) {
          if (ModeTraits<MODE>::vec_sed) {
            if (vec_parity_bits(xq) & 1u) log_event(L, ABFT_EV_VECTOR_SED, A.row_offset + qrow, VEC_SPMV_OPERAND);
            sum = vec_encode(sum);
          }
          y[qrow] = sum;
          dot_part = __dadd_rn(dot_part, __dmul_rn(sum, xq));
        }
      } else {
        // ---------------- regular tile: thread-per-row straight from the staged stream ----------------
        regular_row<MODE, FLEN>(t, (uint32_t)tid, has_row, B.rptr, B.cols, B.vals, A, x, y, L, s_crc, dot_part, crc_regs);
      }
    }
    if (dynamic) {
      const double wp = warp_sum(dot_part);
      if ((tid & 31) == 0) s_wdot[it & 1u][tid >> 5] = wp;
    }
    if (A.rows_ecc) {
      // the NEXT tile's row pointers (its stream was issued at the top of this iteration and has normally landed
      // behind the walk): checked here so that the barrier below also publishes the repaired, stripped pointers
      const int sn = (int)((it + 1) % kStages);
      mbar_wait(&s_full[sn], ((it + 1) / kStages) & 1u);
      const Tile tn = s_tile[sn];
      if (tn.info != kTileNone && (DYN || it + 1 < n_my)) rows_prepass<BLOCK>(stage[sn].rptr, tn, A, L, tid);
    }
    __syncthreads();  // everyone is done with stage `st`: it may be refilled
  }
  // (dynamic schedule: the loop only ends at the marker, whose iteration has already banked the last tile's partial)
  // fused dot(vec, result): CTA partials (static) or the per-tile partials (dynamic), then the grid-wide deterministic sum
  const double mine = dynamic ? 0.0 : block_sum<BLOCK>(dot_part, s_warp);
  const bool last = grid_sum_finalize<BLOCK>(mine, red, s_warp, dynamic ? A.n_tiles : 0u);
  if (dynamic && last && tid == 0) *A.tile_counter = 0u;  // every CTA has drawn its last tile: ready for the next launch
}


}  // namespace abft
