// spmv.cuh -- CSR SpMV with the per-element ABFT check/correct fused into the
// coalesced column/value stream (reference: CSR/CPUContext.cpp:115-133 NONE,
// 162-207 CONSTRAINTS, 214-245 SED, 252-289 SEC7, 297-345 SEC8, 353-411 SECDED;
// CRC32C is this repository's own mode).
//
// Work decomposition ("tiles", built once in create_matrix): a tile is a run of
// whole rows with at most kTileNnz non-zeros and at most kBlock rows; one CTA
// per tile.
//   element phase  every lane owns elements e = u*kBlock + tid of the tile's
//                  nnz range: coalesced streaming loads of cols/values, the
//                  mode's ECC check (the owning lane corrects + writes back +
//                  logs), gather of vec[col], product -> shared memory.
//   row phase      thread r sums the products of row r from shared memory IN
//                  ELEMENT ORDER with separate multiply and add, i.e. exactly
//                  the reference's `tmp += values[i] * vec[col]` (no FMA in the
//                  reference build), so result[] is bit-identical to the CPU
//                  target; then the fused dot(vec, result) partial.
// A row longer than kTileNnz is a tile of its own, streamed in chunks with a
// tree reduction (not bit-identical to the sequential CPU sum; within 1e-10).
// CONSTRAINTS element events carry their global row in `bit` (the host orders
// them against ROW_* events with it and clears it before reporting).
#pragma once
#include "common.cuh"
#include "ecc.cuh"

namespace abft {

constexpr int kEpt = 8;                      // elements per thread per tile
constexpr int kTileNnz = kBlock * kEpt;      // 2048
constexpr int kTileRows = kBlock;            // 256

struct Tile {
  uint32_t row_begin, row_end;  // local rows
  uint32_t nnz_begin, nnz_end;  // local element indices
};

struct MatView {
  const uint32_t *rowptr;  // [n_rows+1] local element indices
  uint32_t *cols;          // [nnz] column word incl. check bits
  double *values;          // [nnz]
  const Tile *tiles;
  uint32_t n_rows, nnz;
  uint32_t n_cols;      // global N: bound for the gather
  uint32_t row_offset;  // global index of local row 0
  uint32_t nnz_offset;  // global index of local element 0 (for reported indices)
};

__device__ __forceinline__ void split_value(double v, uint32_t &w0, uint32_t &w1) {
  w0 = (uint32_t)__double2loint(v);
  w1 = (uint32_t)__double2hiint(v);
}

// Slow path shared by SEC7 / SEC8 / SECDED: undo the flip the syndrome points
// at (CSR/CPUContext.cpp:272-278, 320-324, 376-380).  A multi-bit fault can
// yield bit >= 96, where the reference writes past its 12-byte element (UB);
// here the bit is reported and the element left alone.
__device__ __noinline__ void correct_from_syndrome(uint32_t &w0, uint32_t &w1, uint32_t &col, uint32_t syndrome,
                                                   uint32_t gindex, LogRef L) {
  uint32_t bit = flipped_bit(syndrome);
  if (bit < 32) w0 ^= 1u << bit;
  else if (bit < 64) w1 ^= 1u << (bit - 32);
  else if (bit < 96) col ^= 1u << (bit - 64);
  log_event(L, ABFT_EV_CORRECTED_BIT, gindex, bit);
}

// The per-element check of one mode.  Returns true if the element was modified
// and must be written back.  `poison` is set when a detect-only fault makes the
// column untrustworthy.
template <int MODE>
__device__ __forceinline__ bool check_element(uint32_t &w0, uint32_t &w1, uint32_t &col, uint32_t gindex,
                                              const LogRef &L) {
  if (MODE == ABFT_MODE_SED) {
    if (overall_parity(w0, w1, col)) log_event(L, ABFT_EV_SED_DETECTED, gindex, 0);
    return false;
  }
  if (MODE == ABFT_MODE_SEC7) {
    if (syndrome7_any(w0, w1, col)) {
      correct_from_syndrome(w0, w1, col, syndrome7(w0, w1, col), gindex, L);
      return true;
    }
    return false;
  }
  if (MODE == ABFT_MODE_SEC8) {
    if (overall_parity(w0, w1, col)) {
      uint32_t s = syndrome7(w0, w1, col);
      if (s) correct_from_syndrome(w0, w1, col, s, gindex, L);
      else { col ^= 1u << 24; log_event(L, ABFT_EV_CORRECTED_PARITY, gindex, 0); }
      return true;
    }
    return false;
  }
  if (MODE == ABFT_MODE_SECDED) {
    const uint32_t odd = overall_parity(w0, w1, col);
    const uint32_t syn = syndrome7_any(w0, w1, col);
    if (odd | syn) {
      if (odd) {
        uint32_t s = syndrome7(w0, w1, col);
        if (s) correct_from_syndrome(w0, w1, col, s, gindex, L);
        else { col ^= 1u << 24; log_event(L, ABFT_EV_CORRECTED_PARITY, gindex, 0); }
        return true;
      }
      log_event(L, ABFT_EV_DOUBLE_BIT, gindex, 0);
    }
    return false;
  }
  return false;
}

template <int MODE>
struct ModeTraits {
  static constexpr bool masks_column = MODE >= ABFT_MODE_SED;  // check bits in column bits 24..31
  static constexpr bool keeps_cols = MODE == ABFT_MODE_CONSTRAINTS || MODE == ABFT_MODE_CRC32C;
};

// CRC-32C of one 12-byte element appended to a running (raw) state; slicing-by-4
// tables in shared memory: tab[s*256 + b].
__device__ __forceinline__ uint32_t crc_word(uint32_t crc, uint32_t w, const uint32_t *tab) {
  const uint32_t x = crc ^ w;
  return tab[768 + (x & 0xFF)] ^ tab[512 + ((x >> 8) & 0xFF)] ^ tab[256 + ((x >> 16) & 0xFF)] ^ tab[x >> 24];
}

__device__ uint32_t g_crc_tab[1024];  // filled by the host at library init

template <int MODE>
__global__ void __launch_bounds__(kBlock, (MODE == ABFT_MODE_CRC32C) ? 2 : 3)
spmv_kernel(MatView A, const double *__restrict__ x, double *__restrict__ y, LogRef L, ReduceOut red,
            const int *__restrict__ active) {
  __shared__ __align__(16) double s_prod[kTileNnz];
  __shared__ uint32_t s_col[ModeTraits<MODE>::keeps_cols ? kTileNnz : 1];
  __shared__ double s_val[(MODE == ABFT_MODE_CRC32C) ? kTileNnz : 1];
  __shared__ uint32_t s_crc[(MODE == ABFT_MODE_CRC32C) ? 1024 : 1];
  __shared__ double s_warp[kBlock / 32];

  if (active && *active == 0) return;  // device-resident loop already converged

  const int tid = threadIdx.x;
  const Tile t = A.tiles[blockIdx.x];
  const uint32_t nrows = t.row_end - t.row_begin;
  const uint32_t n = t.nnz_end - t.nnz_begin;

  if (MODE == ABFT_MODE_CRC32C) {
    for (int i = tid; i < 1024; i += kBlock) s_crc[i] = g_crc_tab[i];
    __syncthreads();
  }

  // row pointers of my row, requested early so their latency overlaps the stream
  uint32_t rs = 0, re = 0;
  const bool has_row = (uint32_t)tid < nrows;
  if (has_row) {
    rs = A.rowptr[t.row_begin + tid];
    re = A.rowptr[t.row_begin + tid + 1];
  }

  double dot_part = 0.0;

  if (n <= (uint32_t)kTileNnz) {
    // ---------------- element phase ----------------
    uint32_t c[kEpt];
    double v[kEpt];
#pragma unroll
    for (int u = 0; u < kEpt; u++) {
      const uint32_t e = u * kBlock + tid;
      c[u] = 0;
      v[u] = 0.0;
      if (e < n) {
        c[u] = ld_stream(A.cols + t.nnz_begin + e);
        v[u] = ld_stream(A.values + t.nnz_begin + e);
      }
    }
    double xv[kEpt];
#pragma unroll
    for (int u = 0; u < kEpt; u++) {
      const uint32_t e = u * kBlock + tid;
      xv[u] = 0.0;
      if (e < n) {
        if (MODE >= ABFT_MODE_SED && MODE <= ABFT_MODE_SECDED) {
          uint32_t w0, w1;
          split_value(v[u], w0, w1);
          if (check_element<MODE>(w0, w1, c[u], A.nnz_offset + t.nnz_begin + e, L)) {
            v[u] = __hiloint2double((int)w1, (int)w0);
            A.cols[t.nnz_begin + e] = c[u];
            A.values[t.nnz_begin + e] = v[u];
          }
        }
        const uint32_t col = ModeTraits<MODE>::masks_column ? (c[u] & 0x00FFFFFFu) : c[u];
        if (col < A.n_cols) xv[u] = __ldg(x + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kEpt; u++) {
      const uint32_t e = u * kBlock + tid;
      if (e < n) {
        s_prod[e] = __dmul_rn(v[u], xv[u]);
        if (ModeTraits<MODE>::keeps_cols) s_col[e] = c[u];
        if (MODE == ABFT_MODE_CRC32C) s_val[e] = v[u];
      }
    }
    __syncthreads();

    // ---------------- row phase ----------------
    if (has_row) {
      const uint32_t row = t.row_begin + tid;
      bool row_ok = true;
      if (MODE == ABFT_MODE_CONSTRAINTS) {
        // CSR/CPUContext.cpp:173-182
        if (re > A.nnz) { log_event(L, ABFT_EV_ROW_SIZE, A.row_offset + row, 0); row_ok = false; }
        else if (re < rs) { log_event(L, ABFT_EV_ROW_ORDER, A.row_offset + row, 0); row_ok = false; }
      } else {
        // other modes trust rows[] like the reference; clamp so a bad pointer cannot fault
        rs = min(max(rs, t.nnz_begin), t.nnz_end);
        re = min(max(re, rs), t.nnz_end);
      }
      double sum = 0.0;
      if (row_ok) {
        if (MODE == ABFT_MODE_CONSTRAINTS) {
          for (uint32_t i = rs; i < re; i++) {
            // elements normally sit in the tile's shared-memory window; a corrupted
            // row pointer can send the row outside it, then go to global memory
            const bool in = (i >= t.nnz_begin) && (i < t.nnz_end);
            const uint32_t col = in ? s_col[i - t.nnz_begin] : A.cols[i];
            if (col >= A.n_cols) { log_event(L, ABFT_EV_COL_SIZE, A.nnz_offset + i, A.row_offset + row); break; }
            if (i < re - 1) {
              const bool in2 = (i + 1 >= t.nnz_begin) && (i + 1 < t.nnz_end);
              const uint32_t nxt = in2 ? s_col[i + 1 - t.nnz_begin] : A.cols[i + 1];
              if (nxt <= col) { log_event(L, ABFT_EV_COL_ORDER, A.nnz_offset + i, A.row_offset + row); break; }
            }
            const double prod = in ? s_prod[i - t.nnz_begin] : __dmul_rn(A.values[i], __ldg(x + col));
            sum = __dadd_rn(sum, prod);
          }
        } else {
          if (MODE == ABFT_MODE_CRC32C) {
            const uint32_t len = re - rs;
            bool bad = false;
            if (len >= 4) {
              uint32_t crc = 0xFFFFFFFFu, stored = 0;
              for (uint32_t i = rs; i < re; i++) {
                const uint32_t k = i - t.nnz_begin;
                uint32_t w0, w1;
                split_value(s_val[k], w0, w1);
                const uint32_t cw = s_col[k];
                if (i - rs < 4) stored |= (cw >> 24) << (8 * (i - rs));
                crc = crc_word(crc, w0, s_crc);
                crc = crc_word(crc, w1, s_crc);
                crc = crc_word(crc, cw & 0x00FFFFFFu, s_crc);
              }
              bad = (crc ^ 0xFFFFFFFFu) != stored;
            } else {
              for (uint32_t i = rs; i < re; i++) {
                const uint32_t k = i - t.nnz_begin;
                uint32_t w0, w1;
                split_value(s_val[k], w0, w1);
                if (overall_parity(w0, w1, s_col[k]) | syndrome7_any(w0, w1, s_col[k])) bad = true;
              }
            }
            if (bad) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + row, 0);
          }
          for (uint32_t i = rs - t.nnz_begin; i < re - t.nnz_begin; i++) sum = __dadd_rn(sum, s_prod[i]);
        }
      }
      y[row] = sum;
      dot_part = __dmul_rn(sum, x[A.row_offset + row]);
    }
  } else {
    // ---------------- one long row: stream it in chunks ----------------
    // (tile builder guarantees nrows == 1 here)
    double acc = 0.0;
    bool flagged = false;
    uint32_t crc_dummy = 0;
    (void)crc_dummy;
    for (uint32_t base = 0; base < n; base += kBlock) {
      const uint32_t e = base + tid;
      if (e < n) {
        const uint32_t i = t.nnz_begin + e;
        uint32_t cw = ld_stream(A.cols + i);
        double val = ld_stream(A.values + i);
        if (MODE >= ABFT_MODE_SED && MODE <= ABFT_MODE_SECDED) {
          uint32_t w0, w1;
          split_value(val, w0, w1);
          if (check_element<MODE>(w0, w1, cw, A.nnz_offset + i, L)) {
            val = __hiloint2double((int)w1, (int)w0);
            A.cols[i] = cw;
            A.values[i] = val;
          }
        }
        const uint32_t col = ModeTraits<MODE>::masks_column ? (cw & 0x00FFFFFFu) : cw;
        if (MODE == ABFT_MODE_CONSTRAINTS && !flagged) {
          if (col >= A.n_cols) { log_event(L, ABFT_EV_COL_SIZE, A.nnz_offset + i, A.row_offset + t.row_begin); flagged = true; }
          else if (e + 1 < n && A.cols[i + 1] <= col) { log_event(L, ABFT_EV_COL_ORDER, A.nnz_offset + i, A.row_offset + t.row_begin); flagged = true; }
        }
        if (col < A.n_cols) acc = __dadd_rn(acc, __dmul_rn(val, __ldg(x + col)));
      }
    }
    if (MODE == ABFT_MODE_CRC32C && tid == 0) {
      // long rows are rare: one thread walks the row (global memory) for the CRC
      uint32_t crc = 0xFFFFFFFFu, stored = 0;
      for (uint32_t e = 0; e < n; e++) {
        const uint32_t i = t.nnz_begin + e;
        const uint32_t cw = A.cols[i];
        uint32_t w0, w1;
        split_value(A.values[i], w0, w1);
        if (e < 4) stored |= (cw >> 24) << (8 * e);
        crc = crc_word(crc, w0, s_crc);
        crc = crc_word(crc, w1, s_crc);
        crc = crc_word(crc, cw & 0x00FFFFFFu, s_crc);
      }
      if ((crc ^ 0xFFFFFFFFu) != stored) log_event(L, ABFT_EV_CRC_ROW, A.row_offset + t.row_begin, 0);
    }
    const double total = block_sum(acc, s_warp);
    if (tid == 0) {
      y[t.row_begin] = total;
      dot_part = __dmul_rn(total, x[A.row_offset + t.row_begin]);
    }
  }

  // fused dot(vec, result): CTA partial, then the grid-wide deterministic sum
  const double mine = block_sum(dot_part, s_warp);
  grid_sum_finalize(mine, red, s_warp);
}

}  // namespace abft
