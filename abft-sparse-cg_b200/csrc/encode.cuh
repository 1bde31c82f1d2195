// encode.cuh -- device side of create_matrix (reference: CSR/CPUContext.cpp:11-44,
// generate_ecc_bits :209-212, :247-250, :291-295, :347-351) and of inject_bitflip
// (:150-157).  The reference does this in one serial host loop; here the triplets
// are uploaded once and converted / encoded by element-parallel kernels.
#pragma once
#include "common.cuh"
#include "ecc.cuh"
#include "spmv.cuh"

namespace abft {

// Row pointers from row-sorted triplets.  rowptr[k] = index of the first element
// whose row is >= k (CSR/CPUContext.cpp:36-39); rows past the last element get
// nnz (the reference only sets rows[N], :41, leaving trailing empty rows
// undefined).  coo_rows are GLOBAL row numbers of a partition starting at row_offset.
__global__ void __launch_bounds__(kBlock)
rowptr_kernel(const uint32_t *__restrict__ coo_rows, uint32_t nnz, uint32_t n_rows, uint32_t row_offset,
              uint32_t *__restrict__ rowptr) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i <= nnz; i += stride) {
    // element i closes rows (prev, cur]; the virtual element nnz closes everything left
    const uint32_t cur = (i < nnz) ? min(coo_rows[i] - row_offset, n_rows) : n_rows;
    const uint32_t first = (i == 0) ? 0u : min(coo_rows[i - 1] - row_offset, n_rows) + 1u;
    for (uint32_t k = first; k <= cur; k++) rowptr[k] = i;
  }
}

// per-element check bits (modes SED, SEC7, SEC8, SECDED)
__global__ void __launch_bounds__(kBlock)
encode_elements_kernel(int mode, uint32_t *__restrict__ cols, const double *__restrict__ values, uint32_t nnz) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < nnz; i += stride) {
    uint32_t w0, w1;
    split_value(values[i], w0, w1);
    cols[i] = encode_column(mode, w0, w1, cols[i]);
  }
}

// CRC32C row code (repo-defined): one thread per row.
//   len >= 4: CRC-32C over {8 value bytes, 4 bytes of (col & 0xFFFFFF)} per element,
//             CRC byte k in bits 24..31 of the row's k-th column word.
//   len <  4: every element carries the SECDED check byte instead.
__global__ void __launch_bounds__(kBlock)
encode_crc_rows_kernel(const uint32_t *__restrict__ rowptr, uint32_t *__restrict__ cols,
                       const double *__restrict__ values, uint32_t n_rows) {
  __shared__ uint32_t s_crc[1024];
  for (int i = threadIdx.x; i < 1024; i += kBlock) s_crc[i] = g_crc_tab[i];
  __syncthreads();
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t row = blockIdx.x * kBlock + threadIdx.x; row < n_rows; row += stride) {
    const uint32_t rs = rowptr[row], re = rowptr[row + 1];
    if (re - rs >= 4) {
      uint32_t crc = 0xFFFFFFFFu;
      for (uint32_t i = rs; i < re; i++) {
        uint32_t w0, w1;
        split_value(values[i], w0, w1);
        crc = crc_word(crc, w0, s_crc);
        crc = crc_word(crc, w1, s_crc);
        crc = crc_word(crc, cols[i] & 0x00FFFFFFu, s_crc);
      }
      crc ^= 0xFFFFFFFFu;
      for (uint32_t k = 0; k < 4; k++) cols[rs + k] = (cols[rs + k] & 0x00FFFFFFu) | (((crc >> (8 * k)) & 0xFFu) << 24);
    } else {
      for (uint32_t i = rs; i < re; i++) {
        uint32_t w0, w1;
        split_value(values[i], w0, w1);
        cols[i] = encode_column(ABFT_MODE_SECDED, w0, w1, cols[i]);
      }
    }
  }
}

// row-pointer protection (rowsecc.cuh): check bits into bits 28..31 of every pair {rowptr[2k], rowptr[2k+1]};
// n_words is even (the array is padded with zeros, which the last pair may include)
__global__ void __launch_bounds__(kBlock) encode_rowptr_pairs_kernel(uint32_t *__restrict__ rowptr, uint32_t n_words) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t k = 2 * (blockIdx.x * kBlock + threadIdx.x); k + 1 < n_words; k += 2 * stride) {
    uint32_t w0 = rowptr[k], w1 = rowptr[k + 1];
    rows_encode_pair(w0, w1);
    rowptr[k] = w0;
    rowptr[k + 1] = w1;
  }
}

// the 4-byte read-modify-write of inject_bitflip (CSR/CPUContext.cpp:150-157)
__global__ void flip_bit_kernel(uint32_t *cols, double *values, uint32_t index, int bit) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (bit < 64) {
    uint32_t *words = reinterpret_cast<uint32_t *>(values + index);
    words[bit / 32] ^= 1u << (bit % 32);
  } else {
    cols[index] ^= 1u << (bit % 32);
  }
}

__global__ void poke_u32_kernel(uint32_t *p, uint32_t value) { *p = value; }
__global__ void flip_word_bit_kernel(uint32_t *p, int bit) { *p ^= 1u << bit; }

}  // namespace abft
