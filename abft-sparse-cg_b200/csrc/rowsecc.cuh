// rowsecc.cuh -- SECDED protection of the CSR row pointers (scope row f.2 of DESIGN.md: "protection of rows[]").
//
// NOT IN THE REFERENCE: its modes protect cols[] / values[] only and trust rows[] (CSR/CPUContext.cpp:123-124; only
// `constraints` looks at it).  This is this repository's own definition, in the style of the element codes of
// CSR/ecc.h, and -- like crc32c -- "parity unpinned": checked against oracle/rowptr_ecc.c only.
// Enabled per matrix with ABFT_CUDA_PROTECT_ROWS=secded at create_matrix (tests/test_gpu_rowptr.py on the device,
// tests/test_rowptr_ecc.py on the CPU emulator).
//
// Code word = one PAIR of row pointers {rows[2k], rows[2k+1]}: 56 data bits (bits 0..27 of either word: nnz < 2^28) in
// the non-power-of-two Hamming positions 3..62, six Hamming parities P1..P6 (position 2^(p-1)) in word0 bits 28..31 and
// word1 bits 28..29, the overall parity of all 64 bits in word1 bit 30; word1 bit 31 stays 0.  Code-word bit b = bit
// b%32 of word b/32.
//   odd overall parity:  syndrome s != 0 -> the bit at position s is flipped back   ("corrected bit b of row pointer pair k")
//                        s == 0          -> the overall parity bit (or the spare bit 63) is
//   even overall parity: s != 0          -> "double-bit error detected in row pointer pair k", fatal
// spmv checks every pair of the row pointers it stages BEFORE it looks at them (corrections are written back), so the
// modes' own logic only ever sees clean pointers; the host reports row-pointer events ahead of element events.
#pragma once
#include <stdint.h>

namespace abft {

constexpr uint32_t kRowsDataMask = 0x0FFFFFFFu;

__host__ __device__ constexpr bool rows_is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

// mask of Hamming parity p (1..6) over word (0..1), including the parity bit itself
__host__ __device__ constexpr uint32_t rows_mask(int p, int word) {
  uint32_t m = 0, pos = 3;
  for (int d = 0; d < 56; d++) {
    while (rows_is_pow2(pos)) pos++;
    if (((pos >> (p - 1)) & 1u) && d / 28 == word) m |= 1u << (d % 28);
    pos++;
  }
  if (p <= 4 && word == 0) m |= 1u << (27 + p);
  if (p >= 5 && word == 1) m |= 1u << (23 + p);
  return m;
}

template <int P>
struct RM {
  static constexpr uint32_t w0 = rows_mask(P, 0), w1 = rows_mask(P, 1);
};

template <int P>
__host__ __device__ inline uint32_t rows_parity_bit(uint32_t w0, uint32_t w1) {
  uint32_t t = (w0 & RM<P>::w0) ^ (w1 & RM<P>::w1);
#ifdef __CUDA_ARCH__
  return (uint32_t)__popc(t) & 1u;
#else
  return (uint32_t)__builtin_popcount(t) & 1u;
#endif
}

// bit p-1 set iff Hamming parity p is odd: the code position of a single flipped bit
__host__ __device__ inline uint32_t rows_syndrome(uint32_t w0, uint32_t w1) {
  return rows_parity_bit<1>(w0, w1) | (rows_parity_bit<2>(w0, w1) << 1) | (rows_parity_bit<3>(w0, w1) << 2) |
         (rows_parity_bit<4>(w0, w1) << 3) | (rows_parity_bit<5>(w0, w1) << 4) | (rows_parity_bit<6>(w0, w1) << 5);
}

__host__ __device__ inline uint32_t rows_overall(uint32_t w0, uint32_t w1) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popc(w0 ^ w1) & 1u;
#else
  return (uint32_t)__builtin_popcount(w0 ^ w1) & 1u;
#endif
}

// check bits for two clean pointers (< 2^28 each)
__host__ __device__ inline void rows_encode_pair(uint32_t &w0, uint32_t &w1) {
  w0 &= kRowsDataMask;
  w1 &= kRowsDataMask;
  const uint32_t s = rows_syndrome(w0, w1);
  w0 |= (s & 0xFu) << 28;
  w1 |= ((s >> 4) & 0x3u) << 28;
  w1 |= rows_overall(w0, w1) << 30;
}

// code position (1..63) -> code-word bit (0..63), or 64 if no bit lives there
__host__ __device__ inline uint32_t rows_bit_of_position(uint32_t s) {
  if (rows_is_pow2(s)) {
    uint32_t p = 1;
    while ((1u << (p - 1)) != s) p++;
    return p <= 4 ? 27 + p : 32 + 23 + p;
  }
  uint32_t lg = 0;
  while ((2u << lg) <= s) lg++;          // floor(log2 s)
  const uint32_t d = s - lg - 2;          // rank among the non-powers-of-two from 3 on
  if (d >= 56) return 64;
  return d < 28 ? d : 32 + (d - 28);
}

// 0: clean; 1: corrected (w0/w1 repaired, *bit = the code-word bit); 2: uncorrectable
__host__ __device__ inline int rows_check_pair(uint32_t &w0, uint32_t &w1, uint32_t *bit) {
  const uint32_t odd = rows_overall(w0, w1), s = rows_syndrome(w0, w1);
  if (!odd) return s ? 2 : 0;
  uint32_t b;
  if (s) {
    b = rows_bit_of_position(s);
    if (b >= 64) return 2;               // an odd number (>= 3) of flips pointing at nothing
  } else {
    b = (w1 >> 31) ? 63u : 62u;          // the overall parity bit itself, or the spare bit
  }
  if (b < 32) w0 ^= 1u << b;
  else w1 ^= 1u << (b - 32);
  *bit = b;
  return 1;
}

}  // namespace abft
