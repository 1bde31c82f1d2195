// ecc.cuh -- device-side ECC arithmetic for the 96-bit CSR element
// {double value; uint32_t column} (reference: CSR/ecc.h:7-11).
//
// Element bit numbering (little-endian packed struct): bits 0..31 = word0 =
// low half of the value, 32..63 = word1 = high half, 64..95 = word2 = column.
// The check bits live in column bits 24..31 (element bits 88..95).
//
// Hamming layout behind the reference's 21 mask constants (CSR/ecc.h:13-39):
// the 88 data bits (value 0..63, column 0..23) occupy, in order, the
// non-power-of-two code positions 3,5,6,7,9,...,95; parity p (1..7) sits at
// position 2^(p-1), is stored in column bit 32-p, and covers every position
// whose index has bit p-1 set.  The masks are derived from that rule at
// compile time; two of the reference's constants are static_assert'ed as
// known answers.
#pragma once
#include <stdint.h>

namespace abft {

constexpr bool ce_is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

// mask of parity p (1..7) over word (0..2), including the parity bit itself
constexpr uint32_t hamming_mask(int p, int word) {
  uint32_t m = 0;
  uint32_t pos = 3;
  for (int d = 0; d < 88; d++) {
    while (ce_is_pow2(pos)) pos++;
    if ((pos >> (p - 1)) & 1u)
      if (d / 32 == word) m |= 1u << (d % 32);
    pos++;
  }
  const int self = 96 - p;
  if (self / 32 == word) m |= 1u << (self % 32);
  return m;
}

static_assert(hamming_mask(1, 0) == 0x56AAAD5Bu, "P1 word0 (CSR/ecc.h:13)");
static_assert(hamming_mask(7, 2) == 0x02FFFFFFu, "P7 word2 (CSR/ecc.h:39)");

template <int P>
struct HM {
  static constexpr uint32_t w0 = hamming_mask(P, 0);
  static constexpr uint32_t w1 = hamming_mask(P, 1);
  static constexpr uint32_t w2 = hamming_mask(P, 2);
};

template <int P>
__device__ __forceinline__ uint32_t masked(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t t = 0;
  if (HM<P>::w0) t ^= a & HM<P>::w0;
  if (HM<P>::w1) t ^= b & HM<P>::w1;
  if (HM<P>::w2) t ^= c & HM<P>::w2;
  return t;
}

// ecc_compute_overall_parity (CSR/ecc.h:89-93)
__device__ __forceinline__ uint32_t overall_parity(uint32_t w0, uint32_t w1, uint32_t w2) {
  return __popc(w0 ^ w1 ^ w2) & 1u;
}

// ecc_compute_col8 (CSR/ecc.h:51-81): parity p -> bit 32-p
__device__ __forceinline__ uint32_t syndrome7(uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t s = 0;
  s |= (__popc(masked<1>(w0, w1, w2)) & 1u) << 31;
  s |= (__popc(masked<2>(w0, w1, w2)) & 1u) << 30;
  s |= (__popc(masked<3>(w0, w1, w2)) & 1u) << 29;
  s |= (__popc(masked<4>(w0, w1, w2)) & 1u) << 28;
  s |= (__popc(masked<5>(w0, w1, w2)) & 1u) << 27;
  s |= (__popc(masked<6>(w0, w1, w2)) & 1u) << 26;
  s |= (__popc(masked<7>(w0, w1, w2)) & 1u) << 25;
  return s;
}

// Fast-path predicate: bit 0 set iff any of the seven Hamming parities is odd.
// Same arithmetic as syndrome7 != 0, without assembling the syndrome.
__device__ __forceinline__ uint32_t syndrome7_any(uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t a = __popc(masked<1>(w0, w1, w2)) | __popc(masked<2>(w0, w1, w2)) | __popc(masked<3>(w0, w1, w2));
  uint32_t b = __popc(masked<4>(w0, w1, w2)) | __popc(masked<5>(w0, w1, w2)) | __popc(masked<6>(w0, w1, w2));
  return (a | b | __popc(masked<7>(w0, w1, w2))) & 1u;
}

// ecc_get_flipped_bit_col8 (CSR/ecc.h:97-113).  syndrome != 0.
__device__ __forceinline__ uint32_t flipped_bit(uint32_t syndrome) {
  const uint32_t position = __brev(syndrome) & 0x7Fu;  // syndrome bit 32-p == position bit p-1
  const uint32_t lz = __clz(position);
  if ((position & (position - 1)) == 0) return lz + 64;  // a parity bit itself: element bit 96-p
  return position - (32 - lz) - 1;                       // rank among the non-powers-of-two
}

// generate_ecc_bits of the per-element modes (CSR/CPUContext.cpp:209-212,
// 247-250, 291-295, 347-351).  MODE values follow include/abft_cuda.h.
__device__ __forceinline__ uint32_t encode_column(int mode, uint32_t w0, uint32_t w1, uint32_t col) {
  if (mode == 2) {  // SED: parity in bit 31
    col |= overall_parity(w0, w1, col) << 31;
  } else if (mode == 3) {  // SEC7
    col |= syndrome7(w0, w1, col);
  } else if (mode == 4 || mode == 5) {  // SEC8 / SECDED: Hamming bits, then overall parity in bit 24
    col |= syndrome7(w0, w1, col);
    col |= overall_parity(w0, w1, col) << 24;
  }
  return col;
}

// ---------------------------------------------------------------------------
// Vector protection (mode secdedv; repository-defined): SED over every double of x, r, p, w -- the
// mantissa's least significant bit is given up for the parity of the other 63 bits, so a stored element always has
// even parity and a single flipped bit anywhere in it is detected when the element is next read.  The reference
// protects no vector (CSR/CPUContext.cpp:54-113); this costs one ulp of precision per stored value.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double vec_encode(double v) {
  const uint32_t lo = (uint32_t)__double2loint(v) & ~1u, hi = (uint32_t)__double2hiint(v);
  return __hiloint2double((int)hi, (int)(lo | ((uint32_t)__popc(lo ^ hi) & 1u)));
}
// bit 0 set iff the element's parity is odd (corrupted); callers OR the raw popcounts and test bit 0 once
__device__ __forceinline__ uint32_t vec_parity_bits(double v) {
  return (uint32_t)__popc((uint32_t)__double2loint(v) ^ (uint32_t)__double2hiint(v));
}
// which vector a detected fault was found in (the `bit` field of ABFT_EV_VECTOR_SED)
enum { VEC_SPMV_OPERAND = 1, VEC_X = 2, VEC_R = 3, VEC_P = 4, VEC_W = 5, VEC_DOT_A = 6, VEC_DOT_B = 7, VEC_MAPPED = 8 };

// ---------------------------------------------------------------------------
// CRC-32C (Castagnoli, reflected 0x82F63B78), repo-defined row code
// ---------------------------------------------------------------------------
constexpr uint32_t crc32c_table_entry(uint32_t i, int slice) {
  uint32_t c = i;
  for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
  // slice s: the effect of byte i followed by s zero bytes
  for (int s = 0; s < slice; s++) {
    uint32_t lo = c & 0xFFu;
    uint32_t t = lo;
    for (int k = 0; k < 8; k++) t = (t >> 1) ^ (0x82F63B78u & (0u - (t & 1u)));
    c = (c >> 8) ^ t;
  }
  return c;
}

// The spmv check's tables (spmv.cuh, "CRC32C row check"): F = "advance the raw state over four zero bytes" is linear,
// so F^m(x) is the XOR of one table entry per 5-bit chunk of x.  Entry c of the table of chunk j of map F^m (m = 1..3):
constexpr int kCrcMaps = 3, kCrcChunks = 7;
constexpr int kCrcTabWords = kCrcMaps * kCrcChunks * 32;
constexpr uint32_t crc5_entry(int m, int j, uint32_t c) {   // (host side: library init)
  uint32_t x = (uint32_t)((uint64_t)c << (5 * j));
  for (int r = 0; r < 4 * m; r++) x = (x >> 8) ^ crc32c_table_entry(x & 0xFFu, 0);
  return x;
}

// The same check with the chunks cut from the whole 88-bit element {state ^ w0, w1, col} instead of from each word:
// 18 chunks instead of 7 + 7 + 5 (chunk 6 takes two bits of the first word and three of the second, chunk 12 four of
// the second and one of the column).  Bit p of the element goes through F^3 (p < 32), F^2 (p < 64) or F (p < 88).
constexpr int kCrcElemChunks = 18;
constexpr uint32_t crc_elem_entry(int chunk, uint32_t c) {   // (host side: library init)
  uint32_t r = 0;
  for (int b = 0; b < 5; b++) {
    const int p = 5 * chunk + b;
    if (!((c >> b) & 1u) || p >= 88) continue;
    const int m = p < 32 ? 3 : (p < 64 ? 2 : 1), bit = p & 31;   // (p - 64 == p & 31 for 64 <= p < 88)
    r ^= crc5_entry(m, bit / 5, 1u << (bit % 5));
  }
  return r;
}
}  // namespace abft
