// ecc_emul.cpp -- the ECC arithmetic of csrc/ecc.cuh compiled for the host (test infrastructure; built by
// tests/test_ecc_variants.py with g++ and the shim in tests/helpers/cuda_shim/).  The product's own source, unmodified.
#include "cuda_runtime.h"  // the shim (found first on the include path)

#include "../../abft-sparse-cg_b200/csrc/ecc.cuh"

using namespace abft;

extern "C" {

// bit p-1: Hamming parity p as the kernels compute it today (masked<P> over the three words); bit 7: overall parity
uint32_t ecc_parities_masks(uint32_t w0, uint32_t w1, uint32_t col) {
  const uint32_t s = syndrome7(w0, w1, col);  // parity p at bit 32-p
  uint32_t out = 0;
  for (int p = 1; p <= 7; p++) out |= ((s >> (32 - p)) & 1u) << (p - 1);
  return out | (overall_parity(w0, w1, col) << 7);
}

uint32_t ecc_syndrome7(uint32_t w0, uint32_t w1, uint32_t col) { return syndrome7(w0, w1, col); }
uint32_t ecc_syndrome7_any(uint32_t w0, uint32_t w1, uint32_t col) { return syndrome7_any(w0, w1, col); }
uint32_t ecc_flipped_bit(uint32_t syndrome) { return flipped_bit(syndrome); }
uint32_t ecc_encode_column(int mode, uint32_t w0, uint32_t w1, uint32_t col) { return encode_column(mode, w0, w1, col); }
uint32_t ecc_hamming_mask(int p, int word) { return hamming_mask(p, word); }

// CRC-32C from the very table constructors the library fills its device tables with (abft_cuda.cu, ensure_tables):
// byte-wise with the slice-0 table (create_matrix's tables) ...
uint32_t ecc_crc32c_bytes(const uint8_t *data, int n) {
  uint32_t c = 0xFFFFFFFFu;
  for (int i = 0; i < n; i++) c = (c >> 8) ^ crc32c_table_entry((c ^ data[i]) & 0xFFu, 0);
  return c ^ 0xFFFFFFFFu;
}
// ... slicing-by-4 with tables 0..3, one little-endian word at a time (crc_word of spmv.cuh) ...
uint32_t ecc_crc32c_words_by4(const uint32_t *w, int n_words) {
  uint32_t c = 0xFFFFFFFFu;
  for (int i = 0; i < n_words; i++) {
    const uint32_t x = c ^ w[i];
    c = crc32c_table_entry(x & 0xFF, 3) ^ crc32c_table_entry((x >> 8) & 0xFF, 2) ^ crc32c_table_entry((x >> 16) & 0xFF, 1) ^
        crc32c_table_entry(x >> 24, 0);
  }
  return c ^ 0xFFFFFFFFu;
}
// ... and three words at a time with the 5-bit chunk tables of F, F.F, F.F.F (crc_element of spmv.cuh: the spmv check)
static uint32_t map5(int m, uint32_t x) {
  uint32_t r = 0;
  for (int j = 0; j < kCrcChunks; j++) r ^= crc5_entry(m, j, (x >> (5 * j)) & 31u);
  return r;
}
uint32_t ecc_crc32c_words_by_maps(const uint32_t *w, int n_words) {   // n_words a multiple of 3
  uint32_t c = 0xFFFFFFFFu;
  for (int i = 0; i + 2 < n_words; i += 3) c = map5(3, c ^ w[i]) ^ map5(2, w[i + 1]) ^ map5(1, w[i + 2]);
  return c ^ 0xFFFFFFFFu;
}
// ... and with the 18 chunk tables cut from the whole 88-bit element {state ^ w0, w1, col} (crc_elem_chain / crc_elem_free
// of spmv.cuh: the stencil kernels' check); every third word must fit 24 bits
uint32_t ecc_crc32c_words_by_elem_chunks(const uint32_t *w, int n_words) {
  uint32_t c = 0xFFFFFFFFu;
  for (int i = 0; i + 2 < n_words; i += 3) {
    const uint32_t x = c ^ w[i], w1 = w[i + 1], col = w[i + 2];
    uint32_t r = 0;
    for (int ch = 0; ch < kCrcElemChunks; ch++) {
      const int p = 5 * ch;   // first bit of the chunk in the 88-bit string
      const unsigned __int128 e = (unsigned __int128)x | ((unsigned __int128)w1 << 32) | ((unsigned __int128)col << 64);
      r ^= crc_elem_entry(ch, (uint32_t)(e >> p) & 31u);
    }
    c = r;
  }
  return c ^ 0xFFFFFFFFu;
}

}  // extern "C"
