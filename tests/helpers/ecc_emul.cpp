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

}  // extern "C"
