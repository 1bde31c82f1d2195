// Host program (built and run by tests/test_abi.py, no GPU needed): prints the Hamming masks the CUDA kernels
// derive at compile time, so that they can be compared with the reference's constants (tests/golden/).
#include <cstdio>

#include "../../abft-sparse-cg_b200/csrc/ecc.cuh"
#include "../../abft-sparse-cg_b200/csrc/coo.cuh"

int main() {
  for (int p = 1; p <= 7; p++) printf("csr %d %u %u %u\n", p, abft::hamming_mask(p, 0), abft::hamming_mask(p, 1), abft::hamming_mask(p, 2));
  for (int p = 1; p <= 7; p++)
    printf("coo %d %u %u %u %u\n", p, abft::coo_mask(p, 0), abft::coo_mask(p, 1), abft::coo_mask(p, 2), abft::coo_mask(p, 3));
  return 0;
}
