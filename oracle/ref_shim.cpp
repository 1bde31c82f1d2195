// ref_shim.cpp -- ORACLE SUPPORT (test infrastructure, see abft_oracle.h).
//
// A C-linkage window onto the UNMODIFIED reference: it includes the
// reference's own headers (found through -I$REF -I$REF/CSR at build time; no
// reference source lives in this repository) and forwards to the reference's
// CGContext plugin registry ("cpu", <mode>) and to the inline functions of its
// CSR/ecc.h.  oracle/Makefile links it with the reference's CGContext.cpp and
// CSR/CPUContext.cpp into oracle/_ref/libref_csr.so.
//
// Used to (1) pin oracle/abft_oracle.c against the real thing, (2) generate
// the fixtures in tests/golden/, (3) time the reference CPU target for
// bench.py's cpu_baseline / --impl reference legs.
//
// The reference prints fault messages with printf() and exit(1)s on fatal
// ones; callers capture stdout around a call and must not drive a fatal case
// through this in-process shim (use the oracle/_ref/cg-csr binary for those).
#include "CPUContext.h"  // reference: CSR/CPUContext.h (pulls CGContext.h, ecc.h)

#include <cstdio>
#include <cstring>

extern "C" {

// ---- CSR/ecc.h ----
uint32_t ref_hamming_mask(int p, int word) {
  static const uint32_t M[8][3] = {
      {0, 0, 0},
      {ECC7_P1_0, ECC7_P1_1, ECC7_P1_2},
      {ECC7_P2_0, ECC7_P2_1, ECC7_P2_2},
      {ECC7_P3_0, ECC7_P3_1, ECC7_P3_2},
      {ECC7_P4_0, ECC7_P4_1, ECC7_P4_2},
      {ECC7_P5_0, ECC7_P5_1, ECC7_P5_2},
      {ECC7_P6_0, ECC7_P6_1, ECC7_P6_2},
      {ECC7_P7_0, ECC7_P7_1, ECC7_P7_2},
  };
  return M[p][word];
}

static csr_element make(double value, uint32_t column) {
  csr_element e;
  e.value = value;
  e.column = column;
  return e;
}

uint32_t ref_ecc_compute_col8(double value, uint32_t column) { return ecc_compute_col8(make(value, column)); }
uint32_t ref_ecc_overall_parity(double value, uint32_t column) { return ecc_compute_overall_parity(make(value, column)); }
uint32_t ref_ecc_get_flipped_bit_col8(uint32_t syndrome) { return ecc_get_flipped_bit_col8(syndrome); }

// ---- CGContext plugin API ----
void *ref_context_create(const char *mode) { return CGContext::create("cpu", mode); }
void ref_context_destroy(void *ctx) { delete (CGContext *)ctx; }

void *ref_create_matrix(void *ctx, const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                        int nnz) {
  return ((CGContext *)ctx)->create_matrix(columns, rows, values, N, nnz);
}
void ref_destroy_matrix(void *ctx, void *mat) { ((CGContext *)ctx)->destroy_matrix((cg_matrix *)mat); }

// raw views of the reference's storage (CSR/CPUContext.h:5-18)
uint32_t *ref_matrix_cols(void *mat) { return ((cg_matrix *)mat)->cols; }
uint32_t *ref_matrix_rows(void *mat) { return ((cg_matrix *)mat)->rows; }
double *ref_matrix_values(void *mat) { return ((cg_matrix *)mat)->values; }

void *ref_create_vector(void *ctx, int N) { return ((CGContext *)ctx)->create_vector(N); }
void ref_destroy_vector(void *ctx, void *v) { ((CGContext *)ctx)->destroy_vector((cg_vector *)v); }
double *ref_map_vector(void *ctx, void *v) { return ((CGContext *)ctx)->map_vector((cg_vector *)v); }
void ref_unmap_vector(void *ctx, void *v, double *h) { ((CGContext *)ctx)->unmap_vector((cg_vector *)v, h); }
void ref_copy_vector(void *ctx, void *dst, void *src) {
  ((CGContext *)ctx)->copy_vector((cg_vector *)dst, (cg_vector *)src);
}
double ref_dot(void *ctx, void *a, void *b) { return ((CGContext *)ctx)->dot((cg_vector *)a, (cg_vector *)b); }
double ref_calc_xr(void *ctx, void *x, void *r, void *p, void *w, double alpha) {
  return ((CGContext *)ctx)->calc_xr((cg_vector *)x, (cg_vector *)r, (cg_vector *)p, (cg_vector *)w, alpha);
}
void ref_calc_p(void *ctx, void *p, void *r, double beta) {
  ((CGContext *)ctx)->calc_p((cg_vector *)p, (cg_vector *)r, beta);
}
void ref_spmv(void *ctx, void *mat, void *vec, void *result) {
  ((CGContext *)ctx)->spmv((cg_matrix *)mat, (cg_vector *)vec, (cg_vector *)result);
  fflush(stdout);
}
// kind: 0 ANY, 1 VALUE, 2 INDEX (CGContext.h:11); draws from rand() like the CLI does
void ref_inject_bitflip(void *ctx, void *mat, int kind, int num_flips) {
  ((CGContext *)ctx)->inject_bitflip((cg_matrix *)mat, (CGContext::BitFlipKind)kind, num_flips);
  fflush(stdout);
}

// K iterations of the loop body of cg.cpp:94-118 driven through the plugin API
// (no printf per iteration); returns iterations run, rr via *rr_io.
int ref_cg_iterations(void *ctx, void *A, void *x, void *r, void *p, void *w, int max_itrs, double conv_threshold,
                      double *rr_io, double *rr_hist) {
  CGContext *c = (CGContext *)ctx;
  double rr = *rr_io;
  int itr = 0;
  for (; itr < max_itrs && rr > conv_threshold; itr++) {
    c->spmv((cg_matrix *)A, (cg_vector *)p, (cg_vector *)w);
    double pw = c->dot((cg_vector *)p, (cg_vector *)w);
    double alpha = rr / pw;
    double rr_new = c->calc_xr((cg_vector *)x, (cg_vector *)r, (cg_vector *)p, (cg_vector *)w, alpha);
    double beta = rr_new / rr;
    c->calc_p((cg_vector *)p, (cg_vector *)r, beta);
    rr = rr_new;
    if (rr_hist) rr_hist[itr] = rr;
  }
  fflush(stdout);
  *rr_io = rr;
  return itr;
}

}  // extern "C"
