// ref_shim_coo.cpp -- ORACLE SUPPORT (test infrastructure, see coo_oracle.h).
//
// A C-linkage window onto the UNMODIFIED reference COO target: it includes the reference's own
// headers (found through -I$REF -I$REF/COO at build time; no reference source lives in this
// repository) and forwards to the reference's CGContext plugin registry ("cpu", <mode>) and to the
// inline functions of its COO/ecc.h.  oracle/Makefile links it with the reference's CGContext.cpp
// and COO/CPUContext.cpp into oracle/_ref/libref_coo.so.
//
// Used to pin oracle/coo_oracle.c against the real thing and to generate tests/golden/coo_*.json.
// The reference prints fault messages with printf() and exit(1)s on fatal ones: callers capture
// stdout around a call and must not drive a fatal case through this in-process shim (use the
// oracle/_ref/cg-coo binary for those).
#include "CPUContext.h"  // reference: COO/CPUContext.h (pulls CGContext.h, COO/ecc.h)

#include <cstdio>
#include <cstring>

extern "C" {

// ---- COO/ecc.h ----
uint32_t refcoo_hamming_mask(int p, int word) {
  static const uint32_t M[8][4] = {
      {0, 0, 0, 0},
      {ECC7_P1_0, ECC7_P1_1, ECC7_P1_2, ECC7_P1_3},
      {ECC7_P2_0, ECC7_P2_1, ECC7_P2_2, ECC7_P2_3},
      {ECC7_P3_0, ECC7_P3_1, ECC7_P3_2, ECC7_P3_3},
      {ECC7_P4_0, ECC7_P4_1, ECC7_P4_2, ECC7_P4_3},
      {ECC7_P5_0, ECC7_P5_1, ECC7_P5_2, ECC7_P5_3},
      {ECC7_P6_0, ECC7_P6_1, ECC7_P6_2, ECC7_P6_3},
      {ECC7_P7_0, ECC7_P7_1, ECC7_P7_2, ECC7_P7_3},
  };
  return M[p][word];
}

static coo_element make(uint32_t col, uint32_t row, double value) {
  coo_element e;
  e.col = col;
  e.row = row;
  e.value = value;
  return e;
}

uint32_t refcoo_ecc_compute_col8(uint32_t col, uint32_t row, double value) { return ecc_compute_col8(make(col, row, value)); }
uint32_t refcoo_ecc_overall_parity(uint32_t col, uint32_t row, double value) {
  return ecc_compute_overall_parity(make(col, row, value));
}
uint32_t refcoo_ecc_get_flipped_bit_col8(uint32_t syndrome) { return ecc_get_flipped_bit_col8(syndrome); }

// ---- CGContext plugin API ----
void *refcoo_context_create(const char *mode) { return CGContext::create("cpu", mode); }
void refcoo_context_destroy(void *ctx) { delete (CGContext *)ctx; }

void *refcoo_create_matrix(void *ctx, const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                           int nnz) {
  return ((CGContext *)ctx)->create_matrix(columns, rows, values, N, nnz);
}
void refcoo_destroy_matrix(void *ctx, void *mat) { ((CGContext *)ctx)->destroy_matrix((cg_matrix *)mat); }

// raw view of the reference's storage (COO/CPUContext.h: struct cg_matrix {N, nnz, elements}); 16 bytes per element
void *refcoo_matrix_elements(void *mat) { return ((cg_matrix *)mat)->elements; }

void *refcoo_create_vector(void *ctx, int N) { return ((CGContext *)ctx)->create_vector(N); }
void refcoo_destroy_vector(void *ctx, void *v) { ((CGContext *)ctx)->destroy_vector((cg_vector *)v); }
double *refcoo_map_vector(void *ctx, void *v) { return ((CGContext *)ctx)->map_vector((cg_vector *)v); }
void refcoo_copy_vector(void *ctx, void *dst, void *src) {
  ((CGContext *)ctx)->copy_vector((cg_vector *)dst, (cg_vector *)src);
}
double refcoo_dot(void *ctx, void *a, void *b) { return ((CGContext *)ctx)->dot((cg_vector *)a, (cg_vector *)b); }
void refcoo_spmv(void *ctx, void *mat, void *vec, void *result) {
  ((CGContext *)ctx)->spmv((cg_matrix *)mat, (cg_vector *)vec, (cg_vector *)result);
  fflush(stdout);
}
// kind: 0 ANY, 1 VALUE, 2 INDEX (CGContext.h:11); draws from rand() like the CLI does
void refcoo_inject_bitflip(void *ctx, void *mat, int kind, int num_flips) {
  ((CGContext *)ctx)->inject_bitflip((cg_matrix *)mat, (CGContext::BitFlipKind)kind, num_flips);
  fflush(stdout);
}

// K iterations of the loop body of cg.cpp:94-118 driven through the plugin API
int refcoo_cg_iterations(void *ctx, void *A, void *x, void *r, void *p, void *w, int max_itrs, double conv_threshold,
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
