/*
 * coo_oracle.h -- CPU ORACLE for the COO CG hot path of abft-sparse-cg (the
 * reference's second executable, cg-coo).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as abft_oracle.h:
 * only tests/, bench.py's CPU legs and __graft_entry__.smoke() may load it).
 *
 * Plain-C restatement of COO/CPUContext.cpp and COO/ecc.h; every function
 * cites the reference file:line it follows.  Parity status: PINNED against
 * the reference itself (oracle/_ref/libref_coo.so, built from the reference's
 * own sources by oracle/Makefile) and against tests/golden/coo_*.json.
 *
 * Status of the CUDA side: the COO backend is the next row of the scope table
 * (DESIGN.md 0, row f); this oracle and its fixtures are the parity gate it
 * will be built against.
 */
#ifndef COO_ORACLE_H
#define COO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* modes: numbering shared with abft_oracle.h / include/abft_cuda.h (0 none .. 5 secded) */

/* events: kinds shared with abft_oracle.h; the COO CONSTRAINTS messages differ in wording
 * (COO/CPUContext.cpp:155-187) and all carry the ELEMENT index:
 *   5 "row size constraint violated for index %d"     6 "row index order violated at index %d"
 *   7 "column size constraint violated for index %d"  8 "column index order violated at index %d" */
typedef struct {
  int32_t kind;
  uint32_t index; /* COO element index */
  uint32_t bit;   /* flipped bit (0..127) for CORRECTED_BIT, else 0 */
} coo_event;

/* 128-bit element, COO/ecc.h:7-16: word0 = col (check bits in 24..31), word1 = row,
 * word2 = value bits 0..31, word3 = value bits 32..63 */
typedef struct {
  uint32_t col;
  uint32_t row;
  double value;
} coo_elem;

typedef struct {
  uint32_t N;
  uint32_t nnz;
  int32_t mode;
  coo_elem *elements;
} coo_matrix;

/* ---- ECC primitives (COO/ecc.h) ---- */
uint32_t coo_hamming_mask(int p /*1..7*/, int word /*0..3*/);
uint32_t coo_ecc_compute_col8(uint32_t col, uint32_t row, double value);
uint32_t coo_ecc_overall_parity(uint32_t col, uint32_t row, double value);
uint32_t coo_ecc_get_flipped_bit_col8(uint32_t syndrome);
uint32_t coo_encode_column(int mode, uint32_t col, uint32_t row, double value);

/* ---- matrix / kernels (COO/CPUContext.cpp) ---- */
coo_matrix *coo_create_matrix(int mode, const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                              int nnz);
void coo_destroy_matrix(coo_matrix *m);
/* 0 = completed, 1 = fatal event hit (result[] then holds the accumulation up to the fault) */
int coo_spmv(coo_matrix *m, const double *vec, double *result, coo_event *log, int cap, int *n_events);
void coo_flip_bit(coo_matrix *m, int index, int bit /*0..127*/);
/* the reference's rand()-driven injection; kind 0=ANY 1=VALUE 2=INDEX; picked[0] = index, picked[1+i] = bit */
void coo_inject_bitflip(coo_matrix *m, int kind, int num_flips, int *picked);
int coo_cg_solve(coo_matrix *A, const double *b, double *x, double *r, double *p, double *w, int max_itrs,
                 double conv_threshold, double *rr_hist, coo_event *log, int cap, int *n_events);
int coo_format_event(const coo_event *e, char *buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif
