/*
 * abft_oracle.h -- CPU ORACLE for the CSR CG hot path of abft-sparse-cg.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may link,
 * load or execute anything under oracle/.  The CUDA product path
 * (abft-sparse-cg_b200/) never includes or calls it.
 *
 * It is a plain-C restatement of the reference's algorithm; every function
 * cites the reference file:line it follows (paths relative to the reference
 * root).  Parity status:
 *   - modes none/constraints/sed/sec7/sec8/secded: PINNED against the
 *     reference itself (oracle/_ref, built from the reference's own sources
 *     by oracle/Makefile) and against the committed fixtures in tests/golden/.
 *   - mode crc32c: PARITY UNPINNED.  The reference snapshot has no CRC32C
 *     mode; the format is defined by this repository
 *     (DESIGN.md "CRC32C format") and this file is its only CPU statement.
 */
#ifndef ABFT_ORACLE_H
#define ABFT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABFT modes; the numbering is shared with include/abft_cuda.h */
enum {
  ORC_MODE_NONE = 0,
  ORC_MODE_CONSTRAINTS = 1,
  ORC_MODE_SED = 2,
  ORC_MODE_SEC7 = 3,
  ORC_MODE_SEC8 = 4,
  ORC_MODE_SECDED = 5,
  ORC_MODE_CRC32C = 6
};

/* Fault events; one per message the reference prints from inside spmv.
 * The numbering is shared with include/abft_cuda.h */
enum {
  ORC_EV_SED_DETECTED = 1,     /* "[ECC] error detected at index %d"        fatal */
  ORC_EV_CORRECTED_BIT = 2,    /* "[ECC] corrected bit %u at index %d"            */
  ORC_EV_CORRECTED_PARITY = 3, /* "[ECC] corrected overall parity bit at index %d"*/
  ORC_EV_DOUBLE_BIT = 4,       /* "[ECC] double-bit error detected"         fatal */
  ORC_EV_ROW_SIZE = 5,         /* "row size constraint violated for row %d" fatal */
  ORC_EV_ROW_ORDER = 6,        /* "row order constraint violated for row%d" fatal */
  ORC_EV_COL_SIZE = 7,         /* "column size constraint violated at index %d"  fatal */
  ORC_EV_COL_ORDER = 8,        /* "column order constraint violated at index %d" fatal */
  ORC_EV_CRC_ROW = 9           /* "[CRC32C] error detected in row %d"       fatal (repo-defined) */
};

typedef struct {
  int32_t kind;   /* ORC_EV_* */
  uint32_t index; /* CSR element index, or row for ROW_* / CRC_ROW */
  uint32_t bit;   /* flipped bit (0..95) for CORRECTED_BIT, else 0 */
} orc_event;

typedef struct {
  uint32_t N;
  uint32_t nnz;
  int32_t mode;
  uint32_t *cols;  /* nnz, ECC bits in bits 24..31 for the ECC modes */
  uint32_t *rows;  /* N+1 */
  double *values;  /* nnz */
} orc_matrix;

/* ---- ECC primitives (CSR/ecc.h) ---- */
uint32_t orc_hamming_mask(int p /*1..7*/, int word /*0..2*/);
uint32_t orc_ecc_compute_col8(double value, uint32_t column);
uint32_t orc_ecc_overall_parity(double value, uint32_t column);
uint32_t orc_ecc_get_flipped_bit_col8(uint32_t syndrome);
/* generate_ecc_bits for the per-element modes: returns the encoded column */
uint32_t orc_encode_column(int mode, double value, uint32_t column);

/* ---- CRC32C (repo-defined) ---- */
uint32_t orc_crc32c(const void *data, uint64_t len);
uint32_t orc_crc32c_row(const uint32_t *cols, const double *values, uint32_t n);

/* ---- matrix / kernels (CSR/CPUContext.cpp) ---- */
orc_matrix *orc_create_matrix(int mode, const uint32_t *columns,
                              const uint32_t *rows, const double *values,
                              int N, int nnz);
void orc_destroy_matrix(orc_matrix *m);

/* returns 0 = completed, 1 = a fatal event was hit (the reference exit(1)s
 * there; result[] is then only written for the rows before the fault).
 * Events are appended to log[0..cap) in serial (row, element) order;
 * *n_events gets the number appended (never more than cap). */
int orc_spmv(orc_matrix *m, const double *vec, double *result, orc_event *log,
             int cap, int *n_events);

double orc_dot(const double *a, const double *b, int N);
double orc_calc_xr(double *x, double *r, const double *p, const double *w,
                   double alpha, int N);
void orc_calc_p(double *p, const double *r, double beta, int N);

/* flip one bit (0..95) of element `index`: bits 0..63 = value, 64..95 = column */
void orc_flip_bit(orc_matrix *m, int index, int bit);
/* the reference's rand()-driven injection; kind 0=ANY 1=VALUE 2=INDEX.
 * picked[0] = index, picked[1+i] = bit of flip i (caller gives 1+num_flips ints) */
void orc_inject_bitflip(orc_matrix *m, int kind, int num_flips, int *picked);

/* The CG loop of cg.cpp:83-118.  rr_hist (may be NULL) receives rr after each
 * iteration (at most max_itrs entries).  Returns iterations run, or -1-(itr)
 * if a fatal event stopped iteration itr. */
int orc_cg_solve(orc_matrix *A, const double *b, double *x, double *r,
                 double *p, double *w, int max_itrs, double conv_threshold,
                 double *rr_hist, orc_event *log, int cap, int *n_events);

/* message text exactly as the reference prints it (no trailing newline) */
int orc_format_event(const orc_event *e, char *buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif
