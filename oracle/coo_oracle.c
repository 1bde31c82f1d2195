/*
 * coo_oracle.c -- CPU ORACLE for the COO format (test infrastructure, see coo_oracle.h).
 *
 * Plain-C restatement of the reference's COO hot path (COO/CPUContext.cpp, COO/ecc.h).
 * Citations are `file:line` relative to the reference root.  Nothing is copied: the
 * Hamming masks are derived from the position rule (and checked against the reference's
 * 28 constants by tests/test_oracle_coo.py), the element is handled as four explicit
 * 32-bit words, faults come back as events instead of printf()/exit().
 *
 * Word order of the 128-bit element (COO/ecc.h:7-16, struct {uint32_t col; uint32_t row;
 * double value}): word0 = col (check byte in bits 24..31), word1 = row,
 * word2 = value bits 0..31, word3 = value bits 32..63.
 *
 * Note the orientation (COO/CPUContext.cpp:119): result[col] += value * vec[row] -- the
 * element's COLUMN is the output index.  The matrices of cg.cpp are symmetric, so this is
 * the same product as the CSR target's, accumulated per output entry in element order.
 */
#include "coo_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { MODE_NONE = 0, MODE_CONSTRAINTS = 1, MODE_SED = 2, MODE_SEC7 = 3, MODE_SEC8 = 4, MODE_SECDED = 5 };
enum {
  EV_SED_DETECTED = 1,
  EV_CORRECTED_BIT = 2,
  EV_CORRECTED_PARITY = 3,
  EV_DOUBLE_BIT = 4,
  EV_ROW_SIZE = 5,
  EV_ROW_ORDER = 6,
  EV_COL_SIZE = 7,
  EV_COL_ORDER = 8
};

/* ------------------------------------------------------------------ */
/* ECC primitives                                                      */
/* ------------------------------------------------------------------ */
static uint32_t g_mask[8][4]; /* g_mask[p][word], p = 1..7 */
static int g_ready = 0;

static int is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

/* The layout behind the 28 constants of COO/ecc.h:18-51 (their generator is the commented-out
 * gen_ecc7_masks, :136-169): element bits 24..31 are the check byte -- parity p (1..7) is stored
 * in bit 32-p, bit 24 is the overall parity of SEC8/SECDED -- and the other 120 bits, in element
 * order (col 0..23, row, value), take the non-power-of-two code positions 3,5,6,7,9,...;
 * parity p covers every position whose index has bit p-1 set, and its own storage bit. */
static void build_masks(void) {
  memset(g_mask, 0, sizeof g_mask);
  uint32_t pos = 3;
  for (int ebit = 0; ebit < 128; ebit++) {
    if (ebit >= 24 && ebit < 32) continue; /* the check byte holds no data */
    while (is_pow2(pos)) pos++;
    for (int p = 1; p <= 7; p++)
      if (pos & (1u << (p - 1))) g_mask[p][ebit / 32] |= 1u << (ebit % 32);
    pos++;
  }
  for (int p = 1; p <= 7; p++) g_mask[p][0] |= 1u << (32 - p);
  g_ready = 1;
}

uint32_t coo_hamming_mask(int p, int word) {
  if (!g_ready) build_masks();
  return g_mask[p][word];
}

static void split(uint32_t col, uint32_t row, double value, uint32_t w[4]) {
  uint64_t bits;
  memcpy(&bits, &value, 8);
  w[0] = col;
  w[1] = row;
  w[2] = (uint32_t)bits;
  w[3] = (uint32_t)(bits >> 32);
}

static double join_value(const uint32_t w[4]) {
  uint64_t bits = ((uint64_t)w[3] << 32) | w[2];
  double v;
  memcpy(&v, &bits, 8);
  return v;
}

static uint32_t parity32(uint32_t x) { return (uint32_t)__builtin_parity(x); }

/* COO/ecc.h:63-100 -- seven masked parities, parity p lands in bit 32-p */
static uint32_t syndrome_words(const uint32_t w[4]) {
  if (!g_ready) build_masks();
  uint32_t s = 0;
  for (int p = 1; p <= 7; p++) {
    uint32_t t = (w[0] & g_mask[p][0]) ^ (w[1] & g_mask[p][1]) ^ (w[2] & g_mask[p][2]) ^ (w[3] & g_mask[p][3]);
    s |= parity32(t) << (32 - p);
  }
  return s;
}

/* COO/ecc.h:108-112 */
static uint32_t overall_words(const uint32_t w[4]) { return parity32(w[0] ^ w[1] ^ w[2] ^ w[3]); }

uint32_t coo_ecc_compute_col8(uint32_t col, uint32_t row, double value) {
  uint32_t w[4];
  split(col, row, value, w);
  return syndrome_words(w);
}

uint32_t coo_ecc_overall_parity(uint32_t col, uint32_t row, double value) {
  uint32_t w[4];
  split(col, row, value, w);
  return overall_words(w);
}

/* COO/ecc.h:116-134 -- syndrome -> element bit.  The syndrome read as a number is the code
 * position.  A power of two is parity bit p itself, stored in bit 32-p = clz(position); any
 * other position is a data bit whose rank among the non-powers-of-two is
 * position - bitlength(position) - 1, and ranks from 24 on sit behind the check byte (+8).
 * Positions past the element (multi-bit faults only) continue the same arithmetic. */
uint32_t coo_ecc_get_flipped_bit_col8(uint32_t syndrome) {
  uint32_t position = 0;
  for (int p = 1; p <= 7; p++)
    if (syndrome & (1u << (32 - p))) position |= 1u << (p - 1);
  if (position == 0) return 0xFFFFFFFFu; /* undefined in the reference (clz(0)) */
  uint32_t lz = (uint32_t)__builtin_clz(position);
  if (is_pow2(position)) return lz;
  uint32_t rank = position - (32 - lz) - 1;
  return rank >= 24 ? rank + 8 : rank;
}

/* generate_ecc_bits: NONE COO/CPUContext.cpp:7-9, SED :196-199, SEC7 :230-233, SEC8 :269-273,
 * SECDED :320-324 */
uint32_t coo_encode_column(int mode, uint32_t col, uint32_t row, double value) {
  uint32_t w[4];
  split(col, row, value, w);
  switch (mode) {
    case MODE_SED: w[0] |= overall_words(w) << 31; break;
    case MODE_SEC7: w[0] |= syndrome_words(w); break;
    case MODE_SEC8:
    case MODE_SECDED:
      w[0] |= syndrome_words(w);
      w[0] |= overall_words(w) << 24;
      break;
    default: break;
  }
  return w[0];
}

/* ------------------------------------------------------------------ */
/* matrix                                                              */
/* ------------------------------------------------------------------ */
/* COO/CPUContext.cpp:11-35: one element per triplet, in the caller's order */
coo_matrix *coo_create_matrix(int mode, const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                              int nnz) {
  coo_matrix *m = (coo_matrix *)calloc(1, sizeof *m);
  m->N = (uint32_t)N;
  m->nnz = (uint32_t)nnz;
  m->mode = mode;
  m->elements = (coo_elem *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(coo_elem));
  for (int i = 0; i < nnz; i++) {
    m->elements[i].col = coo_encode_column(mode, columns[i], rows[i], values[i]);
    m->elements[i].row = rows[i];
    m->elements[i].value = values[i];
  }
  return m;
}

void coo_destroy_matrix(coo_matrix *m) {
  if (!m) return;
  free(m->elements);
  free(m);
}

typedef struct {
  coo_event *log;
  int cap, n;
} evlog;

static void push(evlog *l, int kind, uint32_t index, uint32_t bit) {
  if (l->log && l->n < l->cap) {
    l->log[l->n].kind = kind;
    l->log[l->n].index = index;
    l->log[l->n].bit = bit;
  }
  if (l->n < l->cap) l->n++;
}

/* the shared "unflip" of SEC7/SEC8/SECDED (COO/CPUContext.cpp:250-256, 293-299, 345-351).  A multi-bit
 * fault can give bit >= 128, where the reference writes past its element (undefined behaviour): the
 * oracle reports the same bit and leaves the element alone. */
static void correct_bit(uint32_t w[4], uint32_t syndrome, uint32_t index, evlog *l) {
  uint32_t bit = coo_ecc_get_flipped_bit_col8(syndrome);
  if (bit < 128) w[bit / 32] ^= 1u << (bit % 32);
  push(l, EV_CORRECTED_BIT, index, bit);
}

static void store(coo_matrix *m, uint32_t i, const uint32_t w[4]) {
  m->elements[i].col = w[0];
  m->elements[i].row = w[1];
  m->elements[i].value = join_value(w);
}

/* spmv: NONE COO/CPUContext.cpp:104-121, CONSTRAINTS :142-194, SED :201-228, SEC7 :235-267,
 * SEC8 :275-318, SECDED :326-379 */
int coo_spmv(coo_matrix *m, const double *vec, double *result, coo_event *log, int cap, int *n_events) {
  evlog l = {log, cap, 0};
  int fatal = 0;
  for (uint32_t i = 0; i < m->N; i++) result[i] = 0.0;
  for (uint32_t i = 0; i < m->nnz && !fatal; i++) {
    uint32_t w[4];
    split(m->elements[i].col, m->elements[i].row, m->elements[i].value, w);
    switch (m->mode) {
      case MODE_CONSTRAINTS: /* checks in the reference's order, first violation is fatal */
        if (w[1] >= m->N) { push(&l, EV_ROW_SIZE, i, 0); fatal = 1; break; }
        if (w[0] >= m->N) { push(&l, EV_COL_SIZE, i, 0); fatal = 1; break; }
        if (i < m->nnz - 1) {
          uint32_t next_row = m->elements[i + 1].row;
          if (w[1] > next_row) { push(&l, EV_ROW_ORDER, i, 0); fatal = 1; break; }
          if (w[1] == next_row && w[0] >= m->elements[i + 1].col) { push(&l, EV_COL_ORDER, i, 0); fatal = 1; break; }
        }
        break;
      case MODE_SED:
        if (overall_words(w)) { push(&l, EV_SED_DETECTED, i, 0); fatal = 1; }
        break;
      case MODE_SEC7: {
        uint32_t syndrome = syndrome_words(w);
        if (syndrome) {
          correct_bit(w, syndrome, i, &l);
          store(m, i, w);
        }
        break;
      }
      case MODE_SEC8: /* the syndrome is only looked at when the overall parity trips */
        if (overall_words(w)) {
          uint32_t syndrome = syndrome_words(w);
          if (syndrome) correct_bit(w, syndrome, i, &l);
          else { w[0] ^= 1u << 24; push(&l, EV_CORRECTED_PARITY, i, 0); }
          store(m, i, w);
        }
        break;
      case MODE_SECDED: {
        uint32_t overall = overall_words(w);
        uint32_t syndrome = syndrome_words(w);
        if (overall) {
          if (syndrome) correct_bit(w, syndrome, i, &l);
          else { w[0] ^= 1u << 24; push(&l, EV_CORRECTED_PARITY, i, 0); }
          store(m, i, w);
        } else if (syndrome) {
          push(&l, EV_DOUBLE_BIT, i, 0);
          fatal = 1;
        }
        break;
      }
      default: break;
    }
    if (fatal) break;
    uint32_t col = (m->mode >= MODE_SED) ? (w[0] & 0x00FFFFFFu) : w[0];
    uint32_t row = w[1];
    /* the multiply-add of every mode (:119, :192, :226, :265, :316, :377): product rounded, then added.
     * Deviation: an index no check caught (mode none, a mis-corrected multi-bit fault) can point outside
     * result[] / vec[]; the reference then reads or WRITES out of bounds (undefined behaviour).  The
     * oracle drops such an element. */
    if (col < m->N && row < m->N) {
      double prod = join_value(w) * vec[row];
      result[col] = result[col] + prod;
    }
  }
  if (n_events) *n_events = l.n;
  return fatal;
}

/* COO/CPUContext.cpp:138: the element as four words, bit/32 picks the word */
void coo_flip_bit(coo_matrix *m, int index, int bit) {
  uint32_t w[4];
  split(m->elements[index].col, m->elements[index].row, m->elements[index].value, w);
  w[bit / 32] ^= 1u << (bit % 32);
  store(m, (uint32_t)index, w);
}

/* COO/CPUContext.cpp:123-140: one rand() for the element, then one per flip; VALUE = bits 64..127,
 * INDEX = bits 0..63 */
void coo_inject_bitflip(coo_matrix *m, int kind, int num_flips, int *picked) {
  int index = rand() % (int)m->nnz;
  int start = 0, end = 128;
  if (kind == 1) start = 64;
  else if (kind == 2) end = 64;
  if (picked) picked[0] = index;
  for (int i = 0; i < num_flips; i++) {
    int bit = (rand() % (end - start)) + start;
    if (picked) picked[1 + i] = bit;
    coo_flip_bit(m, index, bit);
  }
}

/* COO/CPUContext.cpp:71-102: the same loops as the CSR target's */
static double dot(const double *a, const double *b, int N) {
  double ret = 0.0;
  for (int i = 0; i < N; i++) {
    double prod = a[i] * b[i];
    ret = ret + prod;
  }
  return ret;
}

static double calc_xr(double *x, double *r, const double *p, const double *w, double alpha, int N) {
  double ret = 0.0;
  for (int i = 0; i < N; i++) {
    double ap = alpha * p[i];
    double aw = alpha * w[i];
    x[i] = x[i] + ap;
    r[i] = r[i] - aw;
    double sq = r[i] * r[i];
    ret = ret + sq;
  }
  return ret;
}

static void calc_p(double *p, const double *r, double beta, int N) {
  for (int i = 0; i < N; i++) {
    double bp = beta * p[i];
    p[i] = r[i] + bp;
  }
}

/* cg.cpp:83-118 */
int coo_cg_solve(coo_matrix *A, const double *b, double *x, double *r, double *p, double *w, int max_itrs,
                 double conv_threshold, double *rr_hist, coo_event *log, int cap, int *n_events) {
  int N = (int)A->N;
  int total = 0;
  memcpy(r, b, sizeof(double) * (size_t)N);
  memcpy(p, r, sizeof(double) * (size_t)N);
  double rr = dot(r, r, N);
  int itr = 0;
  for (; itr < max_itrs && rr > conv_threshold; itr++) {
    int n = 0;
    int fatal = coo_spmv(A, p, w, log ? log + total : NULL, cap - total, &n);
    total += n;
    if (fatal) {
      if (n_events) *n_events = total;
      return -1 - itr;
    }
    double pw = dot(p, w, N);
    double alpha = rr / pw;
    double rr_new = calc_xr(x, r, p, w, alpha, N);
    double beta = rr_new / rr;
    calc_p(p, r, beta, N);
    rr = rr_new;
    if (rr_hist) rr_hist[itr] = rr;
  }
  if (n_events) *n_events = total;
  return itr;
}

int coo_format_event(const coo_event *e, char *buf, int buflen) {
  switch (e->kind) {
    case EV_SED_DETECTED: return snprintf(buf, buflen, "[ECC] error detected at index %d", (int)e->index);
    case EV_CORRECTED_BIT: return snprintf(buf, buflen, "[ECC] corrected bit %u at index %d", e->bit, (int)e->index);
    case EV_CORRECTED_PARITY: return snprintf(buf, buflen, "[ECC] corrected overall parity bit at index %d", (int)e->index);
    case EV_DOUBLE_BIT: return snprintf(buf, buflen, "[ECC] double-bit error detected");
    case EV_ROW_SIZE: return snprintf(buf, buflen, "row size constraint violated for index %d", (int)e->index);
    case EV_ROW_ORDER: return snprintf(buf, buflen, "row index order violated at index %d", (int)e->index);
    case EV_COL_SIZE: return snprintf(buf, buflen, "column size constraint violated for index %d", (int)e->index);
    case EV_COL_ORDER: return snprintf(buf, buflen, "column index order violated at index %d", (int)e->index);
    default: return snprintf(buf, buflen, "unknown event %d", e->kind);
  }
}
