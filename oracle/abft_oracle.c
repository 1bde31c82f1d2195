/*
 * abft_oracle.c -- CPU ORACLE (test infrastructure, see abft_oracle.h).
 *
 * Plain-C restatement of the reference CSR hot path.  Citations are
 * `file:line` relative to the reference root (/root/reference in the build
 * container).  Nothing here is copied from the reference: the Hamming masks
 * are *derived* from the position rule (and checked against the reference's
 * constants by tests/test_oracle_ref.py), the element is handled as three
 * explicit 32-bit words instead of a packed struct, and faults are returned
 * as events instead of printf()/exit().
 *
 * Word order of the 96-bit element (CSR/ecc.h:7-11, little-endian packed
 * struct {double value; uint32_t column}):
 *   word0 = value bits 0..31, word1 = value bits 32..63, word2 = column.
 */
#include "abft_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* ECC primitives                                                      */
/* ------------------------------------------------------------------ */

static uint32_t g_mask[8][3]; /* g_mask[p][word], p = 1..7 */
static int g_mask_ready = 0;

static int is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

/* Hamming layout behind the 21 constants of CSR/ecc.h:13-39: the 88 data
 * bits (value bits 0..63, then column bits 0..23) sit, in order, at the
 * non-power-of-two code positions 3,5,6,7,9,...,95; parity bit p (1..7)
 * sits at position 2^(p-1) and is stored in column bit 32-p (element bit
 * 96-p); parity p covers every position whose binary index has bit p-1 set. */
static void build_masks(void) {
  uint32_t pos = 3;
  memset(g_mask, 0, sizeof g_mask);
  for (int d = 0; d < 88; d++) {
    while (is_pow2(pos)) pos++;
    int ebit = d; /* element bit == data bit for d < 88 (column bits 0..23 follow the value) */
    for (int p = 1; p <= 7; p++)
      if (pos & (1u << (p - 1))) g_mask[p][ebit / 32] |= 1u << (ebit % 32);
    pos++;
  }
  for (int p = 1; p <= 7; p++) {
    int ebit = 96 - p; /* the parity bit itself */
    g_mask[p][ebit / 32] |= 1u << (ebit % 32);
  }
  g_mask_ready = 1;
}

uint32_t orc_hamming_mask(int p, int word) {
  if (!g_mask_ready) build_masks();
  return g_mask[p][word];
}

static void split(double value, uint32_t column, uint32_t w[3]) {
  uint64_t bits;
  memcpy(&bits, &value, 8);
  w[0] = (uint32_t)bits;
  w[1] = (uint32_t)(bits >> 32);
  w[2] = column;
}

static double join(const uint32_t w[3]) {
  uint64_t bits = ((uint64_t)w[1] << 32) | w[0];
  double v;
  memcpy(&v, &bits, 8);
  return v;
}

static uint32_t parity32(uint32_t x) { return (uint32_t)__builtin_parity(x); }

/* CSR/ecc.h:51-81 -- seven masked parities, parity p lands in bit 32-p */
static uint32_t syndrome_words(const uint32_t w[3]) {
  if (!g_mask_ready) build_masks();
  uint32_t s = 0;
  for (int p = 1; p <= 7; p++) {
    uint32_t t = (w[0] & g_mask[p][0]) ^ (w[1] & g_mask[p][1]) ^ (w[2] & g_mask[p][2]);
    s |= parity32(t) << (32 - p);
  }
  return s;
}

uint32_t orc_ecc_compute_col8(double value, uint32_t column) {
  uint32_t w[3];
  split(value, column, w);
  return syndrome_words(w);
}

/* CSR/ecc.h:89-93 */
static uint32_t overall_words(const uint32_t w[3]) { return parity32(w[0] ^ w[1] ^ w[2]); }

uint32_t orc_ecc_overall_parity(double value, uint32_t column) {
  uint32_t w[3];
  split(value, column, w);
  return overall_words(w);
}

/* CSR/ecc.h:97-113 -- syndrome -> element bit.  The syndrome is the Hamming
 * position (bit 32-p of the syndrome = bit p-1 of the position).  A power of
 * two is a parity bit itself (element bit 96-p == clz(position)+64); anything
 * else is the data bit whose rank among the non-powers-of-two is
 * position - bitlength(position) - 1.  Positions above 95 (only reachable by
 * multi-bit faults) continue the same arithmetic, as the reference does. */
uint32_t orc_ecc_get_flipped_bit_col8(uint32_t syndrome) {
  uint32_t position = 0;
  for (int p = 1; p <= 7; p++)
    if (syndrome & (1u << (32 - p))) position |= 1u << (p - 1);
  if (position == 0) return 0xFFFFFFFFu; /* undefined in the reference (clz(0)) */
  uint32_t lz = (uint32_t)__builtin_clz(position);
  if (is_pow2(position)) return lz + 64;
  return position - (32 - lz) - 1;
}

/* generate_ecc_bits: NONE CSR/CPUContext.cpp:7-9, SED :209-212, SEC7 :247-250,
 * SEC8 :291-295, SECDED :347-351 */
uint32_t orc_encode_column(int mode, double value, uint32_t column) {
  uint32_t w[3];
  split(value, column, w);
  switch (mode) {
    case ORC_MODE_SED:
      w[2] |= overall_words(w) << 31;
      break;
    case ORC_MODE_SEC7:
      w[2] |= syndrome_words(w);
      break;
    case ORC_MODE_SEC8:
    case ORC_MODE_SECDED:
      w[2] |= syndrome_words(w);
      w[2] |= overall_words(w) << 24; /* parity taken AFTER the Hamming bits are in */
      break;
    default:
      break;
  }
  return w[2];
}

/* ------------------------------------------------------------------ */
/* CRC32C (repo-defined mode; PARITY UNPINNED -- not in the reference)  */
/* ------------------------------------------------------------------ */

static uint32_t g_crc_table[256];
static int g_crc_ready = 0;

static void build_crc_table(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
    g_crc_table[i] = c;
  }
  g_crc_ready = 1;
}

static uint32_t crc_update(uint32_t crc, const uint8_t *p, uint64_t len) {
  if (!g_crc_ready) build_crc_table();
  for (uint64_t i = 0; i < len; i++) crc = g_crc_table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return crc;
}

/* Standard CRC-32C (Castagnoli): reflected poly 0x82F63B78, init and final
 * xor 0xFFFFFFFF; crc32c("123456789") == 0xE3069283. */
uint32_t orc_crc32c(const void *data, uint64_t len) {
  return crc_update(0xFFFFFFFFu, (const uint8_t *)data, len) ^ 0xFFFFFFFFu;
}

/* Row message: per element, 8 little-endian bytes of the value followed by 4
 * little-endian bytes of (column & 0x00FFFFFF). */
uint32_t orc_crc32c_row(const uint32_t *cols, const double *values, uint32_t n) {
  uint32_t crc = 0xFFFFFFFFu;
  for (uint32_t k = 0; k < n; k++) {
    uint8_t msg[12];
    uint32_t c = cols[k] & 0x00FFFFFFu;
    memcpy(msg, &values[k], 8);
    memcpy(msg + 8, &c, 4);
    crc = crc_update(crc, msg, 12);
  }
  return crc ^ 0xFFFFFFFFu;
}

/* Encode one row in place.  n >= 4: CRC byte k goes to bits 24..31 of the
 * k-th column word.  n < 4: there is no room for 32 CRC bits, each element
 * carries the SECDED check byte instead (checked detect-only). */
static void crc_encode_row(uint32_t *cols, const double *values, uint32_t n) {
  if (n >= 4) {
    uint32_t crc = orc_crc32c_row(cols, values, n);
    for (uint32_t k = 0; k < 4; k++) cols[k] = (cols[k] & 0x00FFFFFFu) | (((crc >> (8 * k)) & 0xFFu) << 24);
  } else {
    for (uint32_t k = 0; k < n; k++) cols[k] = orc_encode_column(ORC_MODE_SECDED, values[k], cols[k]);
  }
}

/* returns non-zero if the row fails its check */
static int crc_check_row(const uint32_t *cols, const double *values, uint32_t n) {
  if (n >= 4) {
    uint32_t stored = 0;
    for (uint32_t k = 0; k < 4; k++) stored |= (cols[k] >> 24) << (8 * k);
    return stored != orc_crc32c_row(cols, values, n);
  }
  for (uint32_t k = 0; k < n; k++) {
    uint32_t w[3];
    split(values[k], cols[k], w);
    if (overall_words(w) || syndrome_words(w)) return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* matrix                                                              */
/* ------------------------------------------------------------------ */

/* CSR/CPUContext.cpp:11-44: one pass over row-sorted COO triplets; per
 * element encode then store; row pointers filled lazily so empty rows get the
 * index of the next stored element.  Deviation: the reference leaves the
 * pointers of *trailing* empty rows uninitialised (only rows[N] is set, :41);
 * here they are set to nnz. */
orc_matrix *orc_create_matrix(int mode, const uint32_t *columns, const uint32_t *rows,
                              const double *values, int N, int nnz) {
  orc_matrix *m = (orc_matrix *)malloc(sizeof *m);
  m->N = (uint32_t)N;
  m->nnz = (uint32_t)nnz;
  m->mode = mode;
  m->cols = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nnz > 0 ? nnz : 1));
  m->rows = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)N + 1));
  m->values = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));

  uint32_t next_row = 0;
  for (int i = 0; i < nnz; i++) {
    m->values[i] = values[i];
    m->cols[i] = (mode == ORC_MODE_CRC32C) ? columns[i] : orc_encode_column(mode, values[i], columns[i]);
    while (next_row <= rows[i]) m->rows[next_row++] = (uint32_t)i;
  }
  while (next_row <= (uint32_t)N) m->rows[next_row++] = (uint32_t)nnz;
  m->rows[N] = (uint32_t)nnz;

  if (mode == ORC_MODE_CRC32C)
    for (int r = 0; r < N; r++) {
      uint32_t s = m->rows[r], e = m->rows[r + 1];
      crc_encode_row(m->cols + s, m->values + s, e - s);
    }
  return m;
}

void orc_destroy_matrix(orc_matrix *m) {
  if (!m) return;
  free(m->cols);
  free(m->rows);
  free(m->values);
  free(m);
}

/* ------------------------------------------------------------------ */
/* spmv, all modes                                                     */
/* ------------------------------------------------------------------ */

typedef struct {
  orc_event *log;
  int cap;
  int n;
} evlog;

static void push(evlog *l, int kind, uint32_t index, uint32_t bit) {
  if (l->log && l->n < l->cap) {
    l->log[l->n].kind = kind;
    l->log[l->n].index = index;
    l->log[l->n].bit = bit;
  }
  if (l->n < l->cap) l->n++;
}

/* Undo the flip the syndrome points at and persist it (CSR/CPUContext.cpp:
 * 272-278, 320-324, 376-380).  The reference indexes past the 12-byte element
 * when a multi-bit fault yields bit >= 96 (undefined behaviour); the oracle
 * reports the same bit number but leaves the element alone. */
static void correct_bit(orc_matrix *m, uint32_t i, uint32_t w[3], uint32_t syndrome, evlog *l) {
  uint32_t bit = orc_ecc_get_flipped_bit_col8(syndrome);
  if (bit < 96) w[bit / 32] ^= 1u << (bit % 32);
  push(l, ORC_EV_CORRECTED_BIT, i, bit);
}

int orc_spmv(orc_matrix *m, const double *vec, double *result, orc_event *log, int cap, int *n_events) {
  evlog l = {log, cap, 0};
  int fatal = 0;
  const int mode = m->mode;

  for (uint32_t row = 0; row < m->N && !fatal; row++) {
    double tmp = 0.0;
    uint32_t start = m->rows[row];
    uint32_t end = m->rows[row + 1];

    if (mode == ORC_MODE_CONSTRAINTS) {
      /* CSR/CPUContext.cpp:173-182 */
      if (end > m->nnz) { push(&l, ORC_EV_ROW_SIZE, row, 0); fatal = 1; break; }
      if (end < start) { push(&l, ORC_EV_ROW_ORDER, row, 0); fatal = 1; break; }
    }
    if (mode == ORC_MODE_CRC32C) {
      if (crc_check_row(m->cols + start, m->values + start, end - start)) {
        push(&l, ORC_EV_CRC_ROW, row, 0);
        fatal = 1;
        break;
      }
    }

    for (uint32_t i = start; i < end; i++) {
      uint32_t w[3];
      split(m->values[i], m->cols[i], w);
      uint32_t col = w[2];

      switch (mode) {
        case ORC_MODE_NONE: /* CSR/CPUContext.cpp:115-133 */
          break;

        case ORC_MODE_CONSTRAINTS: /* :184-203 */
          if (col >= m->N) { push(&l, ORC_EV_COL_SIZE, i, 0); fatal = 1; }
          else if (i < end - 1 && m->cols[i + 1] <= col) { push(&l, ORC_EV_COL_ORDER, i, 0); fatal = 1; }
          break;

        case ORC_MODE_SED: /* :214-245 */
          if (overall_words(w)) { push(&l, ORC_EV_SED_DETECTED, i, 0); fatal = 1; }
          col &= 0x00FFFFFFu;
          break;

        case ORC_MODE_SEC7: { /* :252-289 */
          uint32_t syndrome = syndrome_words(w);
          if (syndrome) {
            correct_bit(m, i, w, syndrome, &l);
            m->cols[i] = w[2];
            m->values[i] = join(w);
          }
          col = w[2] & 0x00FFFFFFu;
          break;
        }

        case ORC_MODE_SEC8: { /* :297-345 -- syndrome only looked at when the overall parity trips */
          if (overall_words(w)) {
            uint32_t syndrome = syndrome_words(w);
            if (syndrome) correct_bit(m, i, w, syndrome, &l);
            else { w[2] ^= 1u << 24; push(&l, ORC_EV_CORRECTED_PARITY, i, 0); }
            m->cols[i] = w[2];
            m->values[i] = join(w);
          }
          col = w[2] & 0x00FFFFFFu;
          break;
        }

        case ORC_MODE_SECDED: { /* :353-411 */
          uint32_t overall = overall_words(w);
          uint32_t syndrome = syndrome_words(w);
          if (overall) {
            if (syndrome) correct_bit(m, i, w, syndrome, &l);
            else { w[2] ^= 1u << 24; push(&l, ORC_EV_CORRECTED_PARITY, i, 0); }
            m->cols[i] = w[2];
            m->values[i] = join(w);
          } else if (syndrome) {
            push(&l, ORC_EV_DOUBLE_BIT, i, 0);
            fatal = 1;
          }
          col = w[2] & 0x00FFFFFFu;
          break;
        }

        case ORC_MODE_CRC32C:
          col &= 0x00FFFFFFu;
          break;
      }
      if (fatal) break;

      /* the multiply-add of every mode (:128, :202, :240, :284, :340, :406):
       * product rounded, then added -- the reference build has no FMA.
       * Deviation: a corrupted column that no check caught (mode none, a
       * mis-corrected multi-bit fault, the unchecked modes' index flips) makes
       * the reference read vec[] out of bounds (undefined behaviour); the
       * oracle and the CUDA target both read the LAST entry of vec[] instead
       * (the column is clamped to N-1). */
      double xv = vec[col < m->N ? col : m->N - 1];
      double prod = m->values[i] * xv;
      tmp = tmp + prod;
    }
    if (fatal) break;
    result[row] = tmp;
  }
  if (n_events) *n_events = l.n;
  return fatal;
}

/* CSR/CPUContext.cpp:82-90 */
double orc_dot(const double *a, const double *b, int N) {
  double ret = 0.0;
  for (int i = 0; i < N; i++) {
    double prod = a[i] * b[i];
    ret = ret + prod;
  }
  return ret;
}

/* CSR/CPUContext.cpp:92-105 */
double orc_calc_xr(double *x, double *r, const double *p, const double *w, double alpha, int N) {
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

/* CSR/CPUContext.cpp:107-113 */
void orc_calc_p(double *p, const double *r, double beta, int N) {
  for (int i = 0; i < N; i++) {
    double bp = beta * p[i];
    p[i] = r[i] + bp;
  }
}

/* CSR/CPUContext.cpp:150-157 */
void orc_flip_bit(orc_matrix *m, int index, int bit) {
  if (bit < 64) {
    uint64_t bits;
    memcpy(&bits, &m->values[index], 8);
    bits ^= 1ull << bit;
    memcpy(&m->values[index], &bits, 8);
  } else {
    m->cols[index] ^= 1u << (bit % 32);
  }
}

/* CSR/CPUContext.cpp:135-159: one rand() for the element, then one per flip */
void orc_inject_bitflip(orc_matrix *m, int kind, int num_flips, int *picked) {
  int index = rand() % (int)m->nnz;
  int start = 0, end = 96;
  if (kind == 1) end = 64;
  else if (kind == 2) start = 64;
  if (picked) picked[0] = index;
  for (int i = 0; i < num_flips; i++) {
    int bit = (rand() % (end - start)) + start;
    if (picked) picked[1 + i] = bit;
    orc_flip_bit(m, index, bit);
  }
}

/* cg.cpp:83-118 */
int orc_cg_solve(orc_matrix *A, const double *b, double *x, double *r, double *p, double *w,
                 int max_itrs, double conv_threshold, double *rr_hist, orc_event *log, int cap,
                 int *n_events) {
  int N = (int)A->N;
  int total = 0;
  memcpy(r, b, sizeof(double) * (size_t)N);
  memcpy(p, r, sizeof(double) * (size_t)N);
  double rr = orc_dot(r, r, N);
  int itr = 0;
  for (; itr < max_itrs && rr > conv_threshold; itr++) {
    int n = 0;
    int fatal = orc_spmv(A, p, w, log ? log + total : NULL, cap - total, &n);
    total += n;
    if (fatal) {
      if (n_events) *n_events = total;
      return -1 - itr;
    }
    double pw = orc_dot(p, w, N);
    double alpha = rr / pw;
    double rr_new = orc_calc_xr(x, r, p, w, alpha, N);
    double beta = rr_new / rr;
    orc_calc_p(p, r, beta, N);
    rr = rr_new;
    if (rr_hist) rr_hist[itr] = rr;
  }
  if (n_events) *n_events = total;
  return itr;
}

int orc_format_event(const orc_event *e, char *buf, int buflen) {
  switch (e->kind) {
    case ORC_EV_SED_DETECTED: return snprintf(buf, buflen, "[ECC] error detected at index %d", (int)e->index);
    case ORC_EV_CORRECTED_BIT: return snprintf(buf, buflen, "[ECC] corrected bit %u at index %d", e->bit, (int)e->index);
    case ORC_EV_CORRECTED_PARITY: return snprintf(buf, buflen, "[ECC] corrected overall parity bit at index %d", (int)e->index);
    case ORC_EV_DOUBLE_BIT: return snprintf(buf, buflen, "[ECC] double-bit error detected");
    case ORC_EV_ROW_SIZE: return snprintf(buf, buflen, "row size constraint violated for row %d", (int)e->index);
    case ORC_EV_ROW_ORDER: return snprintf(buf, buflen, "row order constraint violated for row%d", (int)e->index);
    case ORC_EV_COL_SIZE: return snprintf(buf, buflen, "column size constraint violated at index %d", (int)e->index);
    case ORC_EV_COL_ORDER: return snprintf(buf, buflen, "column order constraint violated at index %d", (int)e->index);
    case ORC_EV_CRC_ROW: return snprintf(buf, buflen, "[CRC32C] error detected in row %d", (int)e->index);
    default: return snprintf(buf, buflen, "unknown event %d", e->kind);
  }
}
