/* rowptr_ecc.c -- CPU restatement of the row-pointer SECDED code of abft-sparse-cg_b200/csrc/rowsecc.cuh.
 *
 * TEST INFRASTRUCTURE (only tests/ may call it).  PARITY UNPINNED: the reference does not protect rows[]
 * (CSR/CPUContext.cpp:123-124); this code is defined by this repository (scope row f.2) and the oracle is
 * an independent restatement of that definition -- bit-by-bit loops over code positions, no mask tables -- not of
 * anything in /root/reference.
 *
 * Code word = {rows[2k], rows[2k+1]}: data bit d (0..55; word d/28, bit d%28) sits at the d-th non-power-of-two
 * position >= 3; Hamming parity p (1..6) sits at position 2^(p-1), stored in word0 bits 28..31 (p = 1..4) and word1
 * bits 28..29 (p = 5, 6); word1 bit 30 = parity of all 64 bits; word1 bit 31 = 0.
 */
#include <stdint.h>

static int is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

/* code position (1..63) of code-word bit b (0..63); 0: the overall parity bit / the spare bit (in no Hamming parity) */
static uint32_t position_of_bit(uint32_t b) {
  const uint32_t word = b / 32, bit = b % 32;
  if (bit < 28) {
    uint32_t d = word * 28 + bit, pos = 3;
    for (;;) {
      while (is_pow2(pos)) pos++;
      if (d == 0) return pos;
      d--;
      pos++;
    }
  }
  if (word == 0) return 1u << (bit - 28);        /* P1..P4 */
  if (bit < 30) return 1u << (bit - 28 + 4);     /* P5, P6 */
  return 0;
}

static uint32_t get_bit(uint32_t w0, uint32_t w1, uint32_t b) { return ((b < 32 ? w0 : w1) >> (b % 32)) & 1u; }

/* xor of the positions of all set bits = the syndrome; also the overall parity */
static void measure(uint32_t w0, uint32_t w1, uint32_t *syndrome, uint32_t *odd) {
  uint32_t s = 0, par = 0;
  for (uint32_t b = 0; b < 64; b++)
    if (get_bit(w0, w1, b)) {
      s ^= position_of_bit(b);
      par ^= 1u;
    }
  *syndrome = s;
  *odd = par;
}

void rpo_encode_pair(uint32_t *w0, uint32_t *w1) {
  uint32_t s, odd;
  *w0 &= 0x0FFFFFFFu;
  *w1 &= 0x0FFFFFFFu;
  measure(*w0, *w1, &s, &odd);
  /* set each Hamming parity bit so that its parity class becomes even */
  for (uint32_t p = 1; p <= 6; p++)
    if ((s >> (p - 1)) & 1u) {
      if (p <= 4) *w0 |= 1u << (27 + p);
      else *w1 |= 1u << (23 + p);
    }
  measure(*w0, *w1, &s, &odd);
  if (odd) *w1 |= 1u << 30;
}

/* 0 clean, 1 corrected (*bit = code-word bit), 2 uncorrectable */
int rpo_check_pair(uint32_t *w0, uint32_t *w1, uint32_t *bit) {
  uint32_t s, odd;
  measure(*w0, *w1, &s, &odd);
  if (!odd) return s ? 2 : 0;
  uint32_t b = 64;
  if (s) {
    for (uint32_t c = 0; c < 64; c++)
      if (position_of_bit(c) == s) b = c;
    if (b == 64) return 2;
  } else {
    b = ((*w1 >> 31) & 1u) ? 63u : 62u;
  }
  if (b < 32) *w0 ^= 1u << b;
  else *w1 ^= 1u << (b - 32);
  *bit = b;
  return 1;
}

/* encode n_words (even, caller pads) pointers in place */
void rpo_encode(uint32_t *rows, int n_words) {
  for (int k = 0; k + 1 < n_words; k += 2) rpo_encode_pair(&rows[k], &rows[k + 1]);
}

/* The spmv pre-pass: every pair in ascending order; corrections repair rows[] in place.  Events go to ev[3*i] =
 * {kind (10 corrected, 11 uncorrectable), pair, bit}; stops at the first uncorrectable pair.  Returns the number of
 * events; *fatal says whether the last one is the uncorrectable one. */
int rpo_check_all(uint32_t *rows, int n_words, uint32_t *ev, int cap, int *fatal) {
  int n = 0;
  *fatal = 0;
  for (int k = 0; k + 1 < n_words; k += 2) {
    uint32_t bit = 0;
    const int r = rpo_check_pair(&rows[k], &rows[k + 1], &bit);
    if (r == 0) continue;
    if (n < cap) {
      ev[3 * n] = (r == 1) ? 10u : 11u;
      ev[3 * n + 1] = (uint32_t)(k / 2);
      ev[3 * n + 2] = (r == 1) ? bit : 0u;
    }
    n++;
    if (r == 2) {
      *fatal = 1;
      break;
    }
  }
  return n;
}
