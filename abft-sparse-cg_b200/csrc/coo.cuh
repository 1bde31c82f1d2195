// coo.cuh -- the COO format of the reference's second executable (cg-coo): 128-bit elements
// {col, row, value} with the check byte in col bits 24..31 (reference: COO/ecc.h:7-16, COO/CPUContext.cpp).
//
// STATUS: parity green on a B200 against oracle/coo_oracle.c, which is pinned against the reference
// (tests/test_gpu_coo.py: encoded elements, result[], events, repaired elements bit for bit; CG iteration counts), in
// both element layouts and with both kernels of the grouped layout.  Measured on 256^3, of the format's own roofline
// (16 B per element + the two vectors; profiles/r2_coo_stream_kernel.jsonl, r2_coo_layouts.jsonl): the streamed kernel
// of coo_stream.cuh (the default) 0.33 ms none / sed = 1.00, 0.39 ms sec7 / secded = 0.83; this file's kernels 0.41 /
// 0.46-0.59 ms (grouped layout) and 0.53 / 0.60 ms (caller-order layout).
//
// The reference's loop is a scatter: for every element in index order, result[col] += value * vec[row]
// (COO/CPUContext.cpp:104-121).  Each output entry therefore accumulates its elements in ascending element
// index.  Here: create_matrix groups the element indices by output entry once (`perm`, `colptr`: a
// stable counting sort by column), and thread c walks group c in order -- check, repair, multiply, add --
// so result[] has the reference's bits, without atomics.  The elements stay where the caller put them
// (fault indices and inject_bitflip refer to that order) and are read through `perm` with 128-bit loads.
//
// An index fault that no check repairs (mode none, an undetected double flip, a mis-corrected multi-bit
// fault) changes where an element's product belongs.  The walking thread notices (the element's column is
// not the group's), leaves it out and puts it on a short "stray" list; a one-thread fix-up kernel then
// recomputes the output entries the strays now point at, merging them into those groups by element index.
// Result: the reference's result[] in every case the reference defines (it reads/writes out of bounds when
// an index leaves [0, N): such an element is dropped here and in the oracle).
//
// GROUPED layout (the default; ABFT_CUDA_COO_LAYOUT=caller at create_matrix keeps the elements where the caller put
// them): the elements are STORED in group order -- slot k of `perm` -- so thread c streams its group with
// consecutive 128-bit loads instead of gathering 16-byte elements through `perm` (which is then only read to name an
// element in a fault record); `pos` (element index -> slot) serves inject_bitflip, the CONSTRAINTS neighbour and the
// stray fix-up.  Everything observable keeps the caller's numbering: fault indices, abft_cuda_coo_matrix_read.
#pragma once
#include "common.cuh"

namespace abft {

constexpr bool coo_is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

// mask of parity p (1..7) over word (0..3) of the element {col, row, value lo, value hi}, including the parity
// bit itself (stored in col bit 32-p).  Rule behind the reference's 28 constants (COO/ecc.h:18-51): the 120 data
// bits -- every element bit except the check byte 24..31 -- take the non-power-of-two code positions 3,5,6,7,9,...
constexpr uint32_t coo_mask(int p, int word) {
  uint32_t m = 0;
  uint32_t pos = 3;
  for (int ebit = 0; ebit < 128; ebit++) {
    if (ebit >= 24 && ebit < 32) continue;
    while (coo_is_pow2(pos)) pos++;
    if (((pos >> (p - 1)) & 1u) && ebit / 32 == word) m |= 1u << (ebit % 32);
    pos++;
  }
  if (word == 0) m |= 1u << (32 - p);
  return m;
}
static_assert(coo_mask(1, 0) == 0x80AAAD5Bu, "P1 word0 (COO/ecc.h:18)");
static_assert(coo_mask(7, 2) == 0xFFFFFFFEu, "P7 word2 (COO/ecc.h:50)");
static_assert(coo_mask(6, 1) == 0xFFFFFFFCu, "P6 word1 (COO/ecc.h:44)");

template <int P>
struct CM {
  static constexpr uint32_t w0 = coo_mask(P, 0), w1 = coo_mask(P, 1), w2 = coo_mask(P, 2), w3 = coo_mask(P, 3);
};

template <int P>
__device__ __forceinline__ uint32_t coo_parity_bit(const uint4 &e) {
  uint32_t t = 0;
  if (CM<P>::w0) t ^= e.x & CM<P>::w0;
  if (CM<P>::w1) t ^= e.y & CM<P>::w1;
  if (CM<P>::w2) t ^= e.z & CM<P>::w2;
  if (CM<P>::w3) t ^= e.w & CM<P>::w3;
  return __popc(t) & 1u;
}

// bit 0 set iff parity P of the element is odd (the raw popcount: callers OR several and test bit 0 once)
template <int P>
__device__ __forceinline__ uint32_t coo_parity_pop(const uint4 &e) {
  uint32_t t = 0;
  if (CM<P>::w0) t ^= e.x & CM<P>::w0;
  if (CM<P>::w1) t ^= e.y & CM<P>::w1;
  if (CM<P>::w2) t ^= e.z & CM<P>::w2;
  if (CM<P>::w3) t ^= e.w & CM<P>::w3;
  return (uint32_t)__popc(t);
}
// bit 0 set iff the syndrome (ecc_compute_col8) is non-zero -- without assembling it
__device__ __forceinline__ uint32_t coo_syndrome_any(const uint4 &e) {
  return coo_parity_pop<1>(e) | coo_parity_pop<2>(e) | coo_parity_pop<3>(e) | coo_parity_pop<4>(e) | coo_parity_pop<5>(e) |
         coo_parity_pop<6>(e) | coo_parity_pop<7>(e);
}

// bit 0 set iff one of the mode's checks trips on this element (nothing is repaired or reported here)
template <int MODE>
__device__ __forceinline__ uint32_t coo_suspect(const uint4 &e) {
  if (MODE == ABFT_MODE_SED || MODE == ABFT_MODE_SEC8) return (uint32_t)__popc(e.x ^ e.y ^ e.z ^ e.w);
  if (MODE == ABFT_MODE_SEC7) return coo_syndrome_any(e);
  if (MODE == ABFT_MODE_SECDED) return coo_syndrome_any(e) | (uint32_t)__popc(e.x ^ e.y ^ e.z ^ e.w);
  return 0u;
}

// ecc_compute_overall_parity (COO/ecc.h:108-112)
__device__ __forceinline__ uint32_t coo_overall(const uint4 &e) { return __popc(e.x ^ e.y ^ e.z ^ e.w) & 1u; }

// ecc_compute_col8 (COO/ecc.h:63-100): parity p -> bit 32-p
__device__ __forceinline__ uint32_t coo_syndrome(const uint4 &e) {
  return (coo_parity_bit<1>(e) << 31) | (coo_parity_bit<2>(e) << 30) | (coo_parity_bit<3>(e) << 29) |
         (coo_parity_bit<4>(e) << 28) | (coo_parity_bit<5>(e) << 27) | (coo_parity_bit<6>(e) << 26) |
         (coo_parity_bit<7>(e) << 25);
}

// ecc_get_flipped_bit_col8 (COO/ecc.h:116-134).  syndrome != 0.
__device__ __forceinline__ uint32_t coo_flipped_bit(uint32_t syndrome) {
  const uint32_t position = __brev(syndrome) & 0x7Fu;
  const uint32_t lz = __clz(position);
  if ((position & (position - 1)) == 0) return lz;  // parity bit p itself: col bit 32-p
  const uint32_t rank = position - (32 - lz) - 1;
  return rank >= 24 ? rank + 8 : rank;               // data bits from rank 24 on sit behind the check byte
}

__device__ __forceinline__ uint32_t &coo_word(uint4 &e, uint32_t w) {
  return w == 0 ? e.x : (w == 1 ? e.y : (w == 2 ? e.z : e.w));
}

// generate_ecc_bits (COO/CPUContext.cpp:196-199 SED, 230-233 SEC7, 269-273 SEC8, 320-324 SECDED)
__global__ void __launch_bounds__(kBlock)
coo_encode_kernel(int mode, const uint32_t *__restrict__ cols, const uint32_t *__restrict__ rows,
                  const double *__restrict__ values, uint4 *__restrict__ elements, uint32_t nnz,
                  const uint32_t *__restrict__ pos = nullptr /* grouped layout: element index -> slot */) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < nnz; i += stride) {
    const double v = values[i];
    uint4 e = make_uint4(cols[i], rows[i], (uint32_t)__double2loint(v), (uint32_t)__double2hiint(v));
    if (mode == ABFT_MODE_SED) {
      e.x |= coo_overall(e) << 31;
    } else if (mode == ABFT_MODE_SEC7) {
      e.x |= coo_syndrome(e);
    } else if (mode == ABFT_MODE_SEC8 || mode == ABFT_MODE_SECDED) {
      e.x |= coo_syndrome(e);
      e.x |= coo_overall(e) << 24;
    }
    elements[pos ? pos[i] : i] = e;
  }
}

// the 4-byte read-modify-write of inject_bitflip (COO/CPUContext.cpp:138)
__global__ void coo_flip_bit_kernel(uint4 *elements, uint32_t index, int bit, const uint32_t *pos = nullptr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  reinterpret_cast<uint32_t *>(elements + (pos ? pos[index] : index))[bit / 32] ^= 1u << (bit % 32);
}

constexpr uint32_t kMaxStray = 1024;

// Host side of create_matrix: group the element indices by output entry (= column), ascending inside a group
// (a stable counting sort).  colptr gets N+1 entries, perm nnz.  Returns false if a column is out of range.
inline bool coo_build_groups(const uint32_t *columns, int N, int nnz, uint32_t *colptr, uint32_t *perm) {
  for (int j = 0; j <= N; j++) colptr[j] = 0;
  for (int i = 0; i < nnz; i++) {
    if (columns[i] >= (uint32_t)N) return false;
    colptr[columns[i] + 1]++;
  }
  for (int j = 0; j < N; j++) colptr[j + 1] += colptr[j];
  // fill each group front to back; colptr[j] is advanced while filling and restored afterwards
  for (int i = 0; i < nnz; i++) perm[colptr[columns[i]]++] = (uint32_t)i;
  for (int j = N; j > 0; j--) colptr[j] = colptr[j - 1];
  colptr[0] = 0;
  return true;
}

struct CooView {
  uint4 *elements;         // [nnz] {col | check byte, row, value lo, value hi}, caller's order
  const uint32_t *perm;    // [nnz] element indices grouped by output entry, ascending inside a group
  const uint32_t *colptr;  // [N+1] group boundaries in perm
  uint32_t *stray;         // [1 + kMaxStray]: count, then element indices (see the file comment)
  uint32_t N, nnz;
  const uint32_t *pos;     // grouped layout only: [nnz] element index -> slot in `elements` (= slot in perm); else null
};

// where element i (caller's numbering) is stored
template <bool GROUPED>
__device__ __forceinline__ uint32_t coo_slot_of(const CooView &A, uint32_t i) { return GROUPED ? A.pos[i] : i; }

// Slow path of coo_check (a syndrome or the overall parity tripped): repair / report as the mode demands.  Out of line
// and by value, so that the element of the fast path stays in registers (coo_word indexes it dynamically).
struct CooFault {
  uint4 e;
  uint32_t ok;   // 0: the reference would have exit(1)ed at this element
};
template <int MODE, bool GROUPED>
__device__ __noinline__ CooFault coo_fault(uint4 e, uint32_t ph, uint32_t k, uint4 *elements, const uint32_t *perm, DevLog *log,
                                           volatile unsigned int *flag) {
  const LogRef L{log, flag};
  const uint32_t odd = (MODE == ABFT_MODE_SEC7) ? 0u : coo_overall(e);
  const uint32_t s = coo_syndrome(e);
  const bool repair = (MODE == ABFT_MODE_SEC7) ? (s != 0) : (odd != 0);
  const uint32_t i = GROUPED ? perm[k] : ph;   // the element's index in the caller's numbering
  if (repair) {
    if (s) {
      const uint32_t bit = coo_flipped_bit(s);
      // a multi-bit fault can give bit >= 128, where the reference writes past its element (UB): reported, left alone
      if (bit < 128) coo_word(e, bit / 32) ^= 1u << (bit % 32);
      log_event(L, ABFT_EV_CORRECTED_BIT, i, bit);
    } else {
      e.x ^= 1u << 24;
      log_event(L, ABFT_EV_CORRECTED_PARITY, i, 0);
    }
    elements[ph] = e;  // mat->elements[i] = element (COO/CPUContext.cpp:254, 311, 363)
  } else if (MODE == ABFT_MODE_SECDED && s) {
    log_event(L, ABFT_EV_DOUBLE_BIT, i, 0);
    return CooFault{e, 0u};
  }
  return CooFault{e, 1u};
}

// Check (and repair) one element the way the mode's spmv does; logs what the reference prints.
// Returns false if the reference would have exit(1)ed at this element.
// `ph`: where the element is stored; `k`: its slot in perm (grouped layout: ph == k and the element's index in the
// caller's numbering, which the fault records carry, is perm[k] -- only read when something is reported).
template <int MODE, bool GROUPED>
__device__ __forceinline__ bool coo_check(uint4 &e, uint32_t ph, uint32_t k, const CooView &A, const LogRef &L) {
  if (MODE == ABFT_MODE_NONE) return true;
  if (MODE == ABFT_MODE_CONSTRAINTS) {
    const uint32_t i = GROUPED ? A.perm[k] : ph;
    // COO/CPUContext.cpp:155-187, in the reference's order; the first violation is fatal.  (`bit` carries the
    // index so that the host orders these among themselves by element; it is cleared before reporting.)
    if (e.y >= A.N) { log_event(L, ABFT_EV_ROW_SIZE, i, 0); return false; }
    if (e.x >= A.N) { log_event(L, ABFT_EV_COL_SIZE, i, i); return false; }
    if (i + 1 < A.nnz) {
      const uint4 nx = A.elements[coo_slot_of<GROUPED>(A, i + 1)];
      if (e.y > nx.y) { log_event(L, ABFT_EV_ROW_ORDER, i, 0); return false; }
      if (e.y == nx.y && e.x >= nx.x) { log_event(L, ABFT_EV_COL_ORDER, i, i); return false; }
    }
    return true;
  }
  if (MODE == ABFT_MODE_SED) {
    if (coo_overall(e)) { log_event(L, ABFT_EV_SED_DETECTED, GROUPED ? A.perm[k] : ph, 0); return false; }
    return true;
  }
  if (MODE == ABFT_MODE_SEC7 || MODE == ABFT_MODE_SEC8 || MODE == ABFT_MODE_SECDED) {
    // fast path: is there anything to do at all?  (SEC8 looks at the syndrome only when the overall parity trips;
    // SEC7 has no overall parity; SECDED acts on either.)  The verdicts are OR-ed as raw popcounts, bit 0 tested once.
    uint32_t any;
    if (MODE == ABFT_MODE_SEC7) any = coo_syndrome_any(e);
    else if (MODE == ABFT_MODE_SEC8) any = (uint32_t)__popc(e.x ^ e.y ^ e.z ^ e.w);
    else any = coo_syndrome_any(e) | (uint32_t)__popc(e.x ^ e.y ^ e.z ^ e.w);
    if (!(any & 1u)) return true;
    const CooFault r = coo_fault<MODE, GROUPED>(e, ph, k, A.elements, A.perm, L.log, L.flag);
    e = r.e;
    return r.ok != 0u;
  }
  return true;
}

template <int MODE>
__device__ __forceinline__ uint32_t coo_out_index(const uint4 &e) {
  return (MODE >= ABFT_MODE_SED) ? (e.x & 0x00FFFFFFu) : e.x;
}

__device__ __forceinline__ double coo_value(const uint4 &e) { return __hiloint2double((int)e.w, (int)e.z); }

// result[c] = sum over group c, in element order, of value * vec[row]   (separate multiply and add: no FMA,
// like the reference build)
// one element of group c: check / repair (e is updated), then: does its product belong into this group's sum?  If not it
// goes onto the stray list (its column is no longer c) or nowhere (fatal, or vec[] would be read out of bounds).
template <int MODE, bool GROUPED>
__device__ __forceinline__ bool coo_prepare(uint4 &e, uint32_t ph, uint32_t k, uint32_t c, const CooView &A, const LogRef &L) {
  if (!coo_check<MODE, GROUPED>(e, ph, k, A, L)) return false;  // fatal at this element: the host stops at the first one in index order
  const uint32_t col = coo_out_index<MODE>(e);
  if (col != c) {
    // the element no longer belongs to this output entry: leave it to the fix-up kernel
    const uint32_t slot = atomicAdd(&A.stray[0], 1u);
    if (slot < kMaxStray) A.stray[1 + slot] = GROUPED ? A.perm[k] : ph;
    return false;
  }
  return e.y < A.N;  // else vec[] would be read out of bounds (UB in the reference): dropped, as in the oracle
}

template <int MODE, bool GROUPED>
__device__ __forceinline__ void coo_accumulate(uint4 e, uint32_t ph, uint32_t k, uint32_t c, const CooView &A,
                                               const double *__restrict__ vec, const LogRef &L, double &sum) {
  if (coo_prepare<MODE, GROUPED>(e, ph, k, c, A, L)) sum = __dadd_rn(sum, __dmul_rn(coo_value(e), vec[e.y]));
}

// grouped layout: group c = slots [kb, ke) of `elements`, in order, straight from global memory -- the next element's
// load in flight while this one is checked (a repair only ever rewrites the element in hand, never the one already
// fetched)
template <int MODE>
__device__ __forceinline__ double coo_group_from_global(const CooView &A, uint32_t kb, uint32_t ke, uint32_t c,
                                                        const double *__restrict__ vec, const LogRef &L) {
  double sum = 0.0;
  uint4 nxt = kb < ke ? A.elements[kb] : make_uint4(0, 0, 0, 0);
  for (uint32_t k = kb; k < ke; k++) {
    const uint4 e = nxt;
    if (k + 1 < ke) nxt = A.elements[k + 1];
    coo_accumulate<MODE, true>(e, k, k, c, A, vec, L, sum);
  }
  return sum;
}

template <int MODE, bool GROUPED>
__global__ void __launch_bounds__(kBlock)
coo_spmv_kernel(CooView A, const double *__restrict__ vec, double *__restrict__ result, LogRef L) {
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t c = blockIdx.x * kBlock + threadIdx.x; c < A.N; c += stride) {
    double sum = 0.0;
    const uint32_t kb = A.colptr[c], ke = A.colptr[c + 1];
    if (GROUPED) {
      sum = coo_group_from_global<MODE>(A, kb, ke, c, vec, L);
    } else {
      for (uint32_t k = kb; k < ke; k++) {
        const uint32_t ph = A.perm[k];
        coo_accumulate<MODE, GROUPED>(A.elements[ph], ph, k, c, A, vec, L, sum);
      }
    }
    result[c] = sum;
  }
}

// One thread.  For every output entry a stray element now points at: recompute that entry from its own
// group (minus the members that strayed away) and the strays that target it, in ascending element index.
// The elements were already checked / repaired by coo_spmv_kernel; nothing is logged here.
// dot_dev / dot_host (null: no fused dot): the spmv kernel's dot(vec, result), still unpublished; every entry rewritten
// here replaces its term, and the value then goes to the mapped word the host polls (common.cuh, grid_sum_finalize).
template <int MODE, bool GROUPED>
__global__ void coo_stray_kernel(CooView A, const double *__restrict__ vec, double *__restrict__ result, LogRef L,
                                 double *dot_dev = nullptr, volatile double *dot_host = nullptr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double dot = dot_dev ? *dot_dev : 0.0;
  // more index faults in one spmv than the list holds: the products of the ones that did not fit are lost, so
  // result[] is wrong -- said out loud (fatal), not silently (bit = 1 tells the host which list overflowed)
  if (A.stray[0] > kMaxStray) log_event(L, ABFT_EV_LOG_OVERFLOW, A.stray[0] - kMaxStray, 1u);
  const uint32_t n = min(A.stray[0], kMaxStray);
  A.stray[0] = 0;
  for (uint32_t s = 0; s < n; s++) {
    const uint32_t target = coo_out_index<MODE>(A.elements[coo_slot_of<GROUPED>(A, A.stray[1 + s])]);
    if (target >= A.N) continue;  // result[] would be written out of bounds (UB in the reference): dropped
    bool seen = false;            // this target was recomputed for an earlier stray
    for (uint32_t q = 0; q < s && !seen; q++)
      seen = coo_out_index<MODE>(A.elements[coo_slot_of<GROUPED>(A, A.stray[1 + q])]) == target;
    if (seen) continue;
    double sum = 0.0;
    uint32_t k = A.colptr[target];
    const uint32_t ke = A.colptr[target + 1];
    uint32_t last = 0;       // strays with an index below `last` have been merged
    bool first = true;
    for (;;) {
      // next candidate from the group (members that still belong to it)
      uint32_t gi = 0xFFFFFFFFu;
      while (k < ke) {
        const uint32_t i = A.perm[k];
        if (coo_out_index<MODE>(A.elements[GROUPED ? k : i]) == target) { gi = i; break; }
        k++;
      }
      // smallest stray index >= last that targets this entry
      uint32_t si = 0xFFFFFFFFu;
      for (uint32_t q = 0; q < n; q++) {
        const uint32_t i = A.stray[1 + q];
        if ((first || i >= last) && i < si && coo_out_index<MODE>(A.elements[coo_slot_of<GROUPED>(A, i)]) == target) si = i;
      }
      const uint32_t i = min(gi, si);
      if (i == 0xFFFFFFFFu) break;
      if (i == gi) k++;
      const uint4 e = A.elements[coo_slot_of<GROUPED>(A, i)];
      if (e.y < A.N) sum = __dadd_rn(sum, __dmul_rn(coo_value(e), vec[e.y]));
      last = i + 1;
      first = false;
    }
    if (dot_dev) dot = __dadd_rn(__dadd_rn(dot, -__dmul_rn(result[target], vec[target])), __dmul_rn(sum, vec[target]));
    result[target] = sum;
  }
  if (dot_dev) {
    *dot_dev = dot;
    unsigned long long bits = (unsigned long long)__double_as_longlong(dot);
    if (bits == kSlotEmpty) bits ^= 1ull;
    if (dot_host) *reinterpret_cast<volatile unsigned long long *>(const_cast<double *>(dot_host)) = bits;
  }
}

}  // namespace abft
