// csr_emul.cpp -- runs the CSR kernels of csrc/{spmv,encode,vecops}.cuh on the CPU, every CTA as 256 real threads
// (test infrastructure; built by tests/test_csr_kernel_logic.py with g++ -std=c++20 and the shim in
// tests/helpers/cuda_shim_mt/).  The kernels are the product's own source; only the CUDA built-ins and, under
// ABFT_HOST_EMU, the PTX helpers of spmv.cuh (mbarrier, bulk copy, fences, the predicated add) are replaced.
#include "cuda_runtime.h"  // the shim (found first on the include path)

#include <vector>

#include "../../abft-sparse-cg_b200/csrc/encode.cuh"
#include "../../abft-sparse-cg_b200/csrc/vecops.cuh"

using namespace abft;

namespace {
struct Emu {
  int mode = 0;
  uint32_t n_rows = 0, n_cols = 0, nnz = 0;
  uint32_t row_offset = 0, nnz_offset = 0;  // a row partition: global index of local row 0 / local element 0
  std::vector<uint32_t> cols, rowptr;  // padded like abft_cuda_matrix_create pads the device arrays
  std::vector<double> vals;
  std::vector<Tile> tiles;
  bool dyn = false;
  int kind = 0;  // abft_cuda_plan_matrix: 0 static round-robin, 1 dynamic schedule
  int flen = 0;  // stencil matrix: its longest row if 5 or 7 (abft_cuda.cu picks the same instantiation)
  bool rows_ecc = false;
  unsigned int tile_counter = 0;
  DevLog log;
  unsigned int flag = 0;
  std::vector<double> partials;
  unsigned int counter = 0;
  double scal_dev = 0.0, scal_host = 0.0;
  unsigned char *smem = nullptr;
  int grid = 3;  // few CTAs: each takes several tiles (exercises the stage ring and the grid stride)
};

ReduceOut reduce_of(Emu *e) {
  ReduceOut r{};
  r.partials = e->partials.data();
  r.counter = &e->counter;
  r.dev = &e->scal_dev;
  r.host = &e->scal_host;
  return r;
}

void fill_crc_tables() {
  for (int s = 0; s < 4; s++)
    for (uint32_t i = 0; i < 256; i++) g_crc_tab[s * 256 + i] = crc32c_table_entry(i, s);
  for (int m = 1; m <= kCrcMaps; m++)
    for (int j = 0; j < kCrcChunks; j++)
      for (uint32_t c = 0; c < 32; c++) g_crc5_tab[((m - 1) * kCrcChunks + j) * 32 + c] = crc5_entry(m, j, c);
  for (int i = 0; i < kCrcElemChunks; i++)
    for (uint32_t c = 0; c < 32; c++) g_crc_elem_tab[i * 32 + c] = crc_elem_entry(i, c);
}

template <int MODE, bool DYN>
void spmv_inst(Emu *e, const double *vec, double *result) {
  MatView A{};
  A.rowptr = e->rowptr.data();
  A.cols = e->cols.data();
  A.values = e->vals.data();
  A.tiles = e->tiles.data();
  A.tile_counter = DYN ? &e->tile_counter : nullptr;
  A.n_tiles = (uint32_t)e->tiles.size();
  A.n_rows = e->n_rows;
  A.nnz = e->nnz;
  A.n_cols = e->n_cols;
  A.row_offset = e->row_offset;
  A.nnz_offset = e->nnz_offset;
  A.rows_ecc = e->rows_ecc ? 1u : 0u;
  LogRef L{&e->log, &e->flag};
  const ReduceOut red = reduce_of(e);
  const int grid = std::max(1, std::min((int)e->tiles.size(), e->grid));
  constexpr bool has_stencil = !DYN && ((MODE & 7) == ABFT_MODE_SEC7 || (MODE & 7) == ABFT_MODE_SECDED || (MODE & 7) == ABFT_MODE_CRC32C);
  if (has_stencil && e->flen == 7)
    emu::launch(grid, SpmvCfg<false>::block, [&] { spmv_kernel<has_stencil ? MODE : ABFT_MODE_SECDED, false, 7>(A, vec, result, L, red, nullptr, WaitSpec{nullptr, 0u, 0u}); });
  else if (has_stencil && e->flen == 5)
    emu::launch(grid, SpmvCfg<false>::block, [&] { spmv_kernel<has_stencil ? MODE : ABFT_MODE_SECDED, false, 5>(A, vec, result, L, red, nullptr, WaitSpec{nullptr, 0u, 0u}); });
  else
    emu::launch(grid, SpmvCfg<DYN>::block, [&] { spmv_kernel<MODE, DYN>(A, vec, result, L, red, nullptr, WaitSpec{nullptr, 0u, 0u}); });
}

template <int MODE>
void spmv_mode(Emu *e, const double *vec, double *result) {
  if (e->dyn) spmv_inst<MODE, true>(e, vec, result);
  else spmv_inst<MODE, false>(e, vec, result);
}
}  // namespace

extern "C" {

int emu_tile_nnz(int dyn) { return dyn ? kTileNnzDyn : kTileNnz; }
int emu_spmv_block(int dyn) { return dyn ? SpmvCfg<true>::block : SpmvCfg<false>::block; }

// create_matrix as abft_cuda_matrix_create_part does it: row pointers and check bits by the device kernels
// a row partition, as abft_cuda_matrix_create_part takes it: n_rows local rows starting at global row row_offset,
// coo_rows / cols are GLOBAL indices, nnz_offset = global index of the partition's first element
void *emu_create_part(int mode, const uint32_t *cols, const uint32_t *coo_rows, const double *vals, int N, int nnz,
                      int N_global, int row_offset, int nnz_offset);

void *emu_create(int mode, const uint32_t *cols, const uint32_t *coo_rows, const double *vals, int N, int nnz) {
  return emu_create_part(mode, cols, coo_rows, vals, N, nnz, N, 0, 0);
}

void *emu_create_part(int mode, const uint32_t *cols, const uint32_t *coo_rows, const double *vals, int N, int nnz,
                      int N_global, int row_offset, int nnz_offset) {
  Emu *e = new Emu();
  e->mode = mode;
  e->n_rows = (uint32_t)N;
  e->n_cols = (uint32_t)N_global;
  e->row_offset = (uint32_t)row_offset;
  e->nnz_offset = (uint32_t)nnz_offset;
  e->nnz = (uint32_t)nnz;
  const size_t nz = (size_t)std::max(nnz, 1);
  e->cols.assign(nz + 16, 0u);
  e->vals.assign(nz + 16, 0.0);
  e->rowptr.assign((size_t)N + 1 + 16, 0u);
  for (int i = 0; i < nnz; i++) { e->cols[i] = cols[i]; e->vals[i] = vals[i]; }
  memset(&e->log, 0, sizeof e->log);
  e->partials.assign(1 << 16, 0.0);
  e->smem = (unsigned char *)aligned_alloc(128, (std::max(SpmvCfg<true>::dyn_smem, SpmvCfg<false>::dyn_smem) + 127) / 128 * 128);
  fill_crc_tables();
  emu::launch(2, kBlock, [&] { rowptr_kernel(coo_rows, (uint32_t)nnz, (uint32_t)N, (uint32_t)row_offset, e->rowptr.data()); });
  const int base = mode == ABFT_MODE_SECDEDV ? ABFT_MODE_SECDED : mode;
  if (base >= ABFT_MODE_SED && base <= ABFT_MODE_SECDED && nnz > 0)
    emu::launch(2, kBlock, [&] { encode_elements_kernel(base, e->cols.data(), e->vals.data(), (uint32_t)nnz); });
  else if (base == ABFT_MODE_CRC32C && N > 0)
    emu::launch(2, kBlock, [&] { encode_crc_rows_kernel(e->rowptr.data(), e->cols.data(), e->vals.data(), (uint32_t)N); });
  return e;
}

void emu_destroy(void *h) {
  Emu *e = (Emu *)h;
  free(e->smem);
  delete e;
}

// the work decomposition (from abft_cuda_plan_tiles of the real library: host logic, no device needed)
void emu_set_tiles(void *h, const uint32_t *recs, int n_tiles, int kind, int grid) {
  Emu *e = (Emu *)h;
  e->tiles.resize((size_t)n_tiles);
  for (int i = 0; i < n_tiles; i++) e->tiles[i] = Tile{recs[4 * i], recs[4 * i + 1], recs[4 * i + 2], recs[4 * i + 3]};
  e->kind = kind;
  e->dyn = kind == 1;
  e->grid = grid;
  e->flen = 0;
  if (kind == 0 && n_tiles > 0) {   // the rule of abft_cuda_matrix_create_part
    uint32_t longest = 0;
    bool all_regular = true;
    for (const Tile &t : e->tiles) {
      all_regular &= (t.info & kTileRegular) != 0;
      longest = std::max(longest, (t.info >> kTileMaxLenShift) & 0xFFFu);
    }
    if (all_regular && (longest == 5 || longest == 7) && !getenv("ABFT_EMU_GENERIC")) e->flen = (int)longest;
  }
}
int emu_flen(void *h) { return ((Emu *)h)->flen; }

// row-pointer protection (rowsecc.cuh), as abft_cuda_matrix_create_part applies it
int emu_protect_rows(void *h) {
  Emu *e = (Emu *)h;
  const uint32_t n_words = (e->n_rows + 1u + 1u) & ~1u;
  emu::launch(2, kBlock, [&] { encode_rowptr_pairs_kernel(e->rowptr.data(), n_words); });
  e->rows_ecc = true;
  return 1;
}

uint32_t *emu_cols(void *h) { return ((Emu *)h)->cols.data(); }
uint32_t *emu_rowptr(void *h) { return ((Emu *)h)->rowptr.data(); }
double *emu_values(void *h) { return ((Emu *)h)->vals.data(); }

void emu_flip(void *h, int index, int bit) {
  Emu *e = (Emu *)h;
  emu::launch(1, 32, [&] { flip_bit_kernel(e->cols.data(), e->vals.data(), (uint32_t)index, bit); });
}

// returns the number of raw (unsorted) log records copied to out[0..cap); *dot_out = the fused dot(vec, result)
int emu_spmv(void *h, const double *vec, double *result, abft_event *out, int cap, double *dot_out) {
  Emu *e = (Emu *)h;
  e->log.count = 0;
  e->flag = 0;
  emu::g_dyn_smem = e->smem;
  if (e->tiles.empty()) {
    *dot_out = 0.0;
    return 0;
  }
  switch (e->mode) {
    case ABFT_MODE_NONE: spmv_mode<ABFT_MODE_NONE>(e, vec, result); break;
    case ABFT_MODE_CONSTRAINTS: spmv_mode<ABFT_MODE_CONSTRAINTS>(e, vec, result); break;
    case ABFT_MODE_SED: spmv_mode<ABFT_MODE_SED>(e, vec, result); break;
    case ABFT_MODE_SEC7: spmv_mode<ABFT_MODE_SEC7>(e, vec, result); break;
    case ABFT_MODE_SEC8: spmv_mode<ABFT_MODE_SEC8>(e, vec, result); break;
    case ABFT_MODE_SECDED: spmv_mode<ABFT_MODE_SECDED>(e, vec, result); break;
    case ABFT_MODE_CRC32C: spmv_mode<ABFT_MODE_CRC32C>(e, vec, result); break;
    case ABFT_MODE_SECDEDV: spmv_mode<ABFT_MODE_SECDED | kVecSed>(e, vec, result); break;
    default: return -1;
  }
  *dot_out = e->scal_host;
  if (e->scal_host != e->scal_dev && !(e->scal_host != e->scal_host)) return -2;  // device slot and host mirror differ
  if (e->counter != 0 || e->tile_counter != 0) return -3;                         // not re-armed for the next launch
  const int n = (int)std::min<unsigned>(e->log.count, (unsigned)std::min(cap, kLogCap));
  for (int i = 0; i < n; i++) out[i] = e->log.ev[i];
  return n;
}

// the vector kernels through the same machinery (host-scalar flavour); mode secdedv: with the vectors' parity bit.
// Fault records of these kernels are left in the log for emu_vec_events.
static bool vsed_of(Emu *e) { return e->mode == ABFT_MODE_SECDEDV; }
double emu_dot(void *h, const double *a, const double *b, int N, int grid) {
  Emu *e = (Emu *)h;
  const ReduceOut red = reduce_of(e);
  LogRef L{&e->log, &e->flag};
  if (vsed_of(e)) emu::launch(grid, kBlock, [&] { dot_kernel<true>(a, b, N, 1, red, nullptr, L); });
  else emu::launch(grid, kBlock, [&] { dot_kernel<false>(a, b, N, 1, red, nullptr, L); });
  return e->scal_host;
}
double emu_calc_xr(void *h, double *x, double *r, const double *p, const double *w, double alpha, int N, int grid) {
  Emu *e = (Emu *)h;
  const ReduceOut red = reduce_of(e);
  LogRef L{&e->log, &e->flag};
  if (vsed_of(e)) emu::launch(grid, kBlock, [&] { calc_xr_kernel<false, true>(x, r, p, w, ScalarArg{alpha, nullptr, nullptr}, N, 1, red, nullptr, L); });
  else emu::launch(grid, kBlock, [&] { calc_xr_kernel<false, false>(x, r, p, w, ScalarArg{alpha, nullptr, nullptr}, N, 1, red, nullptr, L); });
  return e->scal_host;
}
void emu_calc_p(void *h, double *p, const double *r, double beta, int N, int grid) {
  Emu *e = (Emu *)h;
  LogRef L{&e->log, &e->flag};
  if (vsed_of(e)) emu::launch(grid, kBlock, [&] { calc_p_kernel<false, true>(p, r, ScalarArg{beta, nullptr, nullptr}, N, 1, nullptr, L); });
  else emu::launch(grid, kBlock, [&] { calc_p_kernel<false, false>(p, r, ScalarArg{beta, nullptr, nullptr}, N, 1, nullptr, L); });
}
void emu_vec_encode(double *v, int N) { emu::launch(2, kBlock, [&] { vec_encode_kernel(v, N); }); }
// raw records the vector kernels logged since the last spmv / emu_vec_events
int emu_vec_events(void *h, abft_event *out, int cap) {
  Emu *e = (Emu *)h;
  const int n = (int)std::min<unsigned>(e->log.count, (unsigned)std::min(cap, kLogCap));
  for (int i = 0; i < n; i++) out[i] = e->log.ev[i];
  e->log.count = 0;
  e->flag = 0;
  return n;
}

}  // extern "C"
