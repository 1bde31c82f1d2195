// coo_emul.cpp -- runs the COO kernels of csrc/coo.cuh on the CPU, one emulated thread after the other (test
// infrastructure; built by tests/test_coo_kernel_logic.py with g++ and the shim in tests/helpers/cuda_shim/).
// The kernels are the product's own source, unmodified; only the CUDA built-ins are replaced.
#include "cuda_runtime.h"  // the shim (found first on the include path)

#include <vector>

#include "../../abft-sparse-cg_b200/csrc/coo.cuh"

using namespace abft;

namespace {
struct Emu {
  int mode = 0;
  std::vector<uint4> elements;
  std::vector<uint32_t> perm, colptr, stray, pos;
  bool grouped = false;
  DevLog log;
  unsigned int flag = 0;
  CooView view{};
};

template <class F>
void run_grid(int grid, int block, F f) {
  gridDim.x = (unsigned)grid;
  blockDim.x = (unsigned)block;
  for (int b = 0; b < grid; b++)
    for (int t = 0; t < block; t++) {
      blockIdx.x = (unsigned)b;
      threadIdx.x = (unsigned)t;
      f();
    }
}

template <int MODE>
void spmv_mode(Emu *e, const double *vec, double *result) {
  LogRef L{&e->log, &e->flag};
  const int grid = std::max(1, std::min((int)((e->view.N + kBlock - 1) / kBlock), 3));  // few CTAs: exercises the grid stride
  if (e->grouped) {
    run_grid(grid, kBlock, [&] { coo_spmv_kernel<MODE, true>(e->view, vec, result, L); });
    run_grid(1, 32, [&] { coo_stray_kernel<MODE, true>(e->view, vec, result, L); });
  } else {
    run_grid(grid, kBlock, [&] { coo_spmv_kernel<MODE, false>(e->view, vec, result, L); });
    run_grid(1, 32, [&] { coo_stray_kernel<MODE, false>(e->view, vec, result, L); });
  }
}
}  // namespace

extern "C" {

void *emu_create(int mode, const uint32_t *cols, const uint32_t *rows, const double *vals, int N, int nnz, int grouped) {
  Emu *e = new Emu();
  e->mode = mode;
  e->grouped = grouped != 0 && nnz > 0;
  e->elements.resize((size_t)std::max(nnz, 1));
  e->perm.resize((size_t)std::max(nnz, 1));
  e->colptr.resize((size_t)N + 1);
  e->stray.assign(1 + kMaxStray, 0u);
  memset(&e->log, 0, sizeof e->log);
  if (!coo_build_groups(cols, N, nnz, e->colptr.data(), e->perm.data())) {
    delete e;
    return nullptr;
  }
  if (e->grouped) {
    e->pos.resize((size_t)nnz);
    for (int k = 0; k < nnz; k++) e->pos[e->perm[(size_t)k]] = (uint32_t)k;
  }
  const uint32_t *pos = e->grouped ? e->pos.data() : nullptr;
  run_grid(2, kBlock, [&] { coo_encode_kernel(mode, cols, rows, vals, e->elements.data(), (uint32_t)nnz, pos); });
  e->view.elements = e->elements.data();
  e->view.perm = e->perm.data();
  e->view.colptr = e->colptr.data();
  e->view.stray = e->stray.data();
  e->view.N = (uint32_t)N;
  e->view.nnz = (uint32_t)nnz;
  e->view.pos = pos;
  return e;
}

void emu_destroy(void *h) { delete (Emu *)h; }
// the elements in the caller's numbering (what abft_cuda_coo_matrix_read hands back), in a buffer owned by the emulator
void *emu_elements(void *h) {
  Emu *e = (Emu *)h;
  if (!e->grouped) return e->elements.data();
  static std::vector<uint4> out;
  out.resize(e->elements.size());
  for (size_t k = 0; k < (size_t)e->view.nnz; k++) out[e->perm[k]] = e->elements[k];
  return out.data();
}

void emu_flip(void *h, int index, int bit) {
  Emu *e = (Emu *)h;
  run_grid(1, 32, [&] { coo_flip_bit_kernel(e->elements.data(), (uint32_t)index, bit, e->view.pos); });
}

// returns the number of raw (unsorted) log records copied to out[0..cap)
int emu_spmv(void *h, const double *vec, double *result, abft_event *out, int cap) {
  Emu *e = (Emu *)h;
  e->log.count = 0;
  switch (e->mode) {
    case ABFT_MODE_NONE: spmv_mode<ABFT_MODE_NONE>(e, vec, result); break;
    case ABFT_MODE_CONSTRAINTS: spmv_mode<ABFT_MODE_CONSTRAINTS>(e, vec, result); break;
    case ABFT_MODE_SED: spmv_mode<ABFT_MODE_SED>(e, vec, result); break;
    case ABFT_MODE_SEC7: spmv_mode<ABFT_MODE_SEC7>(e, vec, result); break;
    case ABFT_MODE_SEC8: spmv_mode<ABFT_MODE_SEC8>(e, vec, result); break;
    case ABFT_MODE_SECDED: spmv_mode<ABFT_MODE_SECDED>(e, vec, result); break;
    default: return -1;
  }
  int n = (int)std::min<unsigned>(e->log.count, (unsigned)std::min(cap, kLogCap));
  for (int i = 0; i < n; i++) out[i] = e->log.ev[i];
  return n;
}

}  // extern "C"
