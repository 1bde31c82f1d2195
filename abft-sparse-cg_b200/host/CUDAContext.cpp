// CUDAContext.cpp -- the `-t cuda` target of abft-sparse-cg's cg-csr.
//
// A CGContext subclass (reference plugin API: CGContext.h:8-67) whose twelve
// virtuals forward 1:1 to the C ABI of include/abft_cuda.h, plus one static
// CGContext::Register<> object per ABFT mode (CGContext.h:55-63; the CPU target
// does the same at CSR/CPUContext.cpp:413-421), so the UNMODIFIED cg.cpp,
// matrix-market loader, -b blocking, -x injection and -m modes work with
// `-t cuda -m none|constraints|sed|sec7|sec8|secded|crc32c`.
//
// Compiled by g++ (no CUDA headers needed) against the reference's CGContext.h,
// found through -I$REF at build time; linked after CGContext.o (static-init
// order, reference Makefile:37-43).
//
// Fault messages: the kernels queue events; this shim prints them with the
// reference's exact strings (CSR/CPUContext.cpp:175,180,190,197,233,278,324,
// 331,380,387,398) and exit(1)s on a fatal one at the first call that needs the
// spmv's result (dot / calc_xr / map_vector / the next spmv / destroy), which is
// before anything else reaches stdout -- cg.cpp's next printf comes after
// calc_p (cg.cpp:116).
//
// Environment: ABFT_CUDA_DEVICE=<n> picks the GPU (default 0).
#include "CGContext.h"  // reference header

#include <cstdio>
#include <cstdlib>

#include "abft_cuda.h"

// The opaque types of the plugin API are defined per backend (CGContext.h:5-6;
// compare CSR/CPUContext.h:5-18 and CSR/OCLContext.h:9-22).
struct cg_matrix {
  abft_matrix *m;
  int nnz;
};
struct cg_vector {
  abft_vector *v;
};

namespace {

[[noreturn]] void die(const char *what) {
  fflush(stdout);
  fprintf(stderr, "\ncuda target: %s failed: %s\n\n", what, abft_cuda_last_error());
  exit(1);
}

template <int MODE>
class CUDAContext : public CGContext {
 public:
  CUDAContext() {
    const char *dev = getenv("ABFT_CUDA_DEVICE");
    ctx = abft_cuda_ctx_create(MODE, dev ? atoi(dev) : 0);
    if (!ctx) die("context creation");
  }
  virtual ~CUDAContext() {
    report_faults();
    abft_cuda_ctx_destroy(ctx);
  }

  virtual cg_matrix *create_matrix(const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                                   int nnz) {
    cg_matrix *M = new cg_matrix;
    M->m = abft_cuda_matrix_create(ctx, columns, rows, values, N, nnz);
    M->nnz = nnz;
    if (!M->m) die("create_matrix");
    return M;
  }
  virtual void destroy_matrix(cg_matrix *mat) {
    report_faults();
    abft_cuda_matrix_destroy(ctx, mat->m);
    delete mat;
  }

  virtual cg_vector *create_vector(int N) {
    cg_vector *v = new cg_vector;
    v->v = abft_cuda_vector_create(ctx, N);
    if (!v->v) die("create_vector");
    return v;
  }
  virtual void destroy_vector(cg_vector *vec) {
    abft_cuda_vector_destroy(ctx, vec->v);
    delete vec;
  }
  virtual double *map_vector(cg_vector *v) {
    double *h = abft_cuda_vector_map(ctx, v->v);
    if (!h) die("map_vector");
    report_faults();
    return h;
  }
  virtual void unmap_vector(cg_vector *v, double *h) {
    if (abft_cuda_vector_unmap(ctx, v->v, h)) die("unmap_vector");
  }
  virtual void copy_vector(cg_vector *dst, const cg_vector *src) {
    if (abft_cuda_vector_copy(ctx, dst->v, src->v)) die("copy_vector");
  }

  virtual double dot(const cg_vector *a, const cg_vector *b) {
    double out = 0.0;
    if (abft_cuda_dot(ctx, a->v, b->v, &out)) die("dot");
    report_faults();
    return out;
  }
  virtual double calc_xr(cg_vector *x, cg_vector *r, const cg_vector *p, const cg_vector *w, double alpha) {
    double out = 0.0;
    if (abft_cuda_calc_xr(ctx, x->v, r->v, p->v, w->v, alpha, &out)) die("calc_xr");
    report_faults();
    return out;
  }
  virtual void calc_p(cg_vector *p, const cg_vector *r, double beta) {
    if (abft_cuda_calc_p(ctx, p->v, r->v, beta)) die("calc_p");
  }
  virtual void spmv(const cg_matrix *mat, const cg_vector *vec, cg_vector *result) {
    if (abft_cuda_events_pending(ctx)) report_faults();
    if (abft_cuda_spmv(ctx, mat->m, vec->v, result->v)) die("spmv");
  }

  // CSR/CPUContext.cpp:135-159: one rand() picks the element, one per flip picks
  // the bit; drawn here, in the same order, so a pinned srand() seed flips the
  // same bits as the cpu target.
  virtual void inject_bitflip(cg_matrix *mat, BitFlipKind kind, int num_flips) {
    int index = rand() % mat->nnz;
    int start = 0;
    int end = 96;
    if (kind == VALUE)
      end = 64;
    else if (kind == INDEX)
      start = 64;
    for (int i = 0; i < num_flips; i++) {
      int bit = (rand() % (end - start)) + start;
      printf("*** flipping bit %d at index %d ***\n", bit, index);
      if (abft_cuda_flip_bit(ctx, mat->m, index, bit)) die("inject_bitflip");
    }
  }

 private:
  // print queued fault events the way the reference does, exit(1) on a fatal one
  void report_faults() {
    if (!abft_cuda_events_pending(ctx)) return;
    abft_event ev[256];
    int fatal = 0;
    int n = abft_cuda_poll_events(ctx, ev, 256, &fatal);
    if (n < 0) die("poll_events");
    char line[128];
    for (int i = 0; i < n; i++) {
      abft_cuda_format_event(&ev[i], line, sizeof line);
      printf("%s\n", line);
    }
    if (fatal) {
      fflush(stdout);
      exit(1);
    }
  }

  abft_ctx *ctx;
};

// lower-case names, no '-' inside (run_tests splits `--list` output on '-')
static CGContext::Register<CUDAContext<ABFT_MODE_NONE> > A("cuda", "none");
static CGContext::Register<CUDAContext<ABFT_MODE_CONSTRAINTS> > B("cuda", "constraints");
static CGContext::Register<CUDAContext<ABFT_MODE_SED> > C("cuda", "sed");
static CGContext::Register<CUDAContext<ABFT_MODE_SEC7> > D("cuda", "sec7");
static CGContext::Register<CUDAContext<ABFT_MODE_SEC8> > E("cuda", "sec8");
static CGContext::Register<CUDAContext<ABFT_MODE_SECDED> > F("cuda", "secded");
static CGContext::Register<CUDAContext<ABFT_MODE_CRC32C> > G("cuda", "crc32c");

}  // namespace
