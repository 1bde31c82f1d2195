// COOCUDAContext.cpp -- the `-t cuda` target of abft-sparse-cg's cg-coo (the COO executable).
//
// A CGContext subclass (reference plugin API: CGContext.h:8-67) whose twelve virtuals forward to the
// C ABI of include/abft_cuda.h -- abft_cuda_coo_* for the matrix, the shared vector / reduction calls
// for the rest -- plus one static CGContext::Register<> object per ABFT mode (CGContext.h:55-63; the
// CPU target does the same at COO/CPUContext.cpp:381-389), so the UNMODIFIED cg.cpp works with
// `cg-coo -t cuda -m none|constraints|sed|sec7|sec8|secded`.
//
// Compiled by g++ (no CUDA headers) against the reference's CGContext.h, found through -I$REF;
// linked after CGContext.o (static-init order, reference Makefile:22-32).
//
// Fault messages: the kernels queue events; this shim prints them with the reference's exact strings
// (COO/CPUContext.cpp:159,164,177,187,214,256,298,307,350,359,371) and exit(1)s on a fatal one at the
// first call that needs the spmv's result, before anything else reaches stdout.
//
// STATUS: the C ABI underneath is parity green on a B200 (tests/test_gpu_coo.py); this binary (cg-coo-cuda) builds and
// registers its six targets but has not been driven end to end on a GPU yet.  Single GPU (ABFT_CUDA_DEVICE selects it).
#include "CGContext.h"  // reference header

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "abft_cuda.h"

// the opaque types of the plugin API are defined per backend (CGContext.h:5-6; compare COO/CPUContext.h:5-22)
struct cg_matrix {
  abft_coo_matrix *m;
  int nnz;
};
struct cg_vector {
  abft_vector *v;
  int N;
};

namespace {

[[noreturn]] void die(const char *what) {
  fflush(stdout);
  fprintf(stderr, "\ncuda target: %s failed: %s\n\n", what, abft_cuda_last_error());
  exit(1);
}

template <int MODE>
class COOCUDAContext : public CGContext {
 public:
  COOCUDAContext() {
    const char *dev = getenv("ABFT_CUDA_DEVICE");
    ctx = abft_cuda_ctx_create(MODE, dev ? atoi(dev) : 0);
    if (!ctx) die("context creation");
  }
  virtual ~COOCUDAContext() {
    report_faults();
    abft_cuda_ctx_destroy(ctx);
  }

  virtual cg_matrix *create_matrix(const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                                   int nnz) {
    cg_matrix *M = new cg_matrix;
    M->nnz = nnz;
    M->m = abft_cuda_coo_matrix_create(ctx, columns, rows, values, N, nnz);
    if (!M->m) die("create_matrix");
    return M;
  }
  virtual void destroy_matrix(cg_matrix *mat) {
    report_faults();
    abft_cuda_coo_matrix_destroy(ctx, mat->m);
    delete mat;
  }

  virtual cg_vector *create_vector(int N) {
    cg_vector *r = new cg_vector;
    r->N = N;
    r->v = abft_cuda_vector_create(ctx, N);
    if (!r->v) die("create_vector");
    return r;
  }
  virtual void destroy_vector(cg_vector *vec) {
    abft_cuda_vector_destroy(ctx, vec->v);
    delete vec;
  }
  virtual double *map_vector(cg_vector *v) {
    report_faults();
    double *h = abft_cuda_vector_map(ctx, v->v);
    if (!h) die("map_vector");
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
    report_faults();
    double out = 0.0;
    if (abft_cuda_calc_xr(ctx, x->v, r->v, p->v, w->v, alpha, &out)) die("calc_xr");
    return out;
  }
  virtual void calc_p(cg_vector *p, const cg_vector *r, double beta) {
    if (abft_cuda_calc_p(ctx, p->v, r->v, beta)) die("calc_p");
  }

  virtual void spmv(const cg_matrix *mat, const cg_vector *vec, cg_vector *result) {
    report_faults();  // anything a previous spmv left behind
    if (abft_cuda_coo_spmv(ctx, mat->m, vec->v, result->v)) die("spmv");
  }

  // COO/CPUContext.cpp:123-140: one rand() for the element, one per flip; VALUE = bits 64..127, INDEX = 0..63
  virtual void inject_bitflip(cg_matrix *mat, BitFlipKind kind, int num_flips) {
    int index = rand() % mat->nnz;
    int start = 0, end = 128;
    if (kind == VALUE) start = 64;
    else if (kind == INDEX) end = 64;
    for (int i = 0; i < num_flips; i++) {
      int bit = (rand() % (end - start)) + start;
      printf("*** flipping bit %d at index %d ***\n", bit, index);
      if (abft_cuda_coo_flip_bit(ctx, mat->m, index, bit)) die("inject_bitflip");
    }
  }

 private:
  abft_ctx *ctx;

  // print what the kernels logged, in the serial order of the reference's loop; exit(1) where it would have
  void report_faults() {
    if (!abft_cuda_events_pending(ctx)) return;
    abft_event ev[256];
    int fatal = 0, n;
    while ((n = abft_cuda_poll_events(ctx, ev, 256, &fatal)) > 0) {  // (handed out 256 at a time until drained)
      char line[160];
      for (int i = 0; i < n; i++) {
        abft_cuda_coo_format_event(&ev[i], line, sizeof line);
        printf("%s\n", line);
      }
      if (fatal) {
        fflush(stdout);
        exit(1);
      }
    }
    if (n < 0) die("poll_events");
  }
};

static CGContext::Register<COOCUDAContext<ABFT_MODE_NONE> > A("cuda", "none");
static CGContext::Register<COOCUDAContext<ABFT_MODE_CONSTRAINTS> > B("cuda", "constraints");
static CGContext::Register<COOCUDAContext<ABFT_MODE_SED> > C("cuda", "sed");
static CGContext::Register<COOCUDAContext<ABFT_MODE_SEC7> > D("cuda", "sec7");
static CGContext::Register<COOCUDAContext<ABFT_MODE_SEC8> > E("cuda", "sec8");
static CGContext::Register<COOCUDAContext<ABFT_MODE_SECDED> > F("cuda", "secded");

}  // namespace
