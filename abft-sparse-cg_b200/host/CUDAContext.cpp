// CUDAContext.cpp -- the `-t cuda` target of abft-sparse-cg's cg-csr.
//
// A CGContext subclass (reference plugin API: CGContext.h:8-67) whose twelve
// virtuals forward to the C ABI of include/abft_cuda.h, plus one static
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
// Environment:
//   ABFT_CUDA_DEVICE=<n>   first GPU to use (default 0)
//   ABFT_CUDA_NGPUS=<g>    row-partition the matrix over g GPUs of this box (default 1).
//                          One host thread drives all of them; the p exchange and the two scalar
//                          all-reduces per iteration happen inside the kernels over peer memory
//                          (abft_cuda_dist_*), so every launch of a step is issued on all GPUs before
//                          the first wait.
#include "CGContext.h"  // reference header

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "abft_cuda.h"

// The opaque types of the plugin API are defined per backend (CGContext.h:5-6;
// compare CSR/CPUContext.h:5-18 and CSR/OCLContext.h:9-22).  One entry per GPU.
struct cg_matrix {
  std::vector<abft_matrix *> part;
  int nnz;
};
struct cg_vector {
  int N;
  std::vector<abft_vector *> full;  // per GPU: the whole vector (G == 1: the vector itself)
  std::vector<abft_vector *> loc;   // per GPU: view of the rows that GPU owns (G > 1)
  std::vector<abft_comm *> comm;    // per GPU: exchange window of this vector (G > 1)
  bool prepared;
  bool dirty;    // local slices changed since the peers' copies were last refreshed
  double *host;  // map_vector staging for the gathered vector (G > 1)
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
  CUDAContext() : have_cuts(false) {
    const char *dev = getenv("ABFT_CUDA_DEVICE");
    const char *ng = getenv("ABFT_CUDA_NGPUS");
    const int base = dev ? atoi(dev) : 0;
    G = ng ? atoi(ng) : 1;
    if (G < 1) G = 1;
    if (G > 8) G = 8;
    for (int d = 0; d < G; d++) {
      abft_ctx *c = abft_cuda_ctx_create(MODE, base + d);
      if (!c) die("context creation");
      ctx.push_back(c);
      device.push_back(base + d);
    }
  }
  virtual ~CUDAContext() {
    report_faults();
    for (int d = 0; d < G; d++) abft_cuda_ctx_destroy(ctx[d]);
  }

  virtual cg_matrix *create_matrix(const uint32_t *columns, const uint32_t *rows, const double *values, int N,
                                   int nnz) {
    cg_matrix *M = new cg_matrix;
    M->nnz = nnz;
    if (G == 1) {
      M->part.push_back(abft_cuda_matrix_create(ctx[0], columns, rows, values, N, nnz));
      if (!M->part[0]) die("create_matrix");
      return M;
    }
    // contiguous row blocks balanced by non-zeros; triplets are sorted by row (cg.cpp:394)
    cut.assign(G + 1, 0);
    off.assign(G + 1, 0);
    cut[G] = N;
    off[G] = nnz;
    for (int k = 1; k < G; k++) {
      long long e = (long long)nnz * k / G;
      int row = (e < nnz) ? (int)rows[e] : N;
      if (MODE == ABFT_MODE_SECDEDV) row &= ~1;  // protected row pointers are coded in pairs: blocks start on even rows
      if (row < cut[k - 1]) row = cut[k - 1];
      cut[k] = row;
      long long lo = off[k - 1], hi = nnz;  // first element whose row >= cut[k]
      while (lo < hi) {
        long long mid = (lo + hi) / 2;
        if ((int)rows[mid] < row) lo = mid + 1; else hi = mid;
      }
      off[k] = lo;
    }
    have_cuts = true;
    // which rows of every block each block's columns reference: the exchange plan
    need_lo.assign(G, std::vector<int>(G, 0));
    need_hi.assign(G, std::vector<int>(G, 0));
    for (int d = 0; d < G; d++)
      for (long long i = off[d]; i < off[d + 1]; i++) {
        const int col = (int)columns[i];
        int s = 0;
        while (s + 1 < G && col >= cut[s + 1]) s++;
        if (need_hi[d][s] == need_lo[d][s]) { need_lo[d][s] = col; need_hi[d][s] = col + 1; }
        else { if (col < need_lo[d][s]) need_lo[d][s] = col; if (col >= need_hi[d][s]) need_hi[d][s] = col + 1; }
      }
    for (int d = 0; d < G; d++) {
      abft_matrix *p = abft_cuda_matrix_create_part(ctx[d], columns + off[d], rows + off[d], values + off[d],
                                                    cut[d + 1] - cut[d], (int)(off[d + 1] - off[d]), N, cut[d], off[d]);
      if (!p) die("create_matrix (partition)");
      M->part.push_back(p);
    }
    return M;
  }
  virtual void destroy_matrix(cg_matrix *mat) {
    report_faults();
    for (int d = 0; d < G; d++) abft_cuda_matrix_destroy(ctx[d], mat->part[d]);
    delete mat;
  }

  virtual cg_vector *create_vector(int N) {
    cg_vector *v = new cg_vector;
    v->N = N;
    v->prepared = false;
    v->dirty = true;
    v->host = NULL;
    for (int d = 0; d < G; d++) {
      abft_vector *a = abft_cuda_vector_create(ctx[d], N);
      if (!a) die("create_vector");
      v->full.push_back(a);
    }
    return v;
  }
  virtual void destroy_vector(cg_vector *vec) {
    for (int d = 0; d < G; d++) {
      if (vec->prepared) {
        abft_cuda_comm_destroy(ctx[d], vec->comm[d]);
        abft_cuda_vector_destroy(ctx[d], vec->loc[d]);
      }
    }
    for (int d = 0; d < G; d++) abft_cuda_vector_destroy(ctx[d], vec->full[d]);
    abft_cuda_host_free(vec->host);
    delete vec;
  }
  virtual double *map_vector(cg_vector *v) {
    if (G == 1) {
      double *h = abft_cuda_vector_map(ctx[0], v->full[0]);
      if (!h) die("map_vector");
      report_faults();
      return h;
    }
    prepare(v);
    // one pinned host image of the vector (usable with every GPU); each GPU copies its slice straight into it, all
    // copies in flight before the first wait
    if (!v->host) v->host = (double *)abft_cuda_host_alloc(sizeof(double) * (size_t)(v->N > 0 ? v->N : 1));
    if (!v->host) die("map_vector (host buffer)");
    for (int d = 0; d < G; d++)
      if (abft_cuda_vector_download_async(ctx[d], v->loc[d], v->host + cut[d])) die("map_vector");
    for (int d = 0; d < G; d++)
      if (abft_cuda_ctx_sync(ctx[d])) die("map_vector (sync)");
    report_faults();
    return v->host;
  }
  virtual void unmap_vector(cg_vector *v, double *h) {
    if (G == 1) {
      if (abft_cuda_vector_unmap(ctx[0], v->full[0], h)) die("unmap_vector");
      return;
    }
    prepare(v);
    for (int d = 0; d < G; d++)
      if (abft_cuda_vector_upload_async(ctx[d], v->loc[d], h + cut[d])) die("unmap_vector");
    for (int d = 0; d < G; d++)   // (the caller may reuse h as soon as this returns)
      if (abft_cuda_ctx_sync(ctx[d])) die("unmap_vector (sync)");
    v->dirty = true;
  }
  virtual void copy_vector(cg_vector *dst, const cg_vector *src) {
    if (G == 1) {
      if (abft_cuda_vector_copy(ctx[0], dst->full[0], src->full[0])) die("copy_vector");
      return;
    }
    prepare(dst);
    prepare(const_cast<cg_vector *>(src));
    for (int d = 0; d < G; d++)
      if (abft_cuda_vector_copy(ctx[d], dst->loc[d], src->loc[d])) die("copy_vector");
    dst->dirty = true;
  }

  virtual double dot(const cg_vector *a, const cg_vector *b) {
    double out = 0.0;
    if (G == 1) {
      if (abft_cuda_dot(ctx[0], a->full[0], b->full[0], &out)) die("dot");
    } else {
      cg_vector *A = const_cast<cg_vector *>(a), *B = const_cast<cg_vector *>(b);
      prepare(A);
      prepare(B);
      int slot[8];
      for (int d = 0; d < G; d++)  // every GPU's part is launched before the first wait
        if (abft_cuda_dist_dot_async(ctx[d], A->comm[d], A->loc[d], B->loc[d], &slot[d])) die("dot");
      for (int d = 0; d < G; d++)
        if (abft_cuda_ctx_sync(ctx[d])) die("dot (sync)");
      out = abft_cuda_scalar_host(ctx[0], slot[0]);
    }
    report_faults();
    return out;
  }
  virtual double calc_xr(cg_vector *x, cg_vector *r, const cg_vector *p, const cg_vector *w, double alpha) {
    double out = 0.0;
    if (G == 1) {
      if (abft_cuda_calc_xr(ctx[0], x->full[0], r->full[0], p->full[0], w->full[0], alpha, &out)) die("calc_xr");
    } else {
      cg_vector *P = const_cast<cg_vector *>(p), *W = const_cast<cg_vector *>(w);
      prepare(x); prepare(r); prepare(P); prepare(W);
      int slot[8];
      for (int d = 0; d < G; d++)
        if (abft_cuda_dist_calc_xr_async(ctx[d], x->comm[d], x->loc[d], r->loc[d], P->loc[d], W->loc[d], alpha, &slot[d]))
          die("calc_xr");
      for (int d = 0; d < G; d++)
        if (abft_cuda_ctx_sync(ctx[d])) die("calc_xr (sync)");
      out = abft_cuda_scalar_host(ctx[0], slot[0]);
      x->dirty = r->dirty = true;
    }
    report_faults();
    return out;
  }
  virtual void calc_p(cg_vector *p, const cg_vector *r, double beta) {
    if (G == 1) {
      if (abft_cuda_calc_p(ctx[0], p->full[0], r->full[0], beta)) die("calc_p");
      return;
    }
    cg_vector *R = const_cast<cg_vector *>(r);
    prepare(p);
    prepare(R);
    // p = r + beta p and, in the same kernel, the new entries go to the GPUs whose rows reference them
    for (int d = 0; d < G; d++)
      if (abft_cuda_dist_calc_p(ctx[d], p->comm[d], p->loc[d], R->loc[d], beta)) die("calc_p");
    p->dirty = false;
  }
  virtual void spmv(const cg_matrix *mat, const cg_vector *vec, cg_vector *result) {
    if (events_pending()) report_faults();
    if (G == 1) {
      if (abft_cuda_spmv(ctx[0], mat->part[0], vec->full[0], result->full[0])) die("spmv");
      return;
    }
    cg_vector *V = const_cast<cg_vector *>(vec);
    prepare(V);
    prepare(result);
    if (V->dirty) {
      // refresh the peers' copies (outside the CG loop only: p is pushed by calc_p).  No reduction orders
      // this push against the peers' earlier reads of the vector, so drain every GPU first.
      for (int d = 0; d < G; d++)
        if (abft_cuda_ctx_sync(ctx[d])) die("spmv (sync)");
      for (int d = 0; d < G; d++)
        if (abft_cuda_dist_push(ctx[d], V->comm[d], V->loc[d])) die("spmv (exchange)");
      V->dirty = false;
    }
    for (int d = 0; d < G; d++)
      if (abft_cuda_dist_spmv(ctx[d], V->comm[d], mat->part[d], V->full[d], result->loc[d])) die("spmv");
    result->dirty = true;
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
    int owner = 0;
    if (G > 1)
      while (owner + 1 < G && index >= off[owner + 1]) owner++;
    for (int i = 0; i < num_flips; i++) {
      int bit = (rand() % (end - start)) + start;
      printf("*** flipping bit %d at index %d ***\n", bit, index);
      if (abft_cuda_flip_bit(ctx[owner], mat->part[owner], index, bit)) die("inject_bitflip");
    }
  }

 private:
  // views of the owned rows + one exchange window per GPU for this vector, peers attached
  void prepare(cg_vector *v) {
    if (v->prepared) return;
    if (!have_cuts) {
      fprintf(stderr, "\ncuda target: with ABFT_CUDA_NGPUS > 1 the matrix must be created before vectors are used\n\n");
      exit(1);
    }
    for (int d = 0; d < G; d++) {
      abft_vector *view = abft_cuda_vector_view(ctx[d], v->full[d], cut[d], cut[d + 1] - cut[d]);
      abft_comm *cm = abft_cuda_comm_create(ctx[d], d, G);
      if (!view || !cm) die("vector partition");
      v->loc.push_back(view);
      v->comm.push_back(cm);
    }
    for (int d = 0; d < G; d++) {
      if (abft_cuda_comm_attach_self(ctx[d], v->comm[d], v->full[d])) die("comm attach");
      int lo[8], hi[8];
      unsigned wait = 0;
      for (int s = 0; s < G; s++) {
        lo[s] = hi[s] = 0;
        if (s == d) continue;
        if (abft_cuda_comm_attach_peer_ptrs(ctx[d], v->comm[d], s, abft_cuda_comm_window_ptr(v->comm[s]),
                                            abft_cuda_vector_dev_ptr(v->full[s]), device[s]))
          die("comm attach peer");
        lo[s] = need_lo[s][d];  // rows of mine that GPU s reads
        hi[s] = need_hi[s][d];
        if (need_hi[d][s] > need_lo[d][s]) wait |= 1u << s;  // GPU s pushes to me
      }
      if (abft_cuda_comm_set_plan(ctx[d], v->comm[d], lo, hi, wait, cut[d])) die("comm plan");
    }
    v->prepared = true;
  }

  bool events_pending() {
    for (int d = 0; d < G; d++)
      if (abft_cuda_events_pending(ctx[d])) return true;
    return false;
  }

  // print queued fault events the way the reference does (GPUs in row order = the serial order of the
  // CPU loop), exit(1) on a fatal one
  void report_faults() {
    if (!events_pending()) return;
    for (int d = 0; d < G; d++) {
      abft_event ev[256];
      int fatal = 0, n;
      while ((n = abft_cuda_poll_events(ctx[d], ev, 256, &fatal)) > 0) {  // (handed out 256 at a time until drained)
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
      if (n < 0) die("poll_events");
    }
  }

  int G;
  std::vector<abft_ctx *> ctx;
  std::vector<int> device;
  bool have_cuts;
  std::vector<int> cut;         // row boundaries, G+1
  std::vector<long long> off;   // element boundaries, G+1
  std::vector<std::vector<int> > need_lo, need_hi;  // need[d][s]: rows of GPU s that GPU d's columns reference
};

// lower-case names, no '-' inside (run_tests splits `--list` output on '-')
static CGContext::Register<CUDAContext<ABFT_MODE_NONE> > A("cuda", "none");
static CGContext::Register<CUDAContext<ABFT_MODE_CONSTRAINTS> > B("cuda", "constraints");
static CGContext::Register<CUDAContext<ABFT_MODE_SED> > C("cuda", "sed");
static CGContext::Register<CUDAContext<ABFT_MODE_SEC7> > D("cuda", "sec7");
static CGContext::Register<CUDAContext<ABFT_MODE_SEC8> > E("cuda", "sec8");
static CGContext::Register<CUDAContext<ABFT_MODE_SECDED> > F("cuda", "secded");
static CGContext::Register<CUDAContext<ABFT_MODE_CRC32C> > H("cuda", "crc32c");
// secded + row pointers (SECDED per pair) + vectors (parity bit per element): what the reference leaves unprotected
static CGContext::Register<CUDAContext<ABFT_MODE_SECDEDV> > I("cuda", "secdedv");

}  // namespace
