/*
 * abft_cuda.h -- C ABI of the B200-native (sm_100a) CUDA target for the CSR CG
 * hot path of abft-sparse-cg.
 *
 * This is the drop-in boundary.  The reference's boundary is the C++ plugin
 * class CGContext (CGContext.h:8-67); its twelve pure virtuals map 1:1 onto
 * the entry points below, and abft-sparse-cg_b200/host/CUDAContext.cpp is the
 * thin CGContext subclass that forwards to them and registers
 * ("cuda", <mode>) in the reference's own registry (CGContext.h:55-63), so the
 * unmodified cg.cpp runs `-t cuda -m <mode>`.  INTEGRATION.md shows the
 * reference-side binding.
 *
 * Plain pointers and sizes only; no C++ or torch types.  All host pointers
 * are ordinary (pageable or pinned) host memory unless a name says `dev`.
 * Unless noted, functions return 0 on success and a non-zero CUDA error code
 * otherwise; abft_cuda_last_error() returns the text.  There is NO CPU
 * fallback: with no usable CUDA device every call fails.
 *
 * Threading: like the reference (single host thread, one context per
 * process, CGContext.h / cg.cpp:42-154) a context must be driven by one
 * thread at a time.  All work of a context is ordered on one CUDA stream.
 */
#ifndef ABFT_CUDA_H
#define ABFT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABFT_CUDA_ABI_VERSION 1

/* ABFT modes == the `-m` strings registered by CSR/CPUContext.cpp:413-421,
 * plus crc32c (defined by this repository). */
enum abft_mode {
  ABFT_MODE_NONE = 0,
  ABFT_MODE_CONSTRAINTS = 1,
  ABFT_MODE_SED = 2,
  ABFT_MODE_SEC7 = 3,
  ABFT_MODE_SEC8 = 4,
  ABFT_MODE_SECDED = 5,
  ABFT_MODE_CRC32C = 6,
  /* secded + what the reference leaves unprotected (CSR/CPUContext.cpp:54-113,123-124; repository-defined, DESIGN.md
   * 9c/9d): the row pointers carry a SECDED code per pair, and every element of every vector a parity bit in place
   * of its mantissa's least significant bit (detect-only: "[ECC] error detected in vector ..." + exit(1)). */
  ABFT_MODE_SECDEDV = 7,
  ABFT_MODE_COUNT = 8
};

/* One record per message the reference printf()s from inside spmv
 * (CSR/CPUContext.cpp:175,180,190,197,233,278,324,331,380,387,398). */
enum abft_event_kind {
  ABFT_EV_SED_DETECTED = 1,     /* "[ECC] error detected at index %d"              fatal */
  ABFT_EV_CORRECTED_BIT = 2,    /* "[ECC] corrected bit %u at index %d"                  */
  ABFT_EV_CORRECTED_PARITY = 3, /* "[ECC] corrected overall parity bit at index %d"      */
  ABFT_EV_DOUBLE_BIT = 4,       /* "[ECC] double-bit error detected"               fatal */
  ABFT_EV_ROW_SIZE = 5,         /* "row size constraint violated for row %d"       fatal */
  ABFT_EV_ROW_ORDER = 6,        /* "row order constraint violated for row%d"       fatal */
  ABFT_EV_COL_SIZE = 7,         /* "column size constraint violated at index %d"   fatal */
  ABFT_EV_COL_ORDER = 8,        /* "column order constraint violated at index %d"  fatal */
  ABFT_EV_CRC_ROW = 9,          /* "[CRC32C] error detected in row %d"             fatal */
  /* row-pointer protection (repository-defined: csrc/rowsecc.cuh, ABFT_CUDA_PROTECT_ROWS=secded); index = pair of row pointers */
  ABFT_EV_ROWPTR_CORRECTED = 10, /* "[ROWS] corrected bit %u of row pointer pair %d"        */
  ABFT_EV_ROWPTR_DOUBLE = 11,    /* "[ROWS] double-bit error detected in row pointer pair %d" fatal */
  /* more faults in one spmv than the device log holds (16384): the records past that were dropped, so the lines
   * printed can no longer equal the reference's -- reported last and fatal instead of silently; index = records dropped */
  ABFT_EV_LOG_OVERFLOW = 12,     /* "fault log overflow: %d events dropped"                 fatal */
  /* mode secdedv: odd parity in a vector element; index = element (the column for spmv's operand), bit = which
   * vector: 1 spmv operand, 2 x, 3 r, 4 p, 5 w, 6/7 dot operands, 8 a vector being mapped to the host */
  ABFT_EV_VECTOR_SED = 13        /* "[ECC] error detected in vector %s at index %d"         fatal */
};

typedef struct abft_event {
  int32_t kind;   /* enum abft_event_kind */
  uint32_t index; /* GLOBAL CSR element index (row number for ROW_* / CRC_ROW) */
  uint32_t bit;   /* element bit 0..95 for CORRECTED_BIT (0..63 value, 64..95 column) */
} abft_event;

typedef struct abft_ctx abft_ctx;       /* replaces a CGContext instance            */
typedef struct abft_matrix abft_matrix; /* replaces cg_matrix (CSR/CPUContext.h:11-18) */
typedef struct abft_vector abft_vector; /* replaces cg_vector (CSR/CPUContext.h:5-9)   */

/* ---- library ---- */
int abft_cuda_abi_version(void);
const char *abft_cuda_last_error(void);
const char *abft_cuda_mode_name(int mode);    /* "none", "constraints", ... or NULL */
int abft_cuda_mode_from_name(const char *s);  /* -1 if unknown */
int abft_cuda_device_count(void);             /* 0 when no CUDA device is usable */
/* reference message text of an event, no trailing newline; returns strlen */
int abft_cuda_format_event(const abft_event *e, char *buf, int buflen);
int abft_cuda_event_is_fatal(int kind);

/* ---- context: CGContext::create(target, mode) (CGContext.cpp:9-25) ---- */
abft_ctx *abft_cuda_ctx_create(int mode, int device); /* NULL on failure */
void abft_cuda_ctx_destroy(abft_ctx *ctx);
int abft_cuda_ctx_mode(const abft_ctx *ctx);
int abft_cuda_ctx_sync(abft_ctx *ctx); /* wait for everything queued on the context */
void *abft_cuda_ctx_stream(abft_ctx *ctx); /* the context's cudaStream_t */
/* Measurement aid (off by default; a null pointer in the kernels' arguments): per CG iteration, nanosecond
 * globaltimer stamps of where the time between the streams goes -- row i (of `rows`, a ring) holds ABFT_PHASE_SLOTS
 * 64-bit values: 0 spmv begins, 1 spmv has seen every peer's push of p, 2 spmv's last CTA enters the cross-rank
 * all-reduce of p.w, 3 p.w published, 4 calc_xr begins, 5 its last CTA enters the all-reduce of r.r, 6 r.r
 * published, 7 calc_p (+ push) begins, 8 every store to the peers has landed.  globaltimer is one clock for all GPUs
 * of a box, so rows of different ranks can be laid side by side.  rows = 0 switches it off again. */
#define ABFT_PHASE_SLOTS 12
int abft_cuda_phases_enable(abft_ctx *ctx, int rows);
int abft_cuda_phases_read(abft_ctx *ctx, uint64_t *out, int cap_rows); /* -> rows copied (oldest first up to `rows`) */

/* ---- matrix: create_matrix / destroy_matrix (CSR/CPUContext.cpp:11-52) ----
 * Triplets sorted by (row, col), 0-based (what cg.cpp:394-418 passes).  The
 * arrays are deep-copied (the caller frees them at once, cg.cpp:420-422); the
 * COO->CSR conversion and the per-element / per-row ECC encode
 * (generate_ecc_bits, CSR/CPUContext.cpp:209-212,247-250,291-295,347-351) run
 * on the device.  ECC modes need N <= 2^24 (column index is 24 bits,
 * CSR/CPUContext.cpp:238,282,338,404). */
abft_matrix *abft_cuda_matrix_create(abft_ctx *ctx, const uint32_t *columns, const uint32_t *rows,
                                     const double *values, int N, int nnz);
/* Row partition [row_offset, row_offset+n_rows) of an N_global-column matrix
 * (multi-GPU): `rows` are GLOBAL row numbers, `columns` stay
 * GLOBAL so the ECC words are bit-identical to the single-GPU encoding;
 * nnz_offset = global index of this partition's first element (used for the
 * `index` of reported events and for flip routing). */
abft_matrix *abft_cuda_matrix_create_part(abft_ctx *ctx, const uint32_t *columns, const uint32_t *rows,
                                          const double *values, int n_rows, int nnz, int N_global,
                                          int row_offset, int64_t nnz_offset);
void abft_cuda_matrix_destroy(abft_ctx *ctx, abft_matrix *mat);
/* Host-only: the spmv work decomposition ("tiles": runs of whole rows) create_matrix would build for
 * these CSR row pointers.  tile_nnz <= 0: the size create_matrix picks.  Writes up to cap records
 * {row_begin, nnz_begin, nnz_end, info} (info: bits 0..15 rows, bit 16 thread-per-row tile, bit 17 single row longer
 * than a tile, bits 18..29 longest row) and returns the number of tiles.  Needs no device. */
int abft_cuda_plan_tiles(const uint32_t *rowptr, int n_rows, int tile_nnz, uint32_t *out, int cap);
/* Host-only: the decomposition AND the kernel create_matrix picks for these row pointers in this mode:
 * *kind = 1 tiles drawn dynamically from a counter (matrices with irregular tiles), 0 tiles in static round-robin (every
 * stencil matrix).  Same records and return value as abft_cuda_plan_tiles. */
int abft_cuda_plan_matrix(const uint32_t *rowptr, int n_rows, int mode, uint32_t *out, int cap, int *kind);
int abft_cuda_matrix_rows(const abft_matrix *mat);
int abft_cuda_matrix_nnz(const abft_matrix *mat);
/* copy the stored (encoded) CSR arrays back; any pointer may be NULL.
 * cols_out[nnz], rowptr_out[n_rows+1], values_out[nnz].  Test/debug aid. */
int abft_cuda_matrix_read(abft_ctx *ctx, const abft_matrix *mat, uint32_t *cols_out, uint32_t *rowptr_out,
                          double *values_out);
/* overwrite one row pointer (fault injection into rows[] for the CONSTRAINTS checks) */
int abft_cuda_matrix_poke_rowptr(abft_ctx *ctx, abft_matrix *mat, int row, uint32_t value);
int abft_cuda_matrix_poke_col(abft_ctx *ctx, abft_matrix *mat, int index, uint32_t value);
int abft_cuda_matrix_poke_value(abft_ctx *ctx, abft_matrix *mat, int index, double value);

/* ---- vectors (CSR/CPUContext.cpp:54-80) ---- */
abft_vector *abft_cuda_vector_create(abft_ctx *ctx, int N); /* zero-filled */
/* non-owning vector over caller-managed DEVICE memory (e.g. a torch tensor) */
abft_vector *abft_cuda_vector_wrap(abft_ctx *ctx, double *dev_ptr, int N);
/* non-owning window [offset, offset+N) of another vector */
abft_vector *abft_cuda_vector_view(abft_ctx *ctx, abft_vector *parent, int offset, int N);
void abft_cuda_vector_destroy(abft_ctx *ctx, abft_vector *v);
/* Injection aid for mode secdedv (the reference injects into the matrix only): flip bit 0..63 of element `index`. */
int abft_cuda_vector_flip_bit(abft_ctx *ctx, abft_vector *vec, int index, int bit);
int abft_cuda_vector_len(const abft_vector *v);
double *abft_cuda_vector_dev_ptr(abft_vector *v);
/* map_vector: device->host copy into a pinned staging buffer owned by the
 * vector; the pointer stays valid until unmap.  Several vectors may be mapped
 * at once (cg.cpp:66-74).  NULL on failure. */
double *abft_cuda_vector_map(abft_ctx *ctx, abft_vector *v);
/* unmap_vector: host->device copy of all N entries from `h` (the pointer map
 * returned, or any host buffer of N doubles). */
int abft_cuda_vector_unmap(abft_ctx *ctx, abft_vector *v, const double *h);
int abft_cuda_vector_copy(abft_ctx *ctx, abft_vector *dst, const abft_vector *src);
/* The same two transfers for a caller that owns the host buffer -- the plugin with several GPUs under one host thread
 * keeps ONE host image of a vector and moves each GPU's slice of it: asynchronous device->host / host->device copy
 * of all N entries on the context's stream (mode secdedv checks / encodes like map / unmap);
 * abft_cuda_ctx_sync() completes it.  abft_cuda_host_alloc: pinned host memory usable with every device of the
 * process (so that the slices move concurrently and without a staging copy); NULL on failure. */
int abft_cuda_vector_download_async(abft_ctx *ctx, abft_vector *v, double *host_dst);
int abft_cuda_vector_upload_async(abft_ctx *ctx, abft_vector *v, const double *host_src);
void *abft_cuda_host_alloc(size_t bytes);
void abft_cuda_host_free(void *p);

/* ---- the hot path (cg.cpp:97-112) ---- */
/* spmv: result = A * vec, ABFT check/correct of the context's mode fused into
 * the element stream (CSR/CPUContext.cpp:115-133,162-207,214-245,252-289,
 * 297-345,353-411).  Asynchronous.  Faults found are queued and surface at the
 * next abft_cuda_poll_events().  Also leaves dot(vec, result) on the device so
 * the abft_cuda_dot(vec, result) that cg.cpp:100 issues next costs no pass. */
int abft_cuda_spmv(abft_ctx *ctx, abft_matrix *mat, const abft_vector *vec, abft_vector *result);
/* dot (CSR/CPUContext.cpp:82-90): blocks until the value is on the host */
int abft_cuda_dot(abft_ctx *ctx, const abft_vector *a, const abft_vector *b, double *out);
/* calc_xr (CSR/CPUContext.cpp:92-105): x += alpha p; r -= alpha w; *out = r.r */
int abft_cuda_calc_xr(abft_ctx *ctx, abft_vector *x, abft_vector *r, const abft_vector *p,
                      const abft_vector *w, double alpha, double *out);
/* calc_p (CSR/CPUContext.cpp:107-113): p = r + beta p.  Asynchronous. */
int abft_cuda_calc_p(abft_ctx *ctx, abft_vector *p, const abft_vector *r, double beta);

/* Asynchronous halves of dot / calc_xr for multi-GPU use: the partial result
 * stays on the device (and in a pinned host slot after abft_cuda_ctx_sync) so
 * the caller can all-reduce it.  slot in [0, 8). */
int abft_cuda_dot_async(abft_ctx *ctx, const abft_vector *a, const abft_vector *b, int slot);
int abft_cuda_calc_xr_async(abft_ctx *ctx, abft_vector *x, abft_vector *r, const abft_vector *p,
                            const abft_vector *w, double alpha, int slot);
double *abft_cuda_scalar_dev_ptr(abft_ctx *ctx, int slot);  /* device address of slot */
double abft_cuda_scalar_host(abft_ctx *ctx, int slot);       /* value after a sync */
#define ABFT_SLOT_SPMV_DOT 0  /* where spmv leaves dot(vec, result) */

/* ---- fault injection: the RMW of inject_bitflip (CSR/CPUContext.cpp:150-157).
 * `index` is a GLOBAL element index, bit 0..63 flips the value, 64..95 the
 * column word.  The caller draws rand() (one for the index, one per flip,
 * CSR/CPUContext.cpp:137,148) so that sequences match the reference. */
int abft_cuda_flip_bit(abft_ctx *ctx, abft_matrix *mat, int64_t index, int bit);

/* ---- fault reporting ----
 * Waits for queued work, then moves the pending events, sorted in the
 * reference's serial order (ascending row / element index), into out[0..cap).
 * Returns the number written (<= cap), or -1 on a CUDA error.  *fatal (may be
 * NULL) is set to 1 if one of them is fatal; events after the first fatal one
 * are dropped, as the reference would have exit(1)ed there. */
int abft_cuda_poll_events(abft_ctx *ctx, abft_event *out, int cap, int *fatal);
/* cheap host-side check (no sync): non-zero if events may be pending */
int abft_cuda_events_pending(abft_ctx *ctx);

/* ---- device-resident CG loop (cg.cpp:83-118 without the two host round trips
 * per iteration).  Runs copy r=b, p=r, rr=r.r and then iterations until
 * max_itrs or rr <= conv_threshold, scalars (alpha, beta, rr) staying on the
 * device; the convergence flag is read back every `check_every` iterations,
 * kernels of iterations past convergence are no-ops, so the iteration count
 * and the vectors equal those of the host-driven loop.  Returns iterations
 * run (>= 0), or -1 on error / -2 if a fatal fault event is pending. */
int abft_cuda_cg_solve(abft_ctx *ctx, abft_matrix *A, const abft_vector *b, abft_vector *x, abft_vector *r,
                       abft_vector *p, abft_vector *w, int max_itrs, double conv_threshold, int check_every,
                       double *rr_out);

/* ---- the plugin path as one call: the loop body of cg.cpp:94-118 over abft_cuda_spmv / _dot / _calc_xr / _calc_p,
 * scalars on the host -- what host/CUDAContext.cpp does for the unmodified driver, compiled, so that a caller in an
 * interpreted language (tests, bench.py) times the library and not its own call overhead.  Iterates until max_itrs
 * or rr <= conv_threshold; *rr_io carries r.r in and out; rr_hist (may be NULL, [max_itrs]) gets r.r per iteration.
 * Returns the iterations done (>= 0) or -1.  If a fault event is pending after an spmv, the loop stops BEFORE using
 * that result: poll the events and, unless one was fatal, call again (the interrupted iteration is redone). */
int abft_cuda_cg_host_loop(abft_ctx *ctx, abft_matrix *A, abft_vector *x, abft_vector *r, abft_vector *p,
                           abft_vector *w, int max_itrs, double conv_threshold, double *rr_io, double *rr_hist);

/* ---- multi-GPU: row-partitioned CG with the exchange fused into the kernels ----
 * One context per GPU (one process per GPU, or several contexts in one process).  Every rank owns a
 * row block of A (abft_cuda_matrix_create_part), the matching slices of x, r, w, b, and a FULL-length
 * copy of the search direction p (the "exchange vector", created with abft_cuda_vector_create).
 * Ranks never read remote memory: dist_calc_p computes p = r + beta p on its slice and, in the same
 * kernel, STORES the entries its peers' rows reference straight into their copies of p over NVLink
 * (peer memory: CUDA IPC between processes, peer access inside one), then raises one flag per peer;
 * dist_spmv spins on those flags before it touches p.  The two scalars per iteration are all-reduced
 * inside the reduction's last CTA the same way (each rank stores its partial into every peer's window
 * and adds the G partials in rank order, so all ranks get identical bits; no NCCL call on the path).
 * The plan (which rows each peer needs) comes from the partition's column ranges: a banded matrix
 * exchanges a halo, an irregular one degenerates to an all-gather.
 * All ranks must issue the same sequence of dist_* calls. */
#define ABFT_IPC_HANDLE_BYTES 64
typedef struct abft_comm abft_comm;
abft_comm *abft_cuda_comm_create(abft_ctx *ctx, int rank, int world); /* world <= 8 */
void abft_cuda_comm_destroy(abft_ctx *ctx, abft_comm *comm);
/* handles to send to the peers (any transport; bench.py uses torch.distributed) */
int abft_cuda_comm_window_handle(abft_ctx *ctx, abft_comm *comm, unsigned char *out64);
int abft_cuda_vector_ipc_handle(abft_ctx *ctx, abft_vector *vec, unsigned char *out64);
void *abft_cuda_comm_window_ptr(abft_comm *comm); /* same-process peers exchange raw pointers */
int abft_cuda_comm_attach_self(abft_ctx *ctx, abft_comm *comm, abft_vector *vec_full);
int abft_cuda_comm_attach_peer(abft_ctx *ctx, abft_comm *comm, int peer, const unsigned char *window_handle64,
                               const unsigned char *vec_handle64);
int abft_cuda_comm_attach_peer_ptrs(abft_ctx *ctx, abft_comm *comm, int peer, void *window_dev, double *vec_dev,
                                    int peer_device);
/* push_lo/hi[s]: GLOBAL rows [lo, hi) of this rank's slice that rank s reads (lo >= hi: none);
 * wait_mask: bit s set if rank s pushes to this rank; row0: this rank's first global row */
int abft_cuda_comm_set_plan(abft_ctx *ctx, abft_comm *comm, const int *push_lo, const int *push_hi,
                            unsigned int wait_mask, int row0);
int abft_cuda_dist_push(abft_ctx *ctx, abft_comm *comm, abft_vector *p_local); /* initial p = r exchange */
int abft_cuda_dist_calc_p(abft_ctx *ctx, abft_comm *comm, abft_vector *p_local, const abft_vector *r, double beta);
int abft_cuda_dist_spmv(abft_ctx *ctx, abft_comm *comm, abft_matrix *mat, const abft_vector *vec_full,
                        abft_vector *result);
int abft_cuda_dist_dot(abft_ctx *ctx, abft_comm *comm, const abft_vector *a, const abft_vector *b, double *out);
int abft_cuda_dist_calc_xr(abft_ctx *ctx, abft_comm *comm, abft_vector *x, abft_vector *r, const abft_vector *p,
                           const abft_vector *w, double alpha, double *out);
/* Asynchronous halves, for one host thread driving several GPUs (host/CUDAContext.cpp with ABFT_CUDA_NGPUS):
 * the reduction of rank r only completes once every rank has launched its part, so all launches must be issued
 * before the first sync.  *slot_out: read abft_cuda_scalar_host(ctx, slot) after abft_cuda_ctx_sync. */
int abft_cuda_dist_dot_async(abft_ctx *ctx, abft_comm *comm, const abft_vector *a, const abft_vector *b, int *slot_out);
int abft_cuda_dist_calc_xr_async(abft_ctx *ctx, abft_comm *comm, abft_vector *x, abft_vector *r, const abft_vector *p,
                                 const abft_vector *w, double alpha, int *slot_out);

/* device-resident loop over a row partition: like abft_cuda_cg_solve, every rank calls it with the same
 * arguments; scalars never visit the host, the exchange and the all-reduces stay inside the kernels.  A fatal
 * fault does not stop it early (the ranks wait on each other and must issue the same kernels); the rank that logged
 * it returns -2 once the loop has run out (poll its events), the others their iteration count. */
int abft_cuda_dist_cg_solve(abft_ctx *ctx, abft_comm *comm, abft_matrix *A, const abft_vector *b, abft_vector *x,
                            abft_vector *r, abft_vector *p_local, abft_vector *p_full, abft_vector *w, int max_itrs,
                            double conv_threshold, int check_every, double *rr_out);
/* abft_cuda_cg_host_loop over a row partition (the abft_cuda_dist_* calls above); every rank calls it in lock-step.
 * Faults do not interrupt it (the ranks wait on each other): poll the events afterwards. */
int abft_cuda_dist_cg_host_loop(abft_ctx *ctx, abft_comm *comm, abft_matrix *A, abft_vector *x, abft_vector *r,
                                abft_vector *p_local, abft_vector *p_full, abft_vector *w, int max_itrs,
                                double conv_threshold, double *rr_io, double *rr_hist);

/* ---- introspection for bench.py / tests ---- */
/* number of kernels this library has launched on the context since creation */
uint64_t abft_cuda_launch_count(const abft_ctx *ctx);
/* tiles of the spmv work decomposition */
int abft_cuda_matrix_tiles(const abft_matrix *mat);

/* ---------------------------------------------------------------------------
 * COO format -- the reference's second executable, cg-coo (COO/CPUContext.h:17-22 struct cg_matrix
 * {N, nnz, coo_element*}; COO/CPUContext.cpp).  Vectors, dot, calc_xr, calc_p and the fault log are the
 * ones above.  STATUS: parity green on a B200 against the pinned COO oracle (tests/test_gpu_coo.py); the default kernel
 * streams the elements through shared memory (csrc/coo_stream.cuh): 0.83 (secded) .. 1.0 (none) of the format's roofline.
 * ------------------------------------------------------------------------- */
typedef struct abft_coo_matrix abft_coo_matrix;

/* CPUContext::create_matrix + generate_ecc_bits (COO/CPUContext.cpp:11-35, 196-199, 230-233, 269-273, 320-324):
 * deep copy, one 128-bit element {col | check byte, row, value} per triplet in the caller's order.
 * Modes none .. secded (crc32c has no COO form). */
abft_coo_matrix *abft_cuda_coo_matrix_create(abft_ctx *ctx, const uint32_t *columns, const uint32_t *rows,
                                             const double *values, int N, int nnz);
void abft_cuda_coo_matrix_destroy(abft_ctx *ctx, abft_coo_matrix *mat);
int abft_cuda_coo_matrix_nnz(const abft_coo_matrix *mat);
/* copies the nnz 16-byte elements {uint32 col, uint32 row, double value} to the host (tests) */
int abft_cuda_coo_matrix_read(abft_ctx *ctx, const abft_coo_matrix *mat, void *elements_out);
/* CPUContext*::spmv (COO/CPUContext.cpp:104-121, 142-194, 201-228, 235-267, 275-318, 326-379):
 * result[col] += value * vec[row] in element order per output entry; faults go to the context's log
 * (abft_cuda_poll_events), formatted with abft_cuda_coo_format_event.  Like abft_cuda_spmv it leaves dot(vec, result)
 * behind, so the abft_cuda_dot(vec, result) that cg.cpp:100 issues next costs no pass. */
int abft_cuda_coo_spmv(abft_ctx *ctx, abft_coo_matrix *mat, const abft_vector *vec, abft_vector *result);
/* the RMW of CPUContext::inject_bitflip (COO/CPUContext.cpp:138): bit 0..127 of element `index`
 * (0..31 col, 32..63 row, 64..127 value); the plugin draws index and bits with rand() like :125,136 */
int abft_cuda_coo_flip_bit(abft_ctx *ctx, abft_coo_matrix *mat, int index, int bit);
/* the reference's COO message texts (its CONSTRAINTS wording differs from the CSR target's) */
int abft_cuda_coo_format_event(const abft_event *e, char *buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif /* ABFT_CUDA_H */
