// abft_cuda.cu -- the C ABI of include/abft_cuda.h: host orchestration of the
// sm_100a kernels in spmv.cuh / vecops.cuh / encode.cuh.  No CPU fallback: if no
// CUDA device is usable every entry point fails and says so.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "ecc.cuh"
#include "encode.cuh"
#include "spmv.cuh"
#include "coo.cuh"
#include "coo_stream.cuh"
#include "vecops.cuh"

using namespace abft;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
  g_last_error = buf;
  return code ? code : -1;
}

#define CK(call)                                                                    \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) return fail((int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKP(call)                                                    \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      fail((int)e_, cudaGetErrorString(e_), __FILE__, __LINE__);     \
      return nullptr;                                                \
    }                                                                \
  } while (0)
#define REQUIRE(cond, msg)                                  \
  do {                                                      \
    if (!(cond)) return fail(-1, msg, __FILE__, __LINE__);  \
  } while (0)

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
enum { SLOT_SPMV_DOT = ABFT_SLOT_SPMV_DOT, SLOT_DOT = 1, SLOT_XR = 2, SLOT_RR0 = 3, SLOT_RR1 = 4 };

struct abft_vector {
  double *d = nullptr;
  int N = 0;
  bool owns = false;
  double *h_stage = nullptr;  // pinned, allocated on first map
};

struct abft_matrix {
  MatView view{};
  uint32_t *d_rowptr = nullptr;
  Tile *d_tiles = nullptr;
  unsigned int *d_tile_counter = nullptr;  // dynamic spmv schedule (matrices with irregular tiles)
  int n_tiles = 0;
  int64_t nnz_offset = 0;
  int flen = 0;  // stencil matrix (static schedule, every tile regular): its longest row if that is 5 or 7, else 0
};

struct abft_ctx {
  int mode = 0;      // public mode (include/abft_cuda.h)
  int base = 0;      // the matrix code it implies (secdedv: secded)
  bool vsed = false; // mode secdedv: parity bit in every vector element, row pointers protected
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  DevLog *d_log = nullptr;
  unsigned int *h_flag = nullptr;   // mapped pinned
  double *d_scalars = nullptr;      // kNumSlots
  double *h_scalars = nullptr;      // mapped pinned mirror
  double *d_partials = nullptr;     // kMaxPartials
  unsigned int *d_counter = nullptr;
  int *d_loop = nullptr;            // [0],[1] = active flags (ring of 2), [2] = iterations done
  int *h_loop = nullptr;            // mapped pinned mirror: [0] = iterations, [1] = next-active
  uint64_t launches = 0;
  int spmv_ctas_per_sm = 0;  // resident CTAs of this mode's spmv kernel (set at first launch)
  int spmv_ctas_per_sm_dyn = 0;  // ... of its dynamic-schedule variant
  // fused dot(vec, result) left behind by the last spmv
  const double *dot_a = nullptr, *dot_b = nullptr;
  int dot_n = 0;
  uint64_t dot_epoch = 0;
  uint64_t epoch = 1;  // bumped by every call that writes vector data
  unsigned long long *d_phase = nullptr;  // phase stamps (abft_cuda_phases_enable): phase_rows x kPhaseSlots
  int phase_rows = 0;
  long long phase_iter = 0;               // spmv launches since the stamps were enabled
  unsigned long long *phase_cur = nullptr;  // the row the running iteration stamps into
  std::vector<abft_event> pending;  // drained from the device log, sorted, not yet handed to the caller
  size_t pending_pos = 0;
};

struct abft_comm {
  int rank = 0, world = 1;
  CommWindow *win[kMaxRanks] = {};     // device-visible pointers; win[rank] is local
  double *peer_vec[kMaxRanks] = {};    // peers' full-length exchange vectors
  void *ipc_opened[2 * kMaxRanks] = {};  // pointers obtained from cudaIpcOpenMemHandle
  int n_ipc = 0;
  int push_lo[kMaxRanks] = {}, push_hi[kMaxRanks] = {};
  unsigned int wait_mask = 0;
  int row0 = 0;
  unsigned int seq = 0;    // reductions issued so far (same on every rank)
  unsigned int epoch = 0;  // pushes issued so far
};

static const char *kModeNames[ABFT_MODE_COUNT] = {"none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c", "secdedv"};

static bool g_tables_ready[64] = {false};

static int ensure_tables(int device) {
  if (device >= 0 && device < 64 && g_tables_ready[device]) return 0;
  std::vector<uint32_t> tab(1024);
  for (int s = 0; s < 4; s++)
    for (uint32_t i = 0; i < 256; i++) tab[s * 256 + i] = crc32c_table_entry(i, s);
  CK(cudaMemcpyToSymbol(g_crc_tab, tab.data(), sizeof(uint32_t) * 1024));
  std::vector<uint32_t> tab5((size_t)kCrcTabWords);
  for (int m = 1; m <= kCrcMaps; m++)
    for (int j = 0; j < kCrcChunks; j++)
      for (uint32_t cc = 0; cc < 32; cc++) tab5[(size_t)((m - 1) * kCrcChunks + j) * 32 + cc] = crc5_entry(m, j, cc);
  CK(cudaMemcpyToSymbol(g_crc5_tab, tab5.data(), sizeof(uint32_t) * (size_t)kCrcTabWords));
  std::vector<uint32_t> tabe((size_t)kCrcElemChunks * 32);
  for (int i = 0; i < kCrcElemChunks; i++)
    for (uint32_t cc = 0; cc < 32; cc++) tabe[(size_t)i * 32 + cc] = crc_elem_entry(i, cc);
  CK(cudaMemcpyToSymbol(g_crc_elem_tab, tabe.data(), sizeof(uint32_t) * tabe.size()));
  if (device >= 0 && device < 64) g_tables_ready[device] = true;
  return 0;
}

static ReduceOut make_reduce(abft_ctx *c, int slot, abft_comm *comm = nullptr) {
  ReduceOut r{};
  if (comm && comm->world > 1) {
    for (int s = 0; s < comm->world; s++) r.peers.win[s] = comm->win[s];
    r.peers.world = comm->world;
    r.peers.rank = comm->rank;
    r.peers.seq = ++comm->seq;
  }
  r.partials = c->d_partials;
  r.counter = c->d_counter;
  r.dev = c->d_scalars + slot;
  r.host = c->h_scalars + slot;
  // (the host polls the mapped word itself: "no value yet" until the reduction's finalizer stores it)
  memcpy(c->h_scalars + slot, &kSlotEmpty, sizeof(double));
  r.phase = (slot == SLOT_DOT) ? nullptr : c->phase_cur;
  r.phase_slot = (slot == SLOT_SPMV_DOT) ? PH_SPMV_REDUCE : PH_XR_REDUCE;
  r.active = nullptr;
  r.iters = nullptr;
  r.threshold = 0.0;
  r.max_itrs = 0;
  r.bump_iters = 0;
  return r;
}

// a new CG iteration starts with its spmv: the phase stamps of this iteration go into the next row
static void phase_next_row(abft_ctx *c) {
  if (c->d_phase) c->phase_cur = c->d_phase + (size_t)(c->phase_iter++ % c->phase_rows) * kPhaseSlots;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int vec_grid(const abft_ctx *c, int N) {
  long long want = ((long long)(N / 2 + 1) + kBlock - 1) / kBlock;
  long long cap = (long long)c->num_sms * 8;
  return (int)std::max(1LL, std::min(want, cap));
}

// The value of the last reduction into `slot`, as soon as its kernel has published it: the finalizer stores the
// value into mapped pinned memory (which make_reduce set to "no value yet"), and the host polls that word instead of
// paying a stream synchronisation per scalar (two per CG iteration on the plugin path).  Falls back to the stream's
// status every few thousand polls, so a failed launch surfaces as an error instead of a hang.
static int wait_scalar(abft_ctx *c, int slot, double *out) {
  volatile unsigned long long *q = reinterpret_cast<volatile unsigned long long *>(c->h_scalars + slot);
  unsigned long long bits;
  for (unsigned int spin = 1; (bits = *q) == kSlotEmpty; spin++) {
    if ((spin & 0xFFFu) == 0) {
      const cudaError_t e = cudaStreamQuery(c->stream);
      if (e == cudaSuccess) {
        if ((bits = *q) == kSlotEmpty) return fail(-1, "reduction finished without publishing its value", __FILE__, __LINE__);
        break;
      }
      if (e != cudaErrorNotReady) return fail((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    }
  }
  memcpy(out, &bits, sizeof(double));
  return 0;
}

// ---------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------
extern "C" {

int abft_cuda_abi_version(void) { return ABFT_CUDA_ABI_VERSION; }
const char *abft_cuda_last_error(void) { return g_last_error.c_str(); }

const char *abft_cuda_mode_name(int mode) {
  return (mode >= 0 && mode < ABFT_MODE_COUNT) ? kModeNames[mode] : nullptr;
}

int abft_cuda_mode_from_name(const char *s) {
  if (!s) return -1;
  for (int i = 0; i < ABFT_MODE_COUNT; i++)
    if (!strcmp(s, kModeNames[i])) return i;
  return -1;
}

int abft_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int abft_cuda_event_is_fatal(int kind) {
  return !(kind == ABFT_EV_CORRECTED_BIT || kind == ABFT_EV_CORRECTED_PARITY || kind == ABFT_EV_ROWPTR_CORRECTED);
}

int abft_cuda_format_event(const abft_event *e, char *buf, int buflen) {
  switch (e->kind) {
    case ABFT_EV_SED_DETECTED: return snprintf(buf, buflen, "[ECC] error detected at index %d", (int)e->index);
    case ABFT_EV_CORRECTED_BIT: return snprintf(buf, buflen, "[ECC] corrected bit %u at index %d", e->bit, (int)e->index);
    case ABFT_EV_CORRECTED_PARITY: return snprintf(buf, buflen, "[ECC] corrected overall parity bit at index %d", (int)e->index);
    case ABFT_EV_DOUBLE_BIT: return snprintf(buf, buflen, "[ECC] double-bit error detected");
    case ABFT_EV_ROW_SIZE: return snprintf(buf, buflen, "row size constraint violated for row %d", (int)e->index);
    case ABFT_EV_ROW_ORDER: return snprintf(buf, buflen, "row order constraint violated for row%d", (int)e->index);
    case ABFT_EV_COL_SIZE: return snprintf(buf, buflen, "column size constraint violated at index %d", (int)e->index);
    case ABFT_EV_COL_ORDER: return snprintf(buf, buflen, "column order constraint violated at index %d", (int)e->index);
    case ABFT_EV_CRC_ROW: return snprintf(buf, buflen, "[CRC32C] error detected in row %d", (int)e->index);
    case ABFT_EV_ROWPTR_CORRECTED: return snprintf(buf, buflen, "[ROWS] corrected bit %u of row pointer pair %d", e->bit, (int)e->index);
    case ABFT_EV_ROWPTR_DOUBLE: return snprintf(buf, buflen, "[ROWS] double-bit error detected in row pointer pair %d", (int)e->index);
    case ABFT_EV_VECTOR_SED: {
      static const char *const names[] = {"?", "spmv operand", "x", "r", "p", "w", "dot a", "dot b", "mapped"};
      return snprintf(buf, buflen, "[ECC] error detected in vector (%s) at index %d", names[e->bit < 9 ? e->bit : 0], (int)e->index);
    }
    case ABFT_EV_LOG_OVERFLOW:
      return e->bit ? snprintf(buf, buflen, "stray list overflow: %d rerouted products dropped", (int)e->index)
                    : snprintf(buf, buflen, "fault log overflow: %d events dropped", (int)e->index);
    default: return snprintf(buf, buflen, "unknown event %d", (int)e->kind);
  }
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
abft_ctx *abft_cuda_ctx_create(int mode, int device) {
  if (mode < 0 || mode >= ABFT_MODE_COUNT) {
    fail(-1, "unknown ABFT mode", __FILE__, __LINE__);
    return nullptr;
  }
  int ndev = abft_cuda_device_count();
  if (ndev <= 0) {
    fail(-1, "no CUDA device available (this target has no CPU fallback)", __FILE__, __LINE__);
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    fail(-1, "device index out of range", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(device));
  abft_ctx *c = new abft_ctx();
  // every early `return nullptr` below releases what was allocated so far
  struct Cleanup {
    abft_ctx *c;
    bool armed = true;
    ~Cleanup() {
      if (armed) abft_cuda_ctx_destroy(c);
    }
  } cleanup{c};
  c->mode = mode;
  c->vsed = mode == ABFT_MODE_SECDEDV;
  c->base = c->vsed ? ABFT_MODE_SECDED : mode;
  c->device = device;
  cudaDeviceProp prop;
  CKP(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  CKP(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CKP(cudaMalloc(&c->d_log, sizeof(DevLog)));
  CKP(cudaMemset(c->d_log, 0, sizeof(DevLog)));
  CKP(cudaHostAlloc(&c->h_flag, sizeof(unsigned int) * 16, cudaHostAllocMapped));
  memset(c->h_flag, 0, sizeof(unsigned int) * 16);
  CKP(cudaMalloc(&c->d_scalars, sizeof(double) * kNumSlots));
  CKP(cudaMemset(c->d_scalars, 0, sizeof(double) * kNumSlots));
  CKP(cudaHostAlloc(&c->h_scalars, sizeof(double) * kNumSlots, cudaHostAllocMapped));
  memset(c->h_scalars, 0, sizeof(double) * kNumSlots);
  CKP(cudaMalloc(&c->d_partials, sizeof(double) * kMaxPartials));
  CKP(cudaMalloc(&c->d_counter, sizeof(unsigned int) * 4));
  CKP(cudaMemset(c->d_counter, 0, sizeof(unsigned int) * 4));
  CKP(cudaMalloc(&c->d_loop, sizeof(int) * 4));
  CKP(cudaMemset(c->d_loop, 0, sizeof(int) * 4));
  CKP(cudaHostAlloc(&c->h_loop, sizeof(int) * 4, cudaHostAllocMapped));
  memset(c->h_loop, 0, sizeof(int) * 4);
  if (ensure_tables(device) != 0) return nullptr;
  CKP(cudaDeviceSynchronize());
  cleanup.armed = false;
  return c;
}

void abft_cuda_ctx_destroy(abft_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_log);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  cudaFree(c->d_scalars);
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  cudaFree(c->d_partials);
  cudaFree(c->d_counter);
  cudaFree(c->d_loop);
  cudaFree(c->d_phase);
  if (c->h_loop) cudaFreeHost(c->h_loop);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int abft_cuda_ctx_mode(const abft_ctx *c) { return c ? c->mode : -1; }

int abft_cuda_ctx_sync(abft_ctx *c) {
  REQUIRE(c, "null context");
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

void *abft_cuda_ctx_stream(abft_ctx *c) { return c ? (void *)c->stream : nullptr; }

int abft_cuda_phases_enable(abft_ctx *c, int rows) {
  REQUIRE(c && rows >= 0, "bad arguments to phases_enable");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_phase);
  c->d_phase = nullptr;
  c->phase_cur = nullptr;
  c->phase_rows = rows;
  c->phase_iter = 0;
  if (rows > 0) {
    CK(cudaMalloc(&c->d_phase, sizeof(unsigned long long) * (size_t)rows * kPhaseSlots));
    CK(cudaMemset(c->d_phase, 0, sizeof(unsigned long long) * (size_t)rows * kPhaseSlots));
  }
  return 0;
}

int abft_cuda_phases_read(abft_ctx *c, uint64_t *out, int cap_rows) {
  REQUIRE(c && out && cap_rows >= 0, "bad arguments to phases_read");
  if (!c->d_phase) return 0;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const int n = (int)std::min<long long>(std::min<long long>(c->phase_iter, c->phase_rows), cap_rows);
  if (n > 0) CK(cudaMemcpy(out, c->d_phase, sizeof(uint64_t) * (size_t)n * kPhaseSlots, cudaMemcpyDeviceToHost));
  return n;
}
uint64_t abft_cuda_launch_count(const abft_ctx *c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------
// matrix
// ---------------------------------------------------------------------------
static void build_tiles(const uint32_t *rowptr, uint32_t n_rows, uint32_t tile_nnz, uint32_t tile_rows, std::vector<Tile> &tiles) {
  tiles.clear();
  uint32_t r = 0;
  while (r < n_rows) {
    const uint32_t start = rowptr[r];
    uint32_t e = r;
    uint32_t max_len = 0;
    while (e < n_rows && (e - r) < tile_rows && rowptr[e + 1] >= start &&
           rowptr[e + 1] - start <= tile_nnz) {
      max_len = std::max(max_len, rowptr[e + 1] - rowptr[e]);
      e++;
    }
    uint32_t info;
    if (e == r) {
      e = r + 1;  // one row longer than a tile: streamed from global memory in chunks
      info = 1u | kTileLongRow;
    } else {
      const uint32_t nrows = e - r, nnz = rowptr[e] - start;
      info = nrows;
      // thread-per-row pays max_len steps (a gather round trip every 4), element-parallel pays nnz/kBlock steps
      // (+ the row phase): rows of similar length go thread-per-row (every stencil matrix) -- unless the tile is
      // a handful of very long rows (3 rows of 680: one thread would walk 680 elements while 253 idle; measured
      // 200k+ cycles for such a tile against 16k for the element-parallel path)
      const uint32_t mean_up = (nnz + nrows - 1) / nrows;
      if (max_len <= 2 * mean_up + 4 && (uint64_t)max_len * 32u <= (uint64_t)nnz + 512u) info |= kTileRegular;
      info |= std::min(max_len, 0xFFFu) << kTileMaxLenShift;
    }
    tiles.push_back(Tile{r, start, rowptr[e], info});
    r = e;
  }
}

// Which spmv kernel a CSR structure gets, and its tile list (host logic, no device):
//   kind 1  some tile is irregular (power-law rows) and the mode is not crc32c: the dynamic-schedule kernel;
//   kind 0  otherwise: the static round-robin kernel (every stencil matrix).
static int plan_matrix(const uint32_t *rowptr, uint32_t n_rows, int mode, std::vector<Tile> &tiles) {
  build_tiles(rowptr, n_rows, (uint32_t)kTileNnz, (uint32_t)kSpmvBlock, tiles);
  bool irregular = false;
  for (const Tile &t : tiles) irregular |= !(t.info & kTileRegular);
  // irregular tiles cost very different amounts (and the gathers make even equal tiles take 0.3x .. 2x the mean):
  // such a matrix is scheduled dynamically, CTAs drawing (smaller: spmv.cuh, kTileNnzDyn) tiles from a counter
  if (irregular && mode != ABFT_MODE_CRC32C) {
    build_tiles(rowptr, n_rows, (uint32_t)kTileNnzDyn, (uint32_t)kSpmvBlockDyn, tiles);
    return 1;
  }
  return 0;
}

int abft_cuda_plan_matrix(const uint32_t *rowptr, int n_rows, int mode, uint32_t *out, int cap, int *kind) {
  if (!rowptr || n_rows < 0 || mode < 0 || mode >= ABFT_MODE_COUNT) return fail(-1, "bad arguments to plan_matrix", __FILE__, __LINE__);
  std::vector<Tile> tiles;
  const int k = plan_matrix(rowptr, (uint32_t)n_rows, mode, tiles);
  if (kind) *kind = k;
  for (size_t i = 0; i < tiles.size() && (int)i < cap && out; i++) {
    out[4 * i + 0] = tiles[i].row_begin;
    out[4 * i + 1] = tiles[i].nnz_begin;
    out[4 * i + 2] = tiles[i].nnz_end;
    out[4 * i + 3] = tiles[i].info;
  }
  return (int)tiles.size();
}

// The spmv work decomposition of a CSR structure, without a device: what create_matrix would build for these row
// pointers (host logic only; used by the CPU tests and by integrators who want to see how a matrix will be cut).
// tile_nnz <= 0: the size create_matrix picks (8 elements per row of a full tile, or 1536 if that cut has irregular tiles).  Writes up to `cap`
// records {row_begin, nnz_begin, nnz_end, info} and returns the number of tiles (which may exceed cap).
int abft_cuda_plan_tiles(const uint32_t *rowptr, int n_rows, int tile_nnz, uint32_t *out, int cap) {
  if (!rowptr || n_rows < 0) return fail(-1, "bad arguments to plan_tiles", __FILE__, __LINE__);
  std::vector<Tile> tiles;
  if (tile_nnz > 0) {
    // (an explicit size: the static kernel's tiles unless it is the dynamic schedule's size)
    build_tiles(rowptr, (uint32_t)n_rows, (uint32_t)tile_nnz,
                (uint32_t)(tile_nnz == kTileNnzDyn && kTileNnzDyn != kTileNnz ? kSpmvBlockDyn : kSpmvBlock), tiles);
  } else {
    build_tiles(rowptr, (uint32_t)n_rows, (uint32_t)kTileNnz, (uint32_t)kSpmvBlock, tiles);
    bool irregular = false;
    for (const Tile &t : tiles) irregular |= !(t.info & kTileRegular);
    if (irregular) build_tiles(rowptr, (uint32_t)n_rows, (uint32_t)kTileNnzDyn, (uint32_t)kSpmvBlockDyn, tiles);
  }
  for (size_t i = 0; i < tiles.size() && (int)i < cap && out; i++) {
    out[4 * i + 0] = tiles[i].row_begin;
    out[4 * i + 1] = tiles[i].nnz_begin;
    out[4 * i + 2] = tiles[i].nnz_end;
    out[4 * i + 3] = tiles[i].info;
  }
  return (int)tiles.size();
}

abft_matrix *abft_cuda_matrix_create_part(abft_ctx *c, const uint32_t *columns, const uint32_t *rows,
                                          const double *values, int n_rows, int nnz, int N_global, int row_offset,
                                          int64_t nnz_offset) {
  if (!c || n_rows < 0 || nnz < 0 || N_global < 0 || (nnz > 0 && (!columns || !rows || !values))) {
    fail(-1, "bad arguments to matrix_create", __FILE__, __LINE__);
    return nullptr;
  }
  if (c->base >= ABFT_MODE_SED && (unsigned)N_global > (1u << 24)) {
    fail(-1, "ECC modes keep the check bits in column bits 24..31: N must be <= 2^24", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(c->device));
  abft_matrix *m = new abft_matrix();
  const size_t nz = (size_t)std::max(nnz, 1);
  uint32_t *d_coo_rows = nullptr;
  // every early `return nullptr` below releases what was allocated so far
  struct Cleanup {
    abft_matrix *&m;
    uint32_t *&tmp;
    bool armed = true;
    ~Cleanup() {
      if (!armed) return;
      cudaFree(tmp);
      cudaFree(m->view.cols);
      cudaFree(m->view.values);
      cudaFree(m->d_rowptr);
      cudaFree(m->d_tiles);
      cudaFree(m->d_tile_counter);
      delete m;
    }
  } cleanup{m, d_coo_rows};
  // +16 elements of padding: the TMA bulk copies of a tile round its window out to 16-byte granules
  CKP(cudaMalloc(&m->view.cols, sizeof(uint32_t) * (nz + 16)));
  CKP(cudaMalloc(&m->view.values, sizeof(double) * (nz + 16)));
  CKP(cudaMalloc(&m->d_rowptr, sizeof(uint32_t) * ((size_t)n_rows + 1 + 16)));
  CKP(cudaMalloc(&d_coo_rows, sizeof(uint32_t) * nz));
  CKP(cudaMemsetAsync(m->view.cols + nz, 0, sizeof(uint32_t) * 16, c->stream));
  CKP(cudaMemsetAsync(m->view.values + nz, 0, sizeof(double) * 16, c->stream));
  CKP(cudaMemsetAsync(m->d_rowptr + (size_t)n_rows + 1, 0, sizeof(uint32_t) * 16, c->stream));
  if (nnz > 0) {
    CKP(cudaMemcpyAsync(m->view.cols, columns, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    CKP(cudaMemcpyAsync(d_coo_rows, rows, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    CKP(cudaMemcpyAsync(m->view.values, values, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  const int grid = c->num_sms * 8;
  rowptr_kernel<<<grid, kBlock, 0, c->stream>>>(d_coo_rows, (uint32_t)nnz, (uint32_t)n_rows, (uint32_t)row_offset,
                                                m->d_rowptr);
  c->launches++;
  if (c->base >= ABFT_MODE_SED && c->base <= ABFT_MODE_SECDED && nnz > 0) {
    encode_elements_kernel<<<grid, kBlock, 0, c->stream>>>(c->base, m->view.cols, m->view.values, (uint32_t)nnz);
    c->launches++;
  } else if (c->base == ABFT_MODE_CRC32C && n_rows > 0) {
    encode_crc_rows_kernel<<<grid, kBlock, 0, c->stream>>>(m->d_rowptr, m->view.cols, m->view.values, (uint32_t)n_rows);
    c->launches++;
  }
  CKP(cudaGetLastError());
  bool protect_rows = false;
  {
    // mode secdedv protects the row pointers; ABFT_CUDA_PROTECT_ROWS=secded asks for it in any other mode
    const char *pr = c->vsed ? "secded" : getenv("ABFT_CUDA_PROTECT_ROWS");
    if (pr && *pr && strcmp(pr, "0") && strcmp(pr, "none")) {
      if (strcmp(pr, "secded") || nnz >= (1 << 28) || (row_offset & 1)) {
        fail(-1, "row-pointer protection (mode secdedv, ABFT_CUDA_PROTECT_ROWS): only `secded` exists; it needs nnz < 2^28 and partitions that start on even rows",
             __FILE__, __LINE__);
        return nullptr;
      }
      protect_rows = true;
    }
  }
  // tiles: greedy over the row pointers (host, O(rows))
  std::vector<uint32_t> h_rowptr((size_t)n_rows + 1);
  CKP(cudaMemcpyAsync(h_rowptr.data(), m->d_rowptr, sizeof(uint32_t) * ((size_t)n_rows + 1), cudaMemcpyDeviceToHost,
                      c->stream));
  if (protect_rows) {
    // (after the plain pointers went to the host for the tile plan)
    const uint32_t n_words = ((uint32_t)n_rows + 1u + 1u) & ~1u;
    encode_rowptr_pairs_kernel<<<grid, kBlock, 0, c->stream>>>(m->d_rowptr, n_words);
    c->launches++;
    CKP(cudaGetLastError());
    m->view.rows_ecc = 1u;
  }
  CKP(cudaStreamSynchronize(c->stream));
  CKP(cudaFree(d_coo_rows));
  d_coo_rows = nullptr;
  std::vector<Tile> tiles;
  const int kind = plan_matrix(h_rowptr.data(), (uint32_t)n_rows, c->base, tiles);
  m->n_tiles = (int)tiles.size();
  if (m->n_tiles > kMaxPartials) {   // (one reduction partial per tile or CTA)
    fail(-1, "matrix needs more spmv tiles than the reduction buffer holds", __FILE__, __LINE__);
    return nullptr;
  }
  if (kind == 1) {
    CKP(cudaMalloc(&m->d_tile_counter, sizeof(unsigned int)));
    CKP(cudaMemset(m->d_tile_counter, 0, sizeof(unsigned int)));
  }
  CKP(cudaMalloc(&m->d_tiles, sizeof(Tile) * (size_t)std::max(m->n_tiles, 1)));
  if (m->n_tiles)
    CKP(cudaMemcpy(m->d_tiles, tiles.data(), sizeof(Tile) * (size_t)m->n_tiles, cudaMemcpyHostToDevice));
  if (kind == 0 && m->n_tiles > 0) {
    // a stencil matrix?  (the stencil instantiations of spmv_kernel: 5- and 7-point rows, what BASELINE's configs are)
    uint32_t longest = 0;
    bool all_regular = true;
    for (const Tile &t : tiles) {
      all_regular &= (t.info & kTileRegular) != 0;
      longest = std::max(longest, (t.info >> kTileMaxLenShift) & 0xFFFu);
    }
    const char *force = getenv("ABFT_CUDA_SPMV_KERNEL");   // test hook: "generic" keeps the general instantiation
    if (all_regular && (longest == 5 || longest == 7) && !(force && !strcmp(force, "generic"))) m->flen = (int)longest;
  }
  m->view.rowptr = m->d_rowptr;
  m->view.tiles = m->d_tiles;
  m->view.tile_counter = m->d_tile_counter;
  m->view.n_tiles = (uint32_t)m->n_tiles;
  m->view.n_rows = (uint32_t)n_rows;
  m->view.nnz = (uint32_t)nnz;
  m->view.n_cols = (uint32_t)N_global;
  m->view.row_offset = (uint32_t)row_offset;
  m->view.nnz_offset = (uint32_t)nnz_offset;
  m->nnz_offset = nnz_offset;
  cleanup.armed = false;
  return m;
}

abft_matrix *abft_cuda_matrix_create(abft_ctx *c, const uint32_t *columns, const uint32_t *rows, const double *values,
                                     int N, int nnz) {
  return abft_cuda_matrix_create_part(c, columns, rows, values, N, nnz, N, 0, 0);
}

void abft_cuda_matrix_destroy(abft_ctx *c, abft_matrix *m) {
  if (!m) return;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  cudaFree(m->view.cols);
  cudaFree(m->view.values);
  cudaFree(m->d_rowptr);
  cudaFree(m->d_tiles);
  cudaFree(m->d_tile_counter);
  delete m;
}

int abft_cuda_matrix_rows(const abft_matrix *m) { return m ? (int)m->view.n_rows : -1; }
int abft_cuda_matrix_nnz(const abft_matrix *m) { return m ? (int)m->view.nnz : -1; }
int abft_cuda_matrix_tiles(const abft_matrix *m) { return m ? m->n_tiles : -1; }

int abft_cuda_matrix_read(abft_ctx *c, const abft_matrix *m, uint32_t *cols_out, uint32_t *rowptr_out,
                          double *values_out) {
  REQUIRE(c && m, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (cols_out && m->view.nnz) CK(cudaMemcpy(cols_out, m->view.cols, sizeof(uint32_t) * m->view.nnz, cudaMemcpyDeviceToHost));
  if (values_out && m->view.nnz) CK(cudaMemcpy(values_out, m->view.values, sizeof(double) * m->view.nnz, cudaMemcpyDeviceToHost));
  if (rowptr_out) CK(cudaMemcpy(rowptr_out, m->d_rowptr, sizeof(uint32_t) * ((size_t)m->view.n_rows + 1), cudaMemcpyDeviceToHost));
  return 0;
}

int abft_cuda_matrix_poke_rowptr(abft_ctx *c, abft_matrix *m, int row, uint32_t value) {
  REQUIRE(c && m, "null argument");
  REQUIRE(row >= 0 && (uint32_t)row <= m->view.n_rows, "row out of range");
  CK(cudaSetDevice(c->device));
  poke_u32_kernel<<<1, 1, 0, c->stream>>>(m->d_rowptr + row, value);
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_matrix_poke_col(abft_ctx *c, abft_matrix *m, int index, uint32_t value) {
  REQUIRE(c && m, "null argument");
  REQUIRE(index >= 0 && (uint32_t)index < m->view.nnz, "index out of range");
  CK(cudaSetDevice(c->device));
  poke_u32_kernel<<<1, 1, 0, c->stream>>>(m->view.cols + index, value);
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_matrix_poke_value(abft_ctx *c, abft_matrix *m, int index, double value) {
  REQUIRE(c && m, "null argument");
  REQUIRE(index >= 0 && (uint32_t)index < m->view.nnz, "index out of range");
  CK(cudaSetDevice(c->device));
  // two 4-byte stores through the same tiny kernel (keeps everything on the context's stream)
  uint32_t w[2];
  memcpy(w, &value, 8);
  uint32_t *words = reinterpret_cast<uint32_t *>(m->view.values + index);
  poke_u32_kernel<<<1, 1, 0, c->stream>>>(words, w[0]);
  poke_u32_kernel<<<1, 1, 0, c->stream>>>(words + 1, w[1]);
  c->launches += 2;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_flip_bit(abft_ctx *c, abft_matrix *m, int64_t index, int bit) {
  REQUIRE(c && m, "null argument");
  REQUIRE(bit >= 0 && bit < 96, "bit out of range");
  const int64_t local = index - m->nnz_offset;
  REQUIRE(local >= 0 && local < (int64_t)m->view.nnz, "element index not in this partition");
  CK(cudaSetDevice(c->device));
  flip_bit_kernel<<<1, 32, 0, c->stream>>>(m->view.cols, m->view.values, (uint32_t)local, bit);
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------
abft_vector *abft_cuda_vector_create(abft_ctx *c, int N) {
  if (!c || N < 0) {
    fail(-1, "bad arguments to vector_create", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(c->device));
  abft_vector *v = new abft_vector();
  v->N = N;
  v->owns = true;
  cudaError_t e = cudaMalloc(&v->d, sizeof(double) * (size_t)std::max(N, 1));
  if (e == cudaSuccess) e = cudaMemsetAsync(v->d, 0, sizeof(double) * (size_t)std::max(N, 1), c->stream);
  if (e != cudaSuccess) {
    fail((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    cudaFree(v->d);
    delete v;
    return nullptr;
  }
  return v;
}

abft_vector *abft_cuda_vector_wrap(abft_ctx *c, double *dev_ptr, int N) {
  if (!c || N < 0 || (!dev_ptr && N > 0)) {
    fail(-1, "bad arguments to vector_wrap", __FILE__, __LINE__);
    return nullptr;
  }
  abft_vector *v = new abft_vector();
  v->N = N;
  v->d = dev_ptr;
  v->owns = false;
  return v;
}

abft_vector *abft_cuda_vector_view(abft_ctx *c, abft_vector *parent, int offset, int N) {
  if (!c || !parent || offset < 0 || N < 0 || offset + N > parent->N) {
    fail(-1, "bad arguments to vector_view", __FILE__, __LINE__);
    return nullptr;
  }
  return abft_cuda_vector_wrap(c, parent->d + offset, N);
}

void abft_cuda_vector_destroy(abft_ctx *c, abft_vector *v) {
  if (!v) return;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->dot_a == v->d || c->dot_b == v->d) c->dot_a = c->dot_b = nullptr;
  }
  if (v->owns) cudaFree(v->d);
  if (v->h_stage) cudaFreeHost(v->h_stage);
  delete v;
}

// fault injection into a vector (the reference injects into the matrix only, CSR/CPUContext.cpp:135-159): flips one bit
// of element `index` on the device, through the same tiny kernel as the matrix poke
int abft_cuda_vector_flip_bit(abft_ctx *c, abft_vector *v, int index, int bit) {
  REQUIRE(c && v, "null argument");
  REQUIRE(index >= 0 && index < v->N && bit >= 0 && bit < 64, "vector_flip_bit: index or bit out of range");
  CK(cudaSetDevice(c->device));
  flip_word_bit_kernel<<<1, 1, 0, c->stream>>>(reinterpret_cast<uint32_t *>(v->d + index) + bit / 32, bit % 32);
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_vector_len(const abft_vector *v) { return v ? v->N : -1; }
double *abft_cuda_vector_dev_ptr(abft_vector *v) { return v ? v->d : nullptr; }

double *abft_cuda_vector_map(abft_ctx *c, abft_vector *v) {
  if (!c || !v) {
    fail(-1, "null argument", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(c->device));
  if (!v->h_stage) CKP(cudaHostAlloc(&v->h_stage, sizeof(double) * (size_t)std::max(v->N, 1), cudaHostAllocDefault));
  if (c->vsed && v->N) {
    // mode secdedv: the vector is checked on its way out (the host gets the stored values, parity bit and all)
    vec_check_kernel<<<vec_grid(c, v->N), kBlock, 0, c->stream>>>(v->d, v->N, LogRef{c->d_log, c->h_flag});
    c->launches++;
  }
  if (v->N) CKP(cudaMemcpyAsync(v->h_stage, v->d, sizeof(double) * (size_t)v->N, cudaMemcpyDeviceToHost, c->stream));
  CKP(cudaStreamSynchronize(c->stream));
  return v->h_stage;
}

int abft_cuda_vector_unmap(abft_ctx *c, abft_vector *v, const double *h) {
  REQUIRE(c && v && (h || v->N == 0), "null argument");
  CK(cudaSetDevice(c->device));
  if (v->N) CK(cudaMemcpyAsync(v->d, h, sizeof(double) * (size_t)v->N, cudaMemcpyHostToDevice, c->stream));
  if (c->vsed && v->N) {
    // mode secdedv: every element gets its parity bit (in place of the mantissa's lowest bit) as it arrives
    vec_encode_kernel<<<vec_grid(c, v->N), kBlock, 0, c->stream>>>(v->d, v->N);
    c->launches++;
  }
  // pageable sources are staged by the driver before the call returns; a pinned
  // source (our own staging buffer) must not be reused before the copy ran
  if (h != v->h_stage) CK(cudaStreamSynchronize(c->stream));
  c->epoch++;
  return 0;
}

int abft_cuda_vector_download_async(abft_ctx *c, abft_vector *v, double *host_dst) {
  REQUIRE(c && v && (host_dst || v->N == 0), "null argument");
  CK(cudaSetDevice(c->device));
  if (c->vsed && v->N) {
    vec_check_kernel<<<vec_grid(c, v->N), kBlock, 0, c->stream>>>(v->d, v->N, LogRef{c->d_log, c->h_flag});
    c->launches++;
  }
  if (v->N) CK(cudaMemcpyAsync(host_dst, v->d, sizeof(double) * (size_t)v->N, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

int abft_cuda_vector_upload_async(abft_ctx *c, abft_vector *v, const double *host_src) {
  REQUIRE(c && v && (host_src || v->N == 0), "null argument");
  CK(cudaSetDevice(c->device));
  if (v->N) CK(cudaMemcpyAsync(v->d, host_src, sizeof(double) * (size_t)v->N, cudaMemcpyHostToDevice, c->stream));
  if (c->vsed && v->N) {
    vec_encode_kernel<<<vec_grid(c, v->N), kBlock, 0, c->stream>>>(v->d, v->N);
    c->launches++;
  }
  c->epoch++;
  return 0;
}

void *abft_cuda_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    fail(-1, "host_alloc: cudaHostAlloc failed", __FILE__, __LINE__);
    return nullptr;
  }
  return p;
}

void abft_cuda_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int abft_cuda_vector_copy(abft_ctx *c, abft_vector *dst, const abft_vector *src) {
  REQUIRE(c && dst && src, "null argument");
  REQUIRE(src->N >= dst->N, "copy_vector: source shorter than destination");
  CK(cudaSetDevice(c->device));
  // like the reference, the destination's length decides (CSR/CPUContext.cpp:79)
  if (dst->N) CK(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * (size_t)dst->N, cudaMemcpyDeviceToDevice, c->stream));
  c->epoch++;
  return 0;
}

// ---------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------
}  // extern "C"

// persistent grid: SMs x resident CTAs of this mode's kernel (never more CTAs than tiles)
template <int MODE, bool DYN>
static int spmv_resident_ctas(abft_ctx *c) {
  int &cached = DYN ? c->spmv_ctas_per_sm_dyn : c->spmv_ctas_per_sm;
  if (cached == 0) {
    const int smem = (int)SpmvCfg<DYN>::dyn_smem;
    cudaFuncSetAttribute(spmv_kernel<MODE, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_kernel<MODE, DYN>, SpmvCfg<DYN>::block, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (DYN) {
      // leave the L1 what the stage buffers of the resident CTAs do not need (gathers in flight hold L1 lines);
      // a fifth CTA (mode none has the registers for it) would take the L1 back
      per_sm = std::min(per_sm, 4);
      const int need_kb = (per_sm * (smem + 1024) + 1023) / 1024;
      cudaFuncSetAttribute(spmv_kernel<MODE, DYN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           std::min(100, (need_kb * 100 + 227) / 228));
    }
    cached = per_sm;
  }
  return cached * c->num_sms;
}

// persistent grid: SMs x resident CTAs of this mode's kernel (never more CTAs than tiles)
template <int MODE, int FLEN>
static int spmv_stencil_ctas(abft_ctx *c) {
  static int cached[64] = {0};   // per device
  int &v = cached[c->device & 63];
  if (v == 0) {
    const int smem = (int)SpmvCfg<false>::dyn_smem;
    cudaFuncSetAttribute(spmv_kernel<MODE, false, FLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_kernel<MODE, false, FLEN>, SpmvCfg<false>::block, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    v = per_sm;
  }
  return v * c->num_sms;
}

template <int MODE, int FLEN>
static void launch_spmv_stencil(abft_ctx *c, const abft_matrix *m, const double *x, double *y, const ReduceOut &red,
                                const int *active, const WaitSpec &wait) {
  LogRef L{c->d_log, c->h_flag};
  const int grid = std::min(m->n_tiles, spmv_stencil_ctas<MODE, FLEN>(c));
  spmv_kernel<MODE, false, FLEN><<<grid, SpmvCfg<false>::block, SpmvCfg<false>::dyn_smem, c->stream>>>(m->view, x, y, L, red, active, wait);
}

template <int MODE>
static void launch_spmv_mode(abft_ctx *c, const abft_matrix *m, const double *x, double *y, const ReduceOut &red,
                             const int *active, const WaitSpec &wait) {
  LogRef L{c->d_log, c->h_flag};
  // the compute-bound modes have stencil instantiations (the others run at the HBM roofline as they are)
  constexpr bool has_stencil = (MODE & 7) == ABFT_MODE_SEC7 || (MODE & 7) == ABFT_MODE_SECDED || (MODE & 7) == ABFT_MODE_CRC32C;
  if (has_stencil && m->flen == 7) {
    launch_spmv_stencil<has_stencil ? MODE : ABFT_MODE_SECDED, 7>(c, m, x, y, red, active, wait);
  } else if (has_stencil && m->flen == 5) {
    launch_spmv_stencil<has_stencil ? MODE : ABFT_MODE_SECDED, 5>(c, m, x, y, red, active, wait);
  } else if (m->view.tile_counter) {
    const int grid = std::min(m->n_tiles, spmv_resident_ctas<MODE, true>(c));
    spmv_kernel<MODE, true><<<grid, SpmvCfg<true>::block, SpmvCfg<true>::dyn_smem, c->stream>>>(m->view, x, y, L, red, active, wait);
  } else {
    const int grid = std::min(m->n_tiles, spmv_resident_ctas<MODE, false>(c));
    spmv_kernel<MODE, false><<<grid, SpmvCfg<false>::block, SpmvCfg<false>::dyn_smem, c->stream>>>(m->view, x, y, L, red, active, wait);
  }
}

static int launch_spmv(abft_ctx *c, const abft_matrix *m, const double *x, double *y, const ReduceOut &red,
                       const int *active, WaitSpec wait = WaitSpec{nullptr, 0u, 0u, nullptr}) {
  wait.phase = c->phase_cur;
  if (m->n_tiles == 0) {
    // no rows: dot(vec, result) over an empty range is 0
    CK(cudaMemsetAsync(red.dev, 0, sizeof(double), c->stream));
    if (red.host) *const_cast<double *>(red.host) = 0.0;
    return 0;
  }
  switch (c->mode) {
    case ABFT_MODE_NONE: launch_spmv_mode<ABFT_MODE_NONE>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_CONSTRAINTS: launch_spmv_mode<ABFT_MODE_CONSTRAINTS>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_SED: launch_spmv_mode<ABFT_MODE_SED>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_SEC7: launch_spmv_mode<ABFT_MODE_SEC7>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_SEC8: launch_spmv_mode<ABFT_MODE_SEC8>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_SECDED: launch_spmv_mode<ABFT_MODE_SECDED>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_CRC32C: launch_spmv_mode<ABFT_MODE_CRC32C>(c, m, x, y, red, active, wait); break;
    case ABFT_MODE_SECDEDV: launch_spmv_mode<ABFT_MODE_SECDED | kVecSed>(c, m, x, y, red, active, wait); break;
    default: return fail(-1, "bad mode", __FILE__, __LINE__);
  }
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int abft_cuda_spmv(abft_ctx *c, abft_matrix *m, const abft_vector *vec, abft_vector *result) {
  REQUIRE(c && m && vec && result, "null argument");
  REQUIRE((uint32_t)result->N >= m->view.n_rows, "spmv: result shorter than the matrix has rows");
  REQUIRE((uint32_t)vec->N >= m->view.n_cols, "spmv: vector shorter than the matrix has columns");
  REQUIRE(vec->d != result->d, "spmv: vec and result must not alias");
  CK(cudaSetDevice(c->device));
  phase_next_row(c);
  int rc = launch_spmv(c, m, vec->d, result->d, make_reduce(c, SLOT_SPMV_DOT), nullptr);
  if (rc) return rc;
  c->epoch++;
  // dot(vec[row_offset .. row_offset+n_rows), result[0 .. n_rows)) is now in slot 0
  c->dot_a = vec->d + m->view.row_offset;
  c->dot_b = result->d;
  c->dot_n = (int)m->view.n_rows;
  c->dot_epoch = c->epoch;
  return 0;
}

static int launch_dot(abft_ctx *c, const double *a, const double *b, int N, const ReduceOut &red, const int *active) {
  const int ok = aligned16(a) && aligned16(b);
  const LogRef L{c->d_log, c->h_flag};
  if (c->vsed) { dot_kernel<true><<<vec_grid(c, N), kBlock, 0, c->stream>>>(a, b, N, ok, red, active, L); }
  else { dot_kernel<false><<<vec_grid(c, N), kBlock, 0, c->stream>>>(a, b, N, ok, red, active, L); }
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_dot_async(abft_ctx *c, const abft_vector *a, const abft_vector *b, int slot) {
  REQUIRE(c && a && b, "null argument");
  REQUIRE(slot >= 0 && slot < kNumSlots, "bad slot");
  REQUIRE(b->N >= a->N, "dot: second vector shorter than the first");
  CK(cudaSetDevice(c->device));
  return launch_dot(c, a->d, b->d, a->N, make_reduce(c, slot), nullptr);
}

int abft_cuda_dot(abft_ctx *c, const abft_vector *a, const abft_vector *b, double *out) {
  REQUIRE(c && a && b && out, "null argument");
  CK(cudaSetDevice(c->device));
  int slot = SLOT_DOT;
  const bool fused = c->dot_epoch == c->epoch && a->N == c->dot_n && b->N >= a->N &&
                     ((a->d == c->dot_a && b->d == c->dot_b) || (a->d == c->dot_b && b->d == c->dot_a));
  if (fused) {
    slot = SLOT_SPMV_DOT;  // the spmv just before this call already reduced it
  } else {
    int rc = abft_cuda_dot_async(c, a, b, SLOT_DOT);
    if (rc) return rc;
  }
  return wait_scalar(c, slot, out);
}

static int launch_calc_xr(abft_ctx *c, double *x, double *r, const double *p, const double *w, const ScalarArg &alpha,
                          bool dev, int N, const ReduceOut &red, const int *active) {
  const int ok = aligned16(x) && aligned16(r) && aligned16(p) && aligned16(w);
  const LogRef L{c->d_log, c->h_flag};
  if (dev && c->vsed) { calc_xr_kernel<true, true><<<vec_grid(c, N), kBlock, 0, c->stream>>>(x, r, p, w, alpha, N, ok, red, active, L); }
  else if (dev) { calc_xr_kernel<true, false><<<vec_grid(c, N), kBlock, 0, c->stream>>>(x, r, p, w, alpha, N, ok, red, active, L); }
  else if (c->vsed) { calc_xr_kernel<false, true><<<vec_grid(c, N), kBlock, 0, c->stream>>>(x, r, p, w, alpha, N, ok, red, active, L); }
  else { calc_xr_kernel<false, false><<<vec_grid(c, N), kBlock, 0, c->stream>>>(x, r, p, w, alpha, N, ok, red, active, L); }
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_calc_xr_async(abft_ctx *c, abft_vector *x, abft_vector *r, const abft_vector *p, const abft_vector *w,
                            double alpha, int slot) {
  REQUIRE(c && x && r && p && w, "null argument");
  REQUIRE(slot >= 0 && slot < kNumSlots, "bad slot");
  REQUIRE(r->N >= x->N && p->N >= x->N && w->N >= x->N, "calc_xr: vector lengths");
  CK(cudaSetDevice(c->device));
  ScalarArg a{alpha, nullptr, nullptr};
  int rc = launch_calc_xr(c, x->d, r->d, p->d, w->d, a, false, x->N, make_reduce(c, slot), nullptr);
  c->epoch++;
  return rc;
}

int abft_cuda_calc_xr(abft_ctx *c, abft_vector *x, abft_vector *r, const abft_vector *p, const abft_vector *w,
                      double alpha, double *out) {
  REQUIRE(out, "null argument");
  int rc = abft_cuda_calc_xr_async(c, x, r, p, w, alpha, SLOT_XR);
  if (rc) return rc;
  return wait_scalar(c, SLOT_XR, out);
}

static int launch_calc_p(abft_ctx *c, double *p, const double *r, const ScalarArg &beta, bool dev, int N,
                         const int *active) {
  const int ok = aligned16(p) && aligned16(r);
  const LogRef L{c->d_log, c->h_flag};
  if (dev && c->vsed) { calc_p_kernel<true, true><<<vec_grid(c, N), kBlock, 0, c->stream>>>(p, r, beta, N, ok, active, L); }
  else if (dev) { calc_p_kernel<true, false><<<vec_grid(c, N), kBlock, 0, c->stream>>>(p, r, beta, N, ok, active, L); }
  else if (c->vsed) { calc_p_kernel<false, true><<<vec_grid(c, N), kBlock, 0, c->stream>>>(p, r, beta, N, ok, active, L); }
  else { calc_p_kernel<false, false><<<vec_grid(c, N), kBlock, 0, c->stream>>>(p, r, beta, N, ok, active, L); }
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_calc_p(abft_ctx *c, abft_vector *p, const abft_vector *r, double beta) {
  REQUIRE(c && p && r, "null argument");
  REQUIRE(r->N >= p->N, "calc_p: vector lengths");
  CK(cudaSetDevice(c->device));
  ScalarArg b{beta, nullptr, nullptr};
  int rc = launch_calc_p(c, p->d, r->d, b, false, p->N, nullptr);
  c->epoch++;
  return rc;
}

double *abft_cuda_scalar_dev_ptr(abft_ctx *c, int slot) {
  return (c && slot >= 0 && slot < kNumSlots) ? c->d_scalars + slot : nullptr;
}
double abft_cuda_scalar_host(abft_ctx *c, int slot) {
  return (c && slot >= 0 && slot < kNumSlots) ? c->h_scalars[slot] : 0.0;
}

// ---------------------------------------------------------------------------
// fault reporting
// ---------------------------------------------------------------------------
int abft_cuda_events_pending(abft_ctx *c) { return (c && (*c->h_flag || c->pending_pos < c->pending.size())) ? 1 : 0; }

// Drains the device log into the context (sorted into the reference's serial order, a persisting fault reported
// once) and hands out up to `cap` records per call; callers poll until it returns 0.  The hand-out stops at the first
// fatal record (the reference has exit(1)ed there) and discards what follows it.
int abft_cuda_poll_events(abft_ctx *c, abft_event *out, int cap, int *fatal) {
  if (fatal) *fatal = 0;
  if (!c) return -1;
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
    fail(-1, "poll_events: sync failed", __FILE__, __LINE__);
    return -1;
  }
  if (c->pending_pos >= c->pending.size() && *c->h_flag) {
    c->pending.clear();
    c->pending_pos = 0;
    unsigned int raw = 0;
    if (cudaMemcpy(&raw, &c->d_log->count, sizeof raw, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const unsigned int count = std::min(raw, (unsigned)kLogCap);
    std::vector<abft_event> ev(count);
    if (count && cudaMemcpy(ev.data(), c->d_log->ev, sizeof(abft_event) * count, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    // on the context's own (non-blocking) stream: a legacy-stream memset would not be ordered before the next kernel
    cudaMemsetAsync(c->d_log, 0, sizeof(unsigned int) * 2, c->stream);
    cudaStreamSynchronize(c->stream);
    *c->h_flag = 0;

    // serial order of the reference: rows ascending; within a row the row checks
    // come before its elements, elements ascending (CSR/CPUContext.cpp:166-206)
    const bool cons = c->base == ABFT_MODE_CONSTRAINTS;
    auto key = [&](const abft_event &e, uint64_t &k1, int &k2, uint64_t &k3) {
      const bool rowev = e.kind == ABFT_EV_ROW_SIZE || e.kind == ABFT_EV_ROW_ORDER || e.kind == ABFT_EV_CRC_ROW;
      if (cons) {
        k1 = rowev ? e.index : e.bit;  // element events of CONSTRAINTS carry their row in `bit`
        k2 = rowev ? 0 : 1;
        k3 = e.index;
      } else {
        k1 = e.index;
        k2 = 0;
        k3 = e.bit;
      }
    };
    auto rowptr_ev = [](const abft_event &e) { return e.kind == ABFT_EV_ROWPTR_CORRECTED || e.kind == ABFT_EV_ROWPTR_DOUBLE; };
    std::sort(ev.begin(), ev.end(), [&](const abft_event &a, const abft_event &b) {
      // the row pointers are checked before the kernel looks at them: their events come first, by pair (rowsecc.cuh)
      if (rowptr_ev(a) != rowptr_ev(b)) return rowptr_ev(a);
      if (rowptr_ev(a)) return a.index != b.index ? a.index < b.index : a.kind < b.kind;
      uint64_t a1, a3, b1, b3;
      int a2, b2;
      key(a, a1, a2, a3);
      key(b, b1, b2, b3);
      if (a1 != b1) return a1 < b1;
      if (a2 != b2) return a2 < b2;
      if (a3 != b3) return a3 < b3;
      return a.kind < b.kind;
    });
    for (unsigned int i = 0; i < count; i++) {
      abft_event e = ev[i];
      if (cons && (e.kind == ABFT_EV_COL_SIZE || e.kind == ABFT_EV_COL_ORDER)) e.bit = 0;
      // a fault that persists is found again by every spmv until it is polled: report it once
      if (!c->pending.empty() && c->pending.back().kind == e.kind && c->pending.back().index == e.index &&
          c->pending.back().bit == e.bit)
        continue;
      c->pending.push_back(e);
    }
    if (raw > (unsigned)kLogCap) c->pending.push_back(abft_event{ABFT_EV_LOG_OVERFLOW, raw - (unsigned)kLogCap, 0u});
  }
  int n = 0;
  while (n < cap && c->pending_pos < c->pending.size()) {
    const abft_event e = c->pending[c->pending_pos++];
    out[n++] = e;
    if (abft_cuda_event_is_fatal(e.kind)) {
      if (fatal) *fatal = 1;
      c->pending.clear();  // the reference exit(1)s here
      c->pending_pos = 0;
      break;
    }
  }
  return n;
}


// ---------------------------------------------------------------------------
// multi-GPU: peer-memory exchange over NVLink
// ---------------------------------------------------------------------------
abft_comm *abft_cuda_comm_create(abft_ctx *c, int rank, int world) {
  if (!c || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
    fail(-1, "bad arguments to comm_create (1..8 ranks)", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(c->device));
  abft_comm *m = new abft_comm();
  m->rank = rank;
  m->world = world;
  CommWindow init;
  memset(&init, 0, sizeof init);
  for (int par = 0; par < 2; par++)
    for (int s2 = 0; s2 < kMaxRanks; s2++) memcpy(&init.slots[par][s2], &kSlotEmpty, sizeof(double));  // no contribution yet
  cudaError_t e = cudaMalloc(&m->win[rank], sizeof(CommWindow));
  if (e == cudaSuccess) e = cudaMemcpy(m->win[rank], &init, sizeof(CommWindow), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fail((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    cudaFree(m->win[rank]);
    delete m;
    return nullptr;
  }
  return m;
}

void abft_cuda_comm_destroy(abft_ctx *c, abft_comm *m) {
  if (!m) return;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  for (int i = 0; i < m->n_ipc; i++) cudaIpcCloseMemHandle(m->ipc_opened[i]);
  cudaFree(m->win[m->rank]);
  delete m;
}

int abft_cuda_comm_window_handle(abft_ctx *c, abft_comm *m, unsigned char *out64) {
  REQUIRE(c && m && out64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == ABFT_IPC_HANDLE_BYTES, "IPC handle size");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, m->win[m->rank]));
  memcpy(out64, &h, sizeof h);
  return 0;
}

void *abft_cuda_comm_window_ptr(abft_comm *m) { return m ? (void *)m->win[m->rank] : nullptr; }

int abft_cuda_vector_ipc_handle(abft_ctx *c, abft_vector *v, unsigned char *out64) {
  REQUIRE(c && v && out64, "null argument");
  REQUIRE(v->owns, "only vectors created by abft_cuda_vector_create can be exported");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, v->d));
  memcpy(out64, &h, sizeof h);
  return 0;
}

int abft_cuda_comm_attach_self(abft_ctx *c, abft_comm *m, abft_vector *vec_full) {
  REQUIRE(c && m && vec_full, "null argument");
  m->peer_vec[m->rank] = vec_full->d;
  return 0;
}

int abft_cuda_comm_attach_peer(abft_ctx *c, abft_comm *m, int peer, const unsigned char *win_handle64,
                               const unsigned char *vec_handle64) {
  REQUIRE(c && m && win_handle64 && vec_handle64, "null argument");
  REQUIRE(peer >= 0 && peer < m->world && peer != m->rank, "bad peer rank");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t hw, hv;
  memcpy(&hw, win_handle64, sizeof hw);
  memcpy(&hv, vec_handle64, sizeof hv);
  void *pw = nullptr, *pv = nullptr;
  CK(cudaIpcOpenMemHandle(&pw, hw, cudaIpcMemLazyEnablePeerAccess));
  m->ipc_opened[m->n_ipc++] = pw;
  CK(cudaIpcOpenMemHandle(&pv, hv, cudaIpcMemLazyEnablePeerAccess));
  m->ipc_opened[m->n_ipc++] = pv;
  m->win[peer] = (CommWindow *)pw;
  m->peer_vec[peer] = (double *)pv;
  return 0;
}

int abft_cuda_comm_attach_peer_ptrs(abft_ctx *c, abft_comm *m, int peer, void *window_dev, double *vec_dev,
                                    int peer_device) {
  REQUIRE(c && m && window_dev && vec_dev, "null argument");
  REQUIRE(peer >= 0 && peer < m->world && peer != m->rank, "bad peer rank");
  CK(cudaSetDevice(c->device));
  if (peer_device != c->device) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    cudaGetLastError();
  }
  m->win[peer] = (CommWindow *)window_dev;
  m->peer_vec[peer] = vec_dev;
  return 0;
}

int abft_cuda_comm_set_plan(abft_ctx *c, abft_comm *m, const int *push_lo, const int *push_hi, unsigned int wait_mask,
                            int row0) {
  REQUIRE(c && m && push_lo && push_hi, "null argument");
  for (int s = 0; s < m->world; s++) {
    m->push_lo[s] = push_lo[s];
    m->push_hi[s] = push_hi[s];
  }
  m->wait_mask = wait_mask & ~(1u << m->rank);
  m->row0 = row0;
  return 0;
}

}  // extern "C"

static int comm_ready(abft_comm *m) {
  for (int s = 0; s < m->world; s++)
    if (!m->win[s] || !m->peer_vec[s]) return 0;
  return 1;
}

static int comm_windows_ready(abft_comm *m) {
  for (int s = 0; s < m->world; s++)
    if (!m->win[s]) return 0;
  return 1;
}

template <bool COMPUTE>
static int launch_push(abft_ctx *c, abft_comm *m, abft_vector *p_local, const abft_vector *r, double beta) {
  REQUIRE(comm_ready(m), "comm: a peer has not been attached");
  REQUIRE(p_local->d >= m->peer_vec[m->rank], "dist: p_local must be a view of the exchange vector");
  REQUIRE((p_local->d - m->peer_vec[m->rank]) == m->row0, "dist: p_local must start at this rank's first row");
  PushPlan plan{};
  for (int s = 0; s < m->world; s++) {
    plan.peer_vec[s] = m->peer_vec[s];
    plan.win[s] = m->win[s];
    plan.lo[s] = m->push_lo[s];
    plan.hi[s] = m->push_hi[s];
  }
  plan.world = m->world;
  plan.rank = m->rank;
  plan.row0 = m->row0;
  plan.epoch = ++m->epoch;
  plan.phase = c->phase_cur;
  ScalarArg b{beta, nullptr, nullptr};
  const int N = p_local->N;
  long long want = ((long long)N + kBlock * 8 - 1) / (kBlock * 8);
  const int grid = (int)std::max(1LL, std::min(want, (long long)c->num_sms * 8));
  const LogRef L{c->d_log, c->h_flag};
  if (c->vsed) { calc_p_push_kernel<false, COMPUTE, true><<<grid, kBlock, 0, c->stream>>>(p_local->d, COMPUTE ? r->d : p_local->d, b, N, plan, nullptr, L); }
  else { calc_p_push_kernel<false, COMPUTE, false><<<grid, kBlock, 0, c->stream>>>(p_local->d, COMPUTE ? r->d : p_local->d, b, N, plan, nullptr, L); }
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int abft_cuda_dist_push(abft_ctx *c, abft_comm *m, abft_vector *p_local) {
  REQUIRE(c && m && p_local, "null argument");
  CK(cudaSetDevice(c->device));
  return launch_push<false>(c, m, p_local, nullptr, 0.0);
}

int abft_cuda_dist_calc_p(abft_ctx *c, abft_comm *m, abft_vector *p_local, const abft_vector *r, double beta) {
  REQUIRE(c && m && p_local && r, "null argument");
  REQUIRE(r->N >= p_local->N, "calc_p: vector lengths");
  CK(cudaSetDevice(c->device));
  return launch_push<true>(c, m, p_local, r, beta);
}

int abft_cuda_dist_spmv(abft_ctx *c, abft_comm *m, abft_matrix *A, const abft_vector *vec_full, abft_vector *result) {
  REQUIRE(c && m && A && vec_full && result, "null argument");
  REQUIRE(comm_ready(m), "comm: a peer has not been attached");
  REQUIRE(vec_full->d == m->peer_vec[m->rank], "dist_spmv: vec must be the exchange vector");
  REQUIRE((uint32_t)result->N >= A->view.n_rows && (uint32_t)vec_full->N >= A->view.n_cols, "dist_spmv: vector lengths");
  CK(cudaSetDevice(c->device));
  phase_next_row(c);
  WaitSpec w{m->win[m->rank], m->epoch, m->wait_mask, nullptr};
  int rc = launch_spmv(c, A, vec_full->d, result->d, make_reduce(c, SLOT_SPMV_DOT, m), nullptr, w);
  if (rc) return rc;
  c->epoch++;
  c->dot_a = vec_full->d + A->view.row_offset;
  c->dot_b = result->d;
  c->dot_n = (int)A->view.n_rows;
  c->dot_epoch = c->epoch;
  return 0;
}

// asynchronous: *slot_out tells where the (global) value will be once the stream has drained
int abft_cuda_dist_dot_async(abft_ctx *c, abft_comm *m, const abft_vector *a, const abft_vector *b, int *slot_out) {
  REQUIRE(c && m && a && b && slot_out, "null argument");
  REQUIRE(comm_windows_ready(m), "comm: a peer window has not been attached");
  CK(cudaSetDevice(c->device));
  const bool fused = c->dot_epoch == c->epoch && a->N == c->dot_n && b->N >= a->N &&
                     ((a->d == c->dot_a && b->d == c->dot_b) || (a->d == c->dot_b && b->d == c->dot_a));
  if (fused) {
    *slot_out = SLOT_SPMV_DOT;  // dist_spmv already reduced it over all ranks
    return 0;
  }
  REQUIRE(b->N >= a->N, "dot: second vector shorter than the first");
  *slot_out = SLOT_DOT;
  return launch_dot(c, a->d, b->d, a->N, make_reduce(c, SLOT_DOT, m), nullptr);
}

int abft_cuda_dist_dot(abft_ctx *c, abft_comm *m, const abft_vector *a, const abft_vector *b, double *out) {
  REQUIRE(out, "null argument");
  int slot = SLOT_DOT;
  int rc = abft_cuda_dist_dot_async(c, m, a, b, &slot);
  if (rc) return rc;
  return wait_scalar(c, slot, out);
}

int abft_cuda_dist_calc_xr_async(abft_ctx *c, abft_comm *m, abft_vector *x, abft_vector *r, const abft_vector *p,
                                 const abft_vector *w, double alpha, int *slot_out) {
  REQUIRE(c && m && x && r && p && w && slot_out, "null argument");
  REQUIRE(comm_windows_ready(m), "comm: a peer window has not been attached");
  REQUIRE(r->N >= x->N && p->N >= x->N && w->N >= x->N, "calc_xr: vector lengths");
  CK(cudaSetDevice(c->device));
  ScalarArg a{alpha, nullptr, nullptr};
  *slot_out = SLOT_XR;
  int rc = launch_calc_xr(c, x->d, r->d, p->d, w->d, a, false, x->N, make_reduce(c, SLOT_XR, m), nullptr);
  c->epoch++;
  return rc;
}

int abft_cuda_dist_calc_xr(abft_ctx *c, abft_comm *m, abft_vector *x, abft_vector *r, const abft_vector *p,
                           const abft_vector *w, double alpha, double *out) {
  REQUIRE(c && m && x && r && p && w && out, "null argument");
  REQUIRE(r->N >= x->N && p->N >= x->N && w->N >= x->N, "calc_xr: vector lengths");
  CK(cudaSetDevice(c->device));
  ScalarArg a{alpha, nullptr, nullptr};
  int rc = launch_calc_xr(c, x->d, r->d, p->d, w->d, a, false, x->N, make_reduce(c, SLOT_XR, m), nullptr);
  c->epoch++;
  if (rc) return rc;
  return wait_scalar(c, SLOT_XR, out);
}

// ---------------------------------------------------------------------------
// device-resident CG loop
// ---------------------------------------------------------------------------
}  // extern "C"

// One implementation for the single-GPU and the row-partitioned loop.  comm == nullptr: p_full == p_local.
static int cg_solve_impl(abft_ctx *c, abft_comm *m, abft_matrix *A, const abft_vector *b, abft_vector *x, abft_vector *r,
                         abft_vector *p_local, abft_vector *p_full, abft_vector *w, int max_itrs, double conv_threshold,
                         int check_every, double *rr_out) {
  const int N = (int)A->view.n_rows;
  if (check_every < 1) check_every = 1;
  if (cudaSetDevice(c->device) != cudaSuccess) return -1;
  cudaStream_t s = c->stream;
  int *flag0 = c->d_loop, *iters = c->d_loop + 2;
  if (cudaMemsetAsync(c->d_loop, 0, sizeof(int) * 4, s) != cudaSuccess) {
    fail(-1, "cg_solve: memset failed", __FILE__, __LINE__);
    return -1;
  }
  c->h_loop[0] = 0;
  c->h_loop[1] = 0;
  const unsigned int seq0 = m ? m->seq : 0u, epoch0 = m ? m->epoch : 0u;

  auto push_plan = [&](PushPlan &plan) {
    for (int k = 0; k < m->world; k++) {
      plan.peer_vec[k] = m->peer_vec[k];
      plan.win[k] = m->win[k];
      plan.lo[k] = m->push_lo[k];
      plan.hi[k] = m->push_hi[k];
    }
    plan.world = m->world;
    plan.rank = m->rank;
    plan.row0 = m->row0;
    plan.epoch = ++m->epoch;
    plan.phase = c->phase_cur;
  };
  const long long want = ((long long)N + kBlock * 8 - 1) / (kBlock * 8);
  const int push_grid = (int)std::max(1LL, std::min(want, (long long)c->num_sms * 8));

  // r = b ; p = r (+ exchange) ; rr = r.r   (cg.cpp:87-91)
  if (cudaMemcpyAsync(r->d, b->d, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(p_local->d, r->d, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
    fail(-1, "cg_solve: device copy failed", __FILE__, __LINE__);
    return -1;
  }
  if (m) {
    PushPlan plan{};
    push_plan(plan);
    ScalarArg none{0.0, nullptr, nullptr};
    calc_p_push_kernel<false, false, false><<<push_grid, kBlock, 0, s>>>(p_local->d, p_local->d, none, N, plan, nullptr, LogRef{c->d_log, c->h_flag});
    c->launches++;
  }
  {
    ReduceOut red = make_reduce(c, SLOT_RR0, m);
    red.host = nullptr;  // (nobody on the host looks at the scalars while the loop runs: no PCIe store per reduction)
    red.active = flag0;  // iteration 0 reads flag[0]
    red.iters = iters;
    red.threshold = conv_threshold;
    red.max_itrs = max_itrs;
    red.bump_iters = 0;
    if (launch_dot(c, r->d, r->d, N, red, nullptr)) return -1;
  }
  c->epoch++;
  int launched = 0;
  int done_iters = 0;
  bool active = true;
  bool fatal_seen = false;
  while (active && launched < max_itrs) {
    const int batch = std::min(check_every, max_itrs - launched);
    for (int j = 0; j < batch; j++, launched++) {
      const int q = launched & 1;
      const int *flag_cur = c->d_loop + q;
      int *flag_next = c->d_loop + (q ^ 1);
      const double *rr_cur = c->d_scalars + (q ? SLOT_RR1 : SLOT_RR0);
      const int slot_new = q ? SLOT_RR0 : SLOT_RR1;
      // w = A p, pw = p.w (summed over all ranks inside the kernel's finalizer)
      phase_next_row(c);
      WaitSpec wait{m ? m->win[m->rank] : nullptr, m ? m->epoch : 0u, m ? m->wait_mask : 0u, nullptr};
      ReduceOut red_pw = make_reduce(c, SLOT_SPMV_DOT, m);
      red_pw.host = nullptr;
      if (launch_spmv(c, A, p_full->d, w->d, red_pw, flag_cur, wait)) return -1;
      // alpha = rr / pw ; x += alpha p ; r -= alpha w ; rr_new = r.r
      ReduceOut red = make_reduce(c, slot_new, m);
      red.host = nullptr;
      red.active = flag_next;
      red.iters = iters;
      red.threshold = conv_threshold;
      red.max_itrs = max_itrs;
      red.bump_iters = 1;
      ScalarArg alpha{0.0, rr_cur, c->d_scalars + SLOT_SPMV_DOT};
      if (launch_calc_xr(c, x->d, r->d, p_local->d, w->d, alpha, true, N, red, flag_cur)) return -1;
      // beta = rr_new / rr ; p = r + beta p (+ push of the new entries to the peers)
      ScalarArg beta{0.0, c->d_scalars + slot_new, rr_cur};
      if (m) {
        PushPlan plan{};
        push_plan(plan);
        if (c->vsed) { calc_p_push_kernel<true, true, true><<<push_grid, kBlock, 0, s>>>(p_local->d, r->d, beta, N, plan, flag_cur, LogRef{c->d_log, c->h_flag}); }
        else { calc_p_push_kernel<true, true, false><<<push_grid, kBlock, 0, s>>>(p_local->d, r->d, beta, N, plan, flag_cur, LogRef{c->d_log, c->h_flag}); }
        c->launches++;
      } else if (launch_calc_p(c, p_local->d, r->d, beta, true, N, flag_cur)) {
        return -1;
      }
    }
    // read back progress: iterations done and whether the next one would run
    if (cudaMemcpyAsync(c->h_loop, c->d_loop + 2, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(c->h_loop + 1, c->d_loop + (launched & 1), sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      fail(-1, "cg_solve: stream sync failed", __FILE__, __LINE__);
      return -1;
    }
    done_iters = c->h_loop[0];
    active = c->h_loop[1] != 0;
    if (*c->h_flag) {
      // a fault was logged; a fatal one stops the single-GPU solve like the reference's exit(1).  Ranks of a
      // partitioned solve must keep issuing the same kernels (they wait on each other): they finish the loop
      // and the caller finds the fault with abft_cuda_poll_events afterwards.
      unsigned int count = 0;
      cudaMemcpy(&count, &c->d_log->count, sizeof count, cudaMemcpyDeviceToHost);
      count = std::min(count, (unsigned)kLogCap);
      std::vector<abft_event> ev(count);
      if (count) cudaMemcpy(ev.data(), c->d_log->ev, sizeof(abft_event) * count, cudaMemcpyDeviceToHost);
      for (auto &e : ev)
        if (abft_cuda_event_is_fatal(e.kind)) fatal_seen = true;
      if (fatal_seen && !m) {
        c->epoch++;
        return -2;
      }
    }
  }
  c->epoch++;
  if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
  if (m) {
    // kernels past convergence never ran: take back the sequence numbers they were given, so that the
    // next reduction / push continues consecutively (the slot parity relies on it)
    m->seq = seq0 + 1u + 2u * (unsigned)done_iters;
    m->epoch = epoch0 + 1u + (unsigned)done_iters;
  }
  if (rr_out && cudaMemcpy(rr_out, c->d_scalars + ((done_iters & 1) ? SLOT_RR1 : SLOT_RR0), sizeof(double),
                           cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  // a partitioned solve that saw a fatal fault on this rank ran on in lock-step with its peers (above); it still
  // reports it: the reference would have exit(1)ed inside that spmv
  if (fatal_seen) return -2;
  return done_iters;
}

extern "C" {

int abft_cuda_cg_solve(abft_ctx *c, abft_matrix *A, const abft_vector *b, abft_vector *x, abft_vector *r,
                       abft_vector *p, abft_vector *w, int max_itrs, double conv_threshold, int check_every,
                       double *rr_out) {
  if (!c || !A || !b || !x || !r || !p || !w) {
    fail(-1, "null argument", __FILE__, __LINE__);
    return -1;
  }
  const int N = (int)A->view.n_rows;
  if (A->view.row_offset != 0 || A->view.n_cols != (uint32_t)N || b->N < N || x->N < N || r->N < N || p->N < N ||
      w->N < N) {
    fail(-1, "cg_solve needs a whole (unpartitioned) matrix and vectors of its size", __FILE__, __LINE__);
    return -1;
  }
  return cg_solve_impl(c, nullptr, A, b, x, r, p, p, w, max_itrs, conv_threshold, check_every, rr_out);
}

int abft_cuda_dist_cg_solve(abft_ctx *c, abft_comm *m, abft_matrix *A, const abft_vector *b, abft_vector *x,
                            abft_vector *r, abft_vector *p_local, abft_vector *p_full, abft_vector *w, int max_itrs,
                            double conv_threshold, int check_every, double *rr_out) {
  if (!c || !m || !A || !b || !x || !r || !p_local || !p_full || !w) {
    fail(-1, "null argument", __FILE__, __LINE__);
    return -1;
  }
  const int N = (int)A->view.n_rows;
  if (!comm_ready(m) || p_full->d != m->peer_vec[m->rank] || p_local->d != p_full->d + m->row0 || p_local->N < N ||
      (uint32_t)p_full->N < A->view.n_cols || b->N < N || x->N < N || r->N < N || w->N < N) {
    fail(-1, "dist_cg_solve: vectors do not match the partition / exchange vector", __FILE__, __LINE__);
    return -1;
  }
  return cg_solve_impl(c, m->world > 1 ? m : nullptr, A, b, x, r, p_local, p_full, w, max_itrs, conv_threshold,
                       check_every, rr_out);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// the plugin path as one call: cg.cpp's loop body over the four public entry points
// ---------------------------------------------------------------------------
extern "C" {

// cg.cpp:94-118 -- spmv, dot, calc_xr, calc_p through the public entry points above, scalars on the host, exactly what
// the CGContext plugin does for the unmodified driver -- compiled, so that a caller in an interpreted language (the
// tests, bench.py) measures the library and not its own call overhead.  Runs until max_itrs iterations are done or
// rr <= conv_threshold; *rr_io carries r.r in and out; rr_hist (may be null) receives r.r after every iteration.
// Returns the iterations done.  If a fault has been logged, the loop stops BEFORE using that spmv's result (the plugin
// reports faults at this point, CSR/CPUContext.cpp prints from inside spmv): the caller polls the events and, unless
// one was fatal, calls again to go on (the interrupted iteration is redone from its spmv, on the repaired matrix).
int abft_cuda_cg_host_loop(abft_ctx *c, abft_matrix *A, abft_vector *x, abft_vector *r, abft_vector *p, abft_vector *w,
                           int max_itrs, double conv_threshold, double *rr_io, double *rr_hist) {
  if (!c || !A || !x || !r || !p || !w || !rr_io) {
    fail(-1, "null argument", __FILE__, __LINE__);
    return -1;
  }
  double rr = *rr_io;
  int itr = 0;
  for (; itr < max_itrs && rr > conv_threshold; itr++) {
    double pw = 0.0, rr_new = 0.0;
    if (abft_cuda_spmv(c, A, p, w) || abft_cuda_dot(c, p, w, &pw)) return -1;
    if (*c->h_flag) break;
    const double alpha = rr / pw;
    if (abft_cuda_calc_xr(c, x, r, p, w, alpha, &rr_new)) return -1;
    const double beta = rr_new / rr;
    if (abft_cuda_calc_p(c, p, r, beta)) return -1;
    rr = rr_new;
    if (rr_hist) rr_hist[itr] = rr;
  }
  *rr_io = rr;
  return itr;
}

// the same over a row partition (abft_cuda_dist_*): every rank runs it in lock-step
int abft_cuda_dist_cg_host_loop(abft_ctx *c, abft_comm *m, abft_matrix *A, abft_vector *x, abft_vector *r,
                                abft_vector *p_local, abft_vector *p_full, abft_vector *w, int max_itrs,
                                double conv_threshold, double *rr_io, double *rr_hist) {
  if (!c || !m || !A || !x || !r || !p_local || !p_full || !w || !rr_io) {
    fail(-1, "null argument", __FILE__, __LINE__);
    return -1;
  }
  double rr = *rr_io;
  int itr = 0;
  for (; itr < max_itrs && rr > conv_threshold; itr++) {
    double pw = 0.0, rr_new = 0.0;
    if (abft_cuda_dist_spmv(c, m, A, p_full, w) || abft_cuda_dist_dot(c, m, p_local, w, &pw)) return -1;
    // (a fault seen by one rank only: the ranks must keep issuing the same kernels, they wait on each other; the
    // caller finds the fault with abft_cuda_poll_events after the loop, like abft_cuda_dist_cg_solve)
    const double alpha = pw != 0.0 ? rr / pw : 0.0;
    if (abft_cuda_dist_calc_xr(c, m, x, r, p_local, w, alpha, &rr_new)) return -1;
    const double beta = rr != 0.0 ? rr_new / rr : 0.0;
    if (abft_cuda_dist_calc_p(c, m, p_local, r, beta)) return -1;
    rr = rr_new;
    if (rr_hist) rr_hist[itr] = rr;
  }
  *rr_io = rr;
  return itr;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// COO format (the reference's cg-coo): see coo.cuh for the design and its status
// ---------------------------------------------------------------------------
struct abft_coo_matrix {
  CooView view{};
  uint32_t *d_perm = nullptr, *d_colptr = nullptr;
  uint32_t *d_pos = nullptr;     // grouped layout (ABFT_CUDA_COO_LAYOUT=grouped): element index -> slot; else null
  std::vector<uint32_t> h_perm;  // grouped layout: host copy, to hand the elements back in the caller's order
  bool stream = true;            // grouped layout: the TMA-staged kernel (coo_stream.cuh); ABFT_CUDA_COO_KERNEL=direct: coo_spmv_kernel
};

// persistent grid of the streamed kernel: SMs x resident CTAs (never more CTAs than tiles)
template <int MODE>
static int coo_stream_ctas(abft_ctx *c) {
  static int cached[64] = {0};   // per device
  int &v = cached[c->device & 63];
  if (v == 0) {
    cudaFuncSetAttribute(coo_spmv_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCooDynSmem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coo_spmv_stream_kernel<MODE>, kBlock, kCooDynSmem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    v = per_sm;
  }
  return v * c->num_sms;
}

// returns true if dot(vec, result) was reduced along the way (slot SLOT_SPMV_DOT)
template <int MODE>
static bool launch_coo_spmv_mode(abft_ctx *c, const abft_coo_matrix *m, const double *vec, double *result) {
  LogRef L{c->d_log, c->h_flag};
  const int grid = std::max(1, std::min((int)((m->view.N + kBlock - 1) / kBlock), c->num_sms * 8));
  if (m->view.pos && m->stream) {
    const int tiles = (int)((m->view.N + kCooTileCols - 1) / kCooTileCols);
    ReduceOut red = make_reduce(c, SLOT_SPMV_DOT);
    volatile double *host = red.host;
    red.host = nullptr;   // published by the stray fix-up kernel, after it
    red.phase = nullptr;
    coo_spmv_stream_kernel<MODE><<<std::max(1, std::min(tiles, coo_stream_ctas<MODE>(c))), kBlock, kCooDynSmem, c->stream>>>(m->view, vec, result, L, red);
    coo_stray_kernel<MODE, true><<<1, 32, 0, c->stream>>>(m->view, vec, result, L, red.dev, host);
    return m->view.N > 0;
  } else if (m->view.pos) {
    coo_spmv_kernel<MODE, true><<<grid, kBlock, 0, c->stream>>>(m->view, vec, result, L);
    coo_stray_kernel<MODE, true><<<1, 32, 0, c->stream>>>(m->view, vec, result, L);
  } else {
    coo_spmv_kernel<MODE, false><<<grid, kBlock, 0, c->stream>>>(m->view, vec, result, L);
    coo_stray_kernel<MODE, false><<<1, 32, 0, c->stream>>>(m->view, vec, result, L);
  }
  return false;
}

extern "C" {

int abft_cuda_coo_format_event(const abft_event *e, char *buf, int buflen) {
  switch (e->kind) {  // COO/CPUContext.cpp:159,164,177,187; the ECC messages are the CSR target's
    case ABFT_EV_ROW_SIZE: return snprintf(buf, buflen, "row size constraint violated for index %d", (int)e->index);
    case ABFT_EV_ROW_ORDER: return snprintf(buf, buflen, "row index order violated at index %d", (int)e->index);
    case ABFT_EV_COL_SIZE: return snprintf(buf, buflen, "column size constraint violated for index %d", (int)e->index);
    case ABFT_EV_COL_ORDER: return snprintf(buf, buflen, "column index order violated at index %d", (int)e->index);
    default: return abft_cuda_format_event(e, buf, buflen);
  }
}

abft_coo_matrix *abft_cuda_coo_matrix_create(abft_ctx *c, const uint32_t *columns, const uint32_t *rows,
                                             const double *values, int N, int nnz) {
  if (!c || N < 0 || nnz < 0 || (nnz > 0 && (!columns || !rows || !values))) {
    fail(-1, "bad arguments to coo_matrix_create", __FILE__, __LINE__);
    return nullptr;
  }
  if (c->mode == ABFT_MODE_CRC32C || c->mode == ABFT_MODE_SECDEDV) {
    fail(-1, "crc32c / secdedv are CSR-only modes (the reference's COO target has none, and this repository defines none)",
         __FILE__, __LINE__);
    return nullptr;
  }
  if (c->mode >= ABFT_MODE_SED && (unsigned)N > (1u << 24)) {
    fail(-1, "ECC modes keep the check bits in column bits 24..31: N must be <= 2^24", __FILE__, __LINE__);
    return nullptr;
  }
  // (colptr: kCooColptrPad entries past the end, so that the streamed kernel's 16-byte-rounded window stays inside)
  std::vector<uint32_t> colptr((size_t)N + 1 + kCooColptrPad, (uint32_t)nnz), perm((size_t)std::max(nnz, 1));
  if (!coo_build_groups(columns, N, nnz, colptr.data(), perm.data())) {
    fail(-1, "coo_matrix_create: column index out of range", __FILE__, __LINE__);
    return nullptr;
  }
  CKP(cudaSetDevice(c->device));
  abft_coo_matrix *m = new abft_coo_matrix();
  const size_t nz = (size_t)std::max(nnz, 1);
  uint32_t *d_cols = nullptr, *d_rows = nullptr;
  double *d_vals = nullptr;
  struct Cleanup {
    abft_coo_matrix *&m;
    uint32_t *&a, *&b;
    double *&v;
    bool armed = true;
    ~Cleanup() {
      cudaFree(a);
      cudaFree(b);
      cudaFree(v);
      if (!armed) return;
      cudaFree(m->view.elements);
      cudaFree(m->view.stray);
      cudaFree(m->d_perm);
      cudaFree(m->d_colptr);
      cudaFree(m->d_pos);
      delete m;
    }
  } cleanup{m, d_cols, d_rows, d_vals};
  // the elements are stored in group order (coo.cuh, GROUPED: 0.46 instead of 0.60 ms per secded spmv on 256^3) unless
  // ABFT_CUDA_COO_LAYOUT=caller asks for the caller's order (kept for the tests: both layouts against the oracle)
  const char *layout = getenv("ABFT_CUDA_COO_LAYOUT");
  const bool grouped = !(layout && !strcmp(layout, "caller")) && nnz > 0;
  const char *kernel = getenv("ABFT_CUDA_COO_KERNEL");   // test hook: "direct" keeps coo_spmv_kernel on the grouped layout
  m->stream = !(kernel && !strcmp(kernel, "direct"));
  std::vector<uint32_t> pos;
  if (grouped) {
    pos.resize((size_t)nnz);
    for (int k = 0; k < nnz; k++) pos[perm[(size_t)k]] = (uint32_t)k;
    CKP(cudaMalloc(&m->d_pos, sizeof(uint32_t) * nz));
    CKP(cudaMemcpyAsync(m->d_pos, pos.data(), sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  CKP(cudaMalloc(&m->view.elements, sizeof(uint4) * nz));
  CKP(cudaMalloc(&m->view.stray, sizeof(uint32_t) * (1 + kMaxStray)));
  CKP(cudaMalloc(&m->d_perm, sizeof(uint32_t) * nz));
  CKP(cudaMalloc(&m->d_colptr, sizeof(uint32_t) * colptr.size()));
  CKP(cudaMalloc(&d_cols, sizeof(uint32_t) * nz));
  CKP(cudaMalloc(&d_rows, sizeof(uint32_t) * nz));
  CKP(cudaMalloc(&d_vals, sizeof(double) * nz));
  CKP(cudaMemsetAsync(m->view.stray, 0, sizeof(uint32_t), c->stream));
  CKP(cudaMemcpyAsync(m->d_colptr, colptr.data(), sizeof(uint32_t) * colptr.size(), cudaMemcpyHostToDevice, c->stream));
  if (nnz > 0) {
    CKP(cudaMemcpyAsync(m->d_perm, perm.data(), sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    CKP(cudaMemcpyAsync(d_cols, columns, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    CKP(cudaMemcpyAsync(d_rows, rows, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    CKP(cudaMemcpyAsync(d_vals, values, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    coo_encode_kernel<<<c->num_sms * 8, kBlock, 0, c->stream>>>(c->mode, d_cols, d_rows, d_vals, m->view.elements,
                                                                 (uint32_t)nnz, m->d_pos);
    c->launches++;
    CKP(cudaGetLastError());
  }
  CKP(cudaStreamSynchronize(c->stream));  // the host staging vectors go out of scope
  m->view.perm = m->d_perm;
  m->view.colptr = m->d_colptr;
  m->view.N = (uint32_t)N;
  m->view.nnz = (uint32_t)nnz;
  m->view.pos = m->d_pos;
  if (grouped) m->h_perm.swap(perm);
  cleanup.armed = false;
  return m;
}

void abft_cuda_coo_matrix_destroy(abft_ctx *c, abft_coo_matrix *m) {
  if (!m) return;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  cudaFree(m->view.elements);
  cudaFree(m->view.stray);
  cudaFree(m->d_perm);
  cudaFree(m->d_colptr);
  cudaFree(m->d_pos);
  delete m;
}

int abft_cuda_coo_matrix_nnz(const abft_coo_matrix *m) { return m ? (int)m->view.nnz : -1; }

int abft_cuda_coo_matrix_read(abft_ctx *c, const abft_coo_matrix *m, void *elements_out) {
  REQUIRE(c && m && elements_out, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (m->view.nnz && !m->view.pos)
    CK(cudaMemcpy(elements_out, m->view.elements, sizeof(uint4) * m->view.nnz, cudaMemcpyDeviceToHost));
  if (m->view.nnz && m->view.pos) {
    // grouped layout: slot k holds element h_perm[k] of the caller's numbering
    std::vector<uint4> tmp(m->view.nnz);
    CK(cudaMemcpy(tmp.data(), m->view.elements, sizeof(uint4) * m->view.nnz, cudaMemcpyDeviceToHost));
    uint4 *out = static_cast<uint4 *>(elements_out);
    for (uint32_t k = 0; k < m->view.nnz; k++) out[m->h_perm[k]] = tmp[k];
  }
  return 0;
}

int abft_cuda_coo_flip_bit(abft_ctx *c, abft_coo_matrix *m, int index, int bit) {
  REQUIRE(c && m, "null argument");
  REQUIRE(bit >= 0 && bit < 128, "bit out of range");
  REQUIRE(index >= 0 && (uint32_t)index < m->view.nnz, "element index out of range");
  CK(cudaSetDevice(c->device));
  coo_flip_bit_kernel<<<1, 32, 0, c->stream>>>(m->view.elements, (uint32_t)index, bit, m->view.pos);
  c->launches++;
  c->epoch++;
  CK(cudaGetLastError());
  return 0;
}

int abft_cuda_coo_spmv(abft_ctx *c, abft_coo_matrix *m, const abft_vector *vec, abft_vector *result) {
  REQUIRE(c && m && vec && result, "null argument");
  REQUIRE((uint32_t)result->N >= m->view.N && (uint32_t)vec->N >= m->view.N, "coo_spmv: vectors shorter than N");
  REQUIRE(vec->d != result->d, "coo_spmv: vec and result must not alias");
  CK(cudaSetDevice(c->device));
  bool fused = false;
  switch (c->mode) {
    case ABFT_MODE_NONE: fused = launch_coo_spmv_mode<ABFT_MODE_NONE>(c, m, vec->d, result->d); break;
    case ABFT_MODE_CONSTRAINTS: fused = launch_coo_spmv_mode<ABFT_MODE_CONSTRAINTS>(c, m, vec->d, result->d); break;
    case ABFT_MODE_SED: fused = launch_coo_spmv_mode<ABFT_MODE_SED>(c, m, vec->d, result->d); break;
    case ABFT_MODE_SEC7: fused = launch_coo_spmv_mode<ABFT_MODE_SEC7>(c, m, vec->d, result->d); break;
    case ABFT_MODE_SEC8: fused = launch_coo_spmv_mode<ABFT_MODE_SEC8>(c, m, vec->d, result->d); break;
    case ABFT_MODE_SECDED: fused = launch_coo_spmv_mode<ABFT_MODE_SECDED>(c, m, vec->d, result->d); break;
    default: return fail(-1, "coo_spmv: mode has no COO form", __FILE__, __LINE__);
  }
  c->launches += 2;
  c->epoch++;
  if (fused) {
    // dot(vec[0 .. N), result[0 .. N)) is in slot SLOT_SPMV_DOT once the stray fix-up kernel has run: a following
    // abft_cuda_dot(vec, result) (cg.cpp:97-100) launches nothing
    c->dot_a = vec->d;
    c->dot_b = result->d;
    c->dot_n = (int)m->view.N;
    c->dot_epoch = c->epoch;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
