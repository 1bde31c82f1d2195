// emu_runtime.h -- a fake CUDA runtime for the host side of csrc/abft_cuda.cu (test infrastructure): "device" memory is
// host memory, streams are synchronous, one device with two SMs, kernels run through emu::launch (cuda_runtime.h in
// this directory).  tests/helpers/build_emu_lib.py rewrites the `kernel<<<grid, block, smem, stream>>>(args);` launches
// of a temporary copy of abft_cuda.cu into EMU_LAUNCH(...) and compiles the result with g++ into libabftcuda_emu.so, so
// that the C ABI -- host orchestration included -- can be driven by the GPU tests' own code on a machine without a GPU
// (small cases; no peer memory, so no multi-GPU entry points).  It sees logic only.
#pragma once
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600, cudaErrorNotSupported = 801, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct emu_stream_t *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaHostAllocPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int multiProcessorCount; };
struct cudaIpcMemHandle_t { char reserved[64]; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated runtime: not supported"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 2; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void *)(sym), (src), (n)), cudaSuccess)
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorNotSupported; }

// kernel<<<grid, block, smem, stream>>>(args);  ->  EMU_LAUNCH(grid, block, smem, kernel(args));
#define EMU_LAUNCH(grid, block, smem, ...)                                                  \
  do {                                                                                       \
    const size_t emu_bytes_ = (size_t)(smem);                                                \
    unsigned char *emu_buf_ = emu_bytes_ ? (unsigned char *)aligned_alloc(128, (emu_bytes_ + 127) / 128 * 128) : nullptr; \
    emu::g_dyn_smem = emu_buf_;                                                              \
    emu::launch((int)(grid), (int)(block), [&] { __VA_ARGS__; });                               \
    emu::g_dyn_smem = nullptr;                                                               \
    free(emu_buf_);                                                                          \
  } while (0)
