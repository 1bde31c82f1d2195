// gather_bench.cu -- how many random 8-byte gathers per second the B200 L2 serves.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
// This is the ceiling of an spmv whose columns are random (the power-law workload): every non-zero costs
// one 32-byte sector request for 8 useful bytes.  Threads draw indices from a hash (a few ALU ops), keep U
// independent loads in flight, 1024 threads per SM like spmv_kernel.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256, 4) gather(const double *__restrict__ x, unsigned mask, int iters, double *sink) {
  unsigned h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  double acc = 0.0;
  for (int it = 0; it < iters; it++) {
    double v[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      h = h * 1664525u + 1013904223u;
      v[k] = __ldg(x + ((h >> 7) & mask));
    }
#pragma unroll
    for (int k = 0; k < U; k++) acc += v[k];
  }
  if (acc == 123.456) sink[0] = acc;
}

template <int U>
void run(const double *x, size_t n, int sms, double *sink) {
  const int blocks = sms * 4, threads = 256, iters = 2048 / U;
  gather<U><<<blocks, threads>>>(x, (unsigned)(n - 1), iters, sink);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  gather<U><<<blocks, threads>>>(x, (unsigned)(n - 1), iters, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double g = (double)blocks * threads * iters * U;
  printf("vec %4zu MB  U=%d  %7.1f G gathers/s  (%.3f ms; %.2f per clk per SM at 1.965 GHz; sector traffic %.2f TB/s)\n",
         n * 8 >> 20, U, g / ms / 1e6, ms, g / ms / 1e6 / sms / 1.965, g * 32 / ms / 1e9);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs, L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
  double *sink;
  CK(cudaMalloc(&sink, 8));
  for (size_t mb : {16, 64, 512}) {
    const size_t n = mb << 17;
    double *x;
    CK(cudaMalloc(&x, n * 8));
    CK(cudaMemset(x, 0, n * 8));
    run<1>(x, n, p.multiProcessorCount, sink);
    run<2>(x, n, p.multiProcessorCount, sink);
    run<4>(x, n, p.multiProcessorCount, sink);
    run<8>(x, n, p.multiProcessorCount, sink);
    CK(cudaFree(x));
  }
  return 0;
}
