// microbench.cu -- instruction-issue rates that bound the fused ECC check on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
// Prints results per warp-instruction class as lane-ops per clock per SM.
// The SECDED check needs 8 POPC + ~25 LOP3 per matrix element: whether
// that fits under the HBM stream rate (~2.2 elements/clk/SM at roofline) is decided here.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;

enum Op { POPC, LOP3, IMAD, SHF, PRMT, BREV, DFMA, DADD, DMUL, IADD3, LDS32, LDS64, MIX_SECDED, FLO, IMADHI, IMADWIDE, MIX_LOP3_IMAD, MIX_LOP3_IMADHI, SHFLIDX, MIX_SHFL_LDS, MIX_SHF_SHFL };
__constant__ unsigned g_mul[2] = {1u << 19, 1u << 9};  // opaque to ptxas: mul.hi stays a multiply

template <int OP>
__global__ void bench(unsigned long long *cycles, unsigned *sink, unsigned seed) {
  __shared__ unsigned smem[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) smem[i] = i * 2654435761u;
  __syncthreads();
  unsigned x[kChains];
  double d[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) { x[c] = seed + threadIdx.x * 977 + c * 131; d[c] = 1.0 + 1e-9 * (threadIdx.x + c); }
  unsigned long long t0 = clock64();
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) {
      if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, 12345;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(g_mul[c & 1]), "r"(x[(c + 1) % kChains]));
      if (OP == IMADWIDE) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[c]), "r"(g_mul[c & 1])); x[c] = (unsigned)(w >> 32) + (unsigned)w; }
      if (OP == MIX_LOP3_IMAD) {
        asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
        asm volatile("mad.lo.u32 %0, %0, %1, 12345;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      }
      if (OP == MIX_LOP3_IMADHI) {
        asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(g_mul[c & 1]), "r"(x[(c + 1) % kChains]));
      }
      if (OP == SHFLIDX) x[c] = __shfl_sync(0xffffffffu, x[(c + 1) % kChains], (int)x[c]) + c;
      if (OP == MIX_SHFL_LDS) {
        x[c] = __shfl_sync(0xffffffffu, x[(c + 1) % kChains], (int)x[c]) + c;
        x[c] = ((volatile unsigned *)smem)[x[c] & 31u];
      }
      if (OP == MIX_SHF_SHFL) {
        asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
        x[c] = __shfl_sync(0xffffffffu, x[(c + 1) % kChains], (int)x[c]) + c;
      }
      if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      if (OP == BREV) asm volatile("brev.b32 %0, %0;" : "+r"(x[c]));
      if (OP == FLO) asm volatile("bfind.u32 %0, %0;" : "+r"(x[c]));
      if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c + 1) % kChains]));
      if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[c]) : "d"(d[(c + 1) % kChains]));
      if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(d[(c + 1) % kChains]));
      if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(d[(c + 1) % kChains]));
      if (OP == LDS32) { x[c] = ((volatile unsigned *)smem)[x[c] & 1023u]; }
      if (OP == LDS64) { const volatile unsigned long long *q = (const volatile unsigned long long *)smem; unsigned long long v = q[x[c] & 511u]; x[c] = (unsigned)v ^ (unsigned)(v >> 32); }
      if (OP == MIX_SECDED) {
        // the per-element fast path: 21 masked-xor LOP3, 8 POPC, 4 OR-combines
        unsigned w0 = x[c], w1 = x[(c + 1) % kChains], w2 = x[(c + 2) % kChains], acc = 0;
#pragma unroll
        for (int p = 0; p < 7; p++) {
          unsigned t;
          asm volatile("and.b32 %0, %1, %2;" : "=r"(t) : "r"(w0), "n"(0x56AAAD5B));
          asm volatile("lop3.b32 %0, %1, 0x2B555555, %0, 0x78;" : "+r"(t) : "r"(w1));
          asm volatile("lop3.b32 %0, %1, 0x00AAAAAA, %0, 0x78;" : "+r"(t) : "r"(w2));
          asm volatile("popc.b32 %0, %0;" : "+r"(t));
          acc |= t;
        }
        unsigned t = w0 ^ w1 ^ w2;
        asm volatile("popc.b32 %0, %0;" : "+r"(t));
        x[c] = (acc | t) + it;
      }
    }
  }
  unsigned long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += x[c] + (unsigned)d[c];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double ops_per_inner, int sms) {
  const int threads = 1024, blocks = sms * 2;  // 2048 threads per SM: full occupancy
  unsigned long long *cyc;
  unsigned *sink;
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * blocks));
  CK(cudaMalloc(&sink, sizeof(unsigned) * blocks * threads));
  bench<OP><<<blocks, threads>>>(cyc, sink, 1);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  bench<OP><<<blocks, threads>>>(cyc, sink, 2);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long h[2048];
  CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < blocks; i++) mean += (double)h[i];
  mean /= blocks;
  const double lane_ops_per_sm = 2.0 * threads * (double)kIters * kChains * ops_per_inner;  // 2 CTAs per SM
  printf("%-12s %8.2f lane-ops/clk/SM   (%.3f ms, %.0f cycles/CTA, implied clock %.0f MHz)\n", name,
         lane_ops_per_sm / mean, ms, mean, mean / (ms * 1e3));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<POPC>("POPC", 1, sms);
  run<LOP3>("LOP3", 1, sms);
  run<IADD3>("IADD", 1, sms);
  run<IMAD>("IMAD", 1, sms);
  run<IMADHI>("IMAD.HI", 1, sms);
  run<IMADWIDE>("IMAD.WIDE(+IADD)", 1, sms);
  run<MIX_LOP3_IMAD>("LOP3+IMAD", 2, sms);
  run<MIX_LOP3_IMADHI>("LOP3+IMAD.HI", 2, sms);
  run<SHFLIDX>("SHFL.IDX(+IADD)", 1, sms);
  run<MIX_SHFL_LDS>("SHFL+LDS.32", 2, sms);
  run<MIX_SHF_SHFL>("SHF+SHFL", 2, sms);
  run<SHF>("SHF", 1, sms);
  run<PRMT>("PRMT", 1, sms);
  run<BREV>("BREV", 1, sms);
  run<FLO>("FLO", 1, sms);
  run<DFMA>("DFMA", 1, sms);
  run<DADD>("DADD", 1, sms);
  run<DMUL>("DMUL", 1, sms);
  run<LDS32>("LDS.32", 1, sms);
  run<LDS64>("LDS.64", 1, sms);
  run<MIX_SECDED>("SECDED-elem", 1, sms);   // "ops" here = elements checked
  return 0;
}
