// micro-benchmark: MUFU tanh throughput, f32 vs packed bf16x2 / f16x2 (lane-ops per clock per SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8];
  for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x * 8 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(x[j]));
      if (MODE == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x[j]));
      if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x[j]));
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[j]));
      if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[j]));
      if (MODE == 5) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(x[j]));
      if (MODE == 6) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(x[j]));
      if (MODE == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(x[j]));
    }
  }
  uint32_t s = 0;
  for (int j = 0; j < 8; ++j) s ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, iters, 1);
  cudaEventRecord(a);
  k<MODE><<<148, 1024>>>(out, iters, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = 148.0 * 1024 * 8.0 * iters;   // lane-instructions
  printf("%-24s %.3f ms  %.2f lane-instr/ns chip  (%.1f per SM per clk @1.9GHz)\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}
int main() {
  run<0>("tanh.f32"); run<1>("tanh.bf16x2"); run<2>("tanh.f16x2"); run<3>("ex2.f32"); run<4>("ex2.bf16x2");
  run<5>("fma.bf16x2"); run<6>("fma.f32"); run<7>("rcp.f32");
  return 0;
}
