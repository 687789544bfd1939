// Throughput of the candidate GELU building blocks on one SM-filling grid (ops per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = x[i];
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(v) : "f"(v));
      if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(v));
      if (MODE == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(v));
      if (MODE == 3) v = fmaf(v, 0.999f, 0.001f);
      x[i] = v;
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 4 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148 * 2, 1024>>>(out, 16, 0.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<148 * 2, 1024>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double ops = 148.0 * 2 * 1024 * 8.0 * iters;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("MUFU %-6s %.3f ms  %.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("tanh"); run<1>("ex2"); run<2>("rcp"); run<3>("fma");
  return 0;
}
