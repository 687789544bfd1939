// Cost of one tcgen05.mma kind::f16 (bf16 x bf16 -> f32) by shape and operand form, measured with clock64 around a stream
// of back-to-back MMAs issued by one thread (as the kernels of this repo do), operands resident in shared memory.
//   M = 128 (cta_group::1) and M = 256 (cta_group::2, one cluster of two CTAs per TPC); N in {64 .. 256}; K = 16;
//   operand forms: K-major / K-major (conv forward + dgrad), K-major A / MN-major B (CLIP dZ), MN-major / MN-major
//   (weight gradient).  Output: cycles per MMA and the fraction of the nominal rate (4096 MAC/clk/SM).
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I meg-decoding_b200/csrc \
//                                   scripts/umma_bench.cu -o /tmp/umma_bench && /tmp/umma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MEGB200_BENCH_STANDALONE
namespace megb200 { void set_last_error(const char*, ...) {} int num_sms() { return 148; } void count_launch() {}
struct Options; }
#include "tc_common.cuh"

using namespace megb200::tc;

struct Res { long long cycles; };

template <bool kPair>
__global__ void __launch_bounds__(192, 1) bench_kernel(int N, int a_mn, int b_mn, int reps, int two_acc, Res* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc_2sm(smem_u32(&slot), 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  }
  fence_proxy_async();
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const bool leader = !kPair || cluster_ctarank() == 0;
  if (warp == 1 && lane == 0 && leader) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32 * 1024;
    // K-major: SBO 1024, LBO unused;  MN-major: LBO = 8 KB atom stride, SBO 1024
    const uint64_t da = a_mn ? make_smem_desc(sa, 8192, 1024) : make_smem_desc(sa, 16, 1024);
    const uint64_t db = b_mn ? make_smem_desc(sb, 8192, 1024) : make_smem_desc(sb, 16, 1024);
    const uint32_t idesc = make_idesc_m(kPair ? 256 : 128, N, a_mn, b_mn);
    const uint32_t ka = a_mn ? kUmmaK * 128 : 32, kb = b_mn ? kUmmaK * 128 : 32;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tmem_base + (uint32_t)((two_acc && (k & 1)) ? 256 : 0);
        if (kPair) umma_bf16_2sm(d, desc_advance(da, k * ka), desc_advance(db, k * kb), idesc, 1u);
        else umma_bf16(d, desc_advance(da, k * ka), desc_advance(db, k * kb), idesc, 1u);
      }
    }
    if (kPair) umma_commit_2sm(smem_u32(&bar), 1); else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x].cycles = t1 - t0;
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <bool kPair>
static double run(int N, int a_mn, int b_mn, int two_acc) {
  Res* out;
  const int blocks = 148;
  cudaMalloc(&out, sizeof(Res) * blocks);
  cudaMemset(out, 0, sizeof(Res) * blocks);
  const int reps = 256;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(bench_kernel<kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int w = 0; w < 2; ++w) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<kPair>, N, a_mn, b_mn, reps, two_acc, out);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(e));
      return -1;
    }
  }
  Res host[148];
  cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
  double sum = 0;
  int n = 0;
  for (int i = 0; i < blocks; ++i)
    if (host[i].cycles > 0) { sum += (double)host[i].cycles; ++n; }
  cudaFree(out);
  return n ? sum / n / (reps * 4.0) : -1;
}

int main() {
  const char* forms[3] = {"K-major A, K-major B  ", "K-major A, MN-major B ", "MN-major A, MN-major B"};
  const int fa[3] = {0, 0, 1}, fb[3] = {0, 1, 1};
  const int Ns[] = {64, 96, 128, 160, 192, 256};
  printf("cycles per tcgen05.mma (K = 16), all SMs issuing; ideal = M*N*16 / (4096 MAC/clk/SM * SMs in the group)\n");
  for (int f = 0; f < 3; ++f)
    for (int pair = 0; pair < 2; ++pair)
      for (int two = 0; two < 2; ++two) {
        printf("%s %s %s:", forms[f], pair ? "M=256 cta_group::2" : "M=128 cta_group::1", two ? "two accumulators " : "one accumulator  ");
        for (int N : Ns) {
          if (pair && (N % 32) != 0) { printf("   N=%3d    -   ", N); continue; }
          const double c = pair ? run<true>(N, fa[f], fb[f], two) : run<false>(N, fa[f], fb[f], two);
          const double ideal = 128.0 * N * 16 / 4096.0;
          printf("  N=%3d %6.1f (%3.0f%%)", N, c, 100.0 * ideal / c);
        }
        printf("\n");
      }
  return 0;
}
