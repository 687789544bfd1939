// Shared helpers for the megb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/megb200.h"

namespace megb200 {

// ------------------------------------------------------------------------------------------ error plumbing
void set_last_error(const char* fmt, ...);

#define MEGB_CHECK_ARG(cond, ...)                              \
  do {                                                         \
    if (!(cond)) {                                             \
      ::megb200::set_last_error(__VA_ARGS__);                  \
      return MEGB200_ERR_ARG;                                  \
    }                                                          \
  } while (0)

#define MEGB_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::megb200::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                        \
    }                                                                                        \
  } while (0)

#define MEGB_LAUNCH_OK()                                                                     \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::megb200::set_last_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                        \
    }                                                                                        \
    ::megb200::count_launch();                                                               \
  } while (0)

int num_sms();
void count_launch();

// Kernel-variant selection (megb200_set_option).  Defaults come from the MEGB200_* environment variables of the same
// meaning, read once; tests flip them in-process so that every variant is reachable on small shapes.
struct Options {
  long long tapgemm_pair_min_items;   // work items from which the CTA-pair tap-GEMM is used; -1 = 2 x SM count
  int tapgemm_force_1cta;             // 1: always the single-CTA tap-GEMM
  int tapgemm_pair_halo;              // 0: plain pair kernel for k > 1 instead of the halo-staged one
  int tapgemm_halo;                   // 1: single-CTA halo form for k > 1 (opt-in experiment)
  int tapgemm_ragged_n;               // 1: 256-wide N tiles + narrower tail
  int tapgemm_direct_epi;             // 1: per-thread global stores instead of the TMA-store epilogue
  int tapgemm_l2_prefetch;            // 1: the pair tap-GEMM producers prefetch their next tile's activations into L2 (default 0)
  int tapgemm_ksub;                   // 1: one k chunk per stage in the plain pair kernel (deeper ring; default: two)
  int clip_dz_large;                  // 1: always the large-K dZ kernel (default: small-K form when By <= 512)
  int rowloop_direct;                 // 1: direct-load streaming kernels instead of the bulk-copy staged ones
  int wgrad_splits;                   // > 0: force the split count of the weight-gradient GEMM
  int wgrad_pair;                     // 1: CTA-pair weight-gradient kernel for ungrouped problems (default 0: measured slower)
  int clip_dz_jt;                     // > 0: j tiles (128 brain rows) per CTA of the large-K dZ kernel (default: automatic)
  int clip_fused;                     // host-side policy read by the Python engine: -1 auto (large batch, short
                                      // embeddings), 0 never, 1 whenever supported -- the never-materialised CLIP form
  // written by the launchers (read back with megb200_get_option; the tests assert which form actually ran)
  int last_tapgemm_kernel;            // 0 = single CTA, 1 = CTA pair, 2 = CTA pair + halo, 3 = single CTA + halo
  int last_clip_dz_kernel;            // 0 = small-K (stream-bound) form, 1 = large-K form
  int last_wgrad_splits;
  int last_wgrad_kernel;              // 0 = single CTA, 1 = CTA pair
};
Options& options();

// ------------------------------------------------------------------------------------------ dtype helpers
template <typename T> struct Vec8;              // 8 consecutive channels
template <> struct Vec8<float> {
  float4 a, b;
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
};

__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&x)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&x)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// value after one rounding to the storage type (what the next kernel will read back)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f(from_f<T>(v)); }

// ------------------------------------------------------------------------------------------ math
// exact (erf) GELU, as torch.nn.functional.gelu default (reference models.py:314,317,373,374)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// MUFU.RCP (1 ulp-ish): the IEEE reciprocal / division expands to ~10 instructions plus a slow path, which makes the
// streaming passes ALU-bound instead of HBM-bound
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_f(float x) { return fast_rcp(1.0f + __expf(-x)); }

// Fast erf-GELU for the bf16 path: Phi(x) = 0.5 (1 + tanh(g(x))) with g the odd quintic least-max-error fit of
// atanh(erf(x / sqrt 2)) (max |Phi error| 1.9e-5 on the real line once x is clamped to |x| <= 8, where tanh has saturated;
// tanh.approx adds <= 2^-11 relative) -- an order below the bf16 rounding of the stored activation.  One MUFU.TANH and
// six FP32 ops instead of erff (the Abramowitz-Stegun form this replaced took 2 MUFU + 16 ops and made the streaming
// passes issue-bound: ncu showed 41 % issue utilisation at 34 % of DRAM peak).  The check mode (f32) keeps erff.
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_cdf_fast(float x, float& x2) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  x2 = xc * xc;
  const float g = xc * fmaf(x2, fmaf(x2, -3.16786541e-4f, 3.68025929e-2f), 7.97715065e-1f);
  return fmaf(0.5f, tanh_approx(g), 0.5f);
}
// storage-type dispatch: exact erff in the f32 check mode, fast form for bf16
template <typename T> __device__ __forceinline__ float gelu_t(float x) { return gelu_f(x); }
template <> __device__ __forceinline__ float gelu_t<__nv_bfloat16>(float x) {
  float x2;
  return x * gelu_cdf_fast(x, x2);
}
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) { return gelu_grad_f(x); }
template <> __device__ __forceinline__ float gelu_grad_t<__nv_bfloat16>(float x) {
  float x2;
  const float cdf = gelu_cdf_fast(x, x2);
  const float pdf = 0.3989422804014327f * fast_ex2(-0.72134752044448170f * x2);   // exp(-x^2/2); MUFU.EX2
  return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace megb200
