// fp32 CUDA-core contraction engine ("check mode", MEGB200_SIMT): the same contractions as the tcgen05 engine in
// tc_gemm.cu, accumulated in fp32 FMA on any storage dtype.  It is (a) the fp32 check mode the parity contract asks
// for (1e-3 vs the fp32 reference) and (b) the on-device cross-check for every tensor-core kernel.  One 64x64x16
// register-tiled core parameterised by operand loaders and an epilogue.
#include "common.cuh"

namespace megb200 {

constexpr int kBM = 64, kBN = 64, kBK = 16;

// kAK / kBK_: true when the operand's k index is the contiguous one in memory (keeps loads coalesced)
// kGluCols: the 64-column tile is [32 "a" channels | their 32 gate channels]; a thread then owns 2 a-columns
// (j = 0,1 -> 2*tx + j) and the 2 matching gate columns (j = 2,3 -> 32 + 2*tx + j - 2) so that GLU is thread-local.
template <bool kAKMajor, bool kBKMajor, bool kGluCols = false, class LoadA, class LoadB>
__device__ __forceinline__ void simt_tile(const LoadA& la, const LoadB& lb, int k0, int k1, float (&acc)[4][4]) {
  __shared__ float As[kBK][kBM + 1];
  __shared__ float Bs[kBK][kBN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = k0; k < k1; k += kBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      {
        const int r = kAKMajor ? (e >> 4) : (e & 63);
        const int kk = kAKMajor ? (e & 15) : (e >> 6);
        As[kk][r] = (k + kk < k1) ? la(r, k + kk) : 0.f;
      }
      {
        const int r = kBKMajor ? (e >> 4) : (e & 63);
        const int kk = kBKMajor ? (e & 15) : (e >> 6);
        Bs[kk][r] = (k + kk < k1) ? lb(r, k + kk) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][kGluCols ? ((j < 2) ? (tx * 2 + j) : (32 + tx * 2 + j - 2)) : (tx * 4 + j)];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ tap GEMM
template <typename T, int kEpi>
__global__ void __launch_bounds__(256)
simt_tapgemm_kernel(const T* __restrict__ A, int Tn, int Ca, int Cas, const T* __restrict__ W, int N, int Np, int Kp,
                    int taps, int dil, const float* __restrict__ bias, const T* __restrict__ Res, int Crs,
                    const int32_t* __restrict__ group, void* __restrict__ OutV, int Cos, T* __restrict__ Out2, int Co2s,
                    int tiles_per_sample) {
  constexpr bool kGlu = (kEpi == MEGB200_EPI_GLU);
  const int b = blockIdx.x / tiles_per_sample;
  const int t0 = (blockIdx.x % tiles_per_sample) * kBM;
  const int Ch = N / 2;                                   // GLU only
  const int n0 = blockIdx.y * (kGlu ? kBN / 2 : kBN);     // first output channel of this tile
  const int g = group ? group[b] : 0;
  const T* Ab = A + (long long)b * Tn * Cas;
  const T* Wg = W + (long long)g * Np * taps * Kp;
  const int half = taps / 2;
  auto la = [&](int r, int k) -> float {
    const int j = k / Kp, c = k - j * Kp;
    const int t = t0 + r + (j - half) * dil;
    if (c >= Ca || t < 0 || t >= Tn || t0 + r >= Tn) return 0.f;
    return to_f(Ab[(long long)t * Cas + c]);
  };
  auto lb = [&](int r, int k) -> float {
    int n;
    if (kGlu) {
      const int c = n0 + (r & 31);
      if (c >= Ch) return 0.f;
      n = (r < 32) ? c : Ch + c;
    } else {
      n = n0 + r;
      if (n >= N) return 0.f;
    }
    return to_f(Wg[(long long)n * taps * Kp + k]);
  };
  float acc[4][4];
  simt_tile<true, true, kGlu>(la, lb, 0, taps * Kp, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T* Out = reinterpret_cast<T*>(OutV);
  float* OutF = reinterpret_cast<float*>(OutV);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= Tn) continue;
    const long long row = (long long)b * Tn + t;
    if (kGlu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = n0 + tx * 2 + j;
        if (c < Ch) {
          const float a = acc[i][j] + (bias ? bias[(long long)g * N + c] : 0.f);
          const float gt = acc[i][j + 2] + (bias ? bias[(long long)g * N + Ch + c] : 0.f);
          Out2[row * Co2s + c] = from_f<T>(a);
          Out2[row * Co2s + Ch + c] = from_f<T>(gt);
          Out[row * Cos + c] = from_f<T>(a * sigmoid_f(gt));
        } else if (c < Cos) {
          Out[row * Cos + c] = from_f<T>(0.f);
        }
      }
      if (blockIdx.y == 0 && tx == 0)
        for (int c = N; c < Co2s; ++c) Out2[row * Co2s + c] = from_f<T>(0.f);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Cos) continue;
      float v = 0.f;
      if (n < N) {
        v = acc[i][j];
        if (bias) v += bias[(long long)g * N + n];
        if (Res) v += to_f(Res[row * Crs + n]);
      }
      if (kEpi == MEGB200_EPI_STORE_F32) {
        OutF[row * Cos + n] = v;
      } else if (kEpi == MEGB200_EPI_GELU) {
        if (n < Co2s) Out2[row * Co2s + n] = from_f<T>(v);
        Out[row * Cos + n] = from_f<T>(n < N ? gelu_t<T>(v) : 0.f);
      } else {
        Out[row * Cos + n] = from_f<T>(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ weight gradient
// block = (m tile, c tile) x tap x sample-range; fp32 register accumulation over the samples of one group run
template <typename T>
__global__ void __launch_bounds__(256)
simt_wgrad_kernel(const T* __restrict__ dU, int M, int Ms, const T* __restrict__ Xin, int C, int Cs, int B, int Tn,
                  int taps, int dil, const int32_t* __restrict__ group, float* __restrict__ dW, int ctiles,
                  int samples_per_block) {
  const int m0 = (blockIdx.x / ctiles) * kBM;
  const int c0 = (blockIdx.x % ctiles) * kBN;
  const int j = blockIdx.y;
  const int b0 = blockIdx.z * samples_per_block;
  int b1 = b0 + samples_per_block;
  if (b1 > B) b1 = B;
  const int off = (j - taps / 2) * dil;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) tot[i][q] = 0.f;
  auto flush = [&](int g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + tx * 4 + q;
        if (m < M && c < C && tot[i][q] != 0.f)
          atomicAdd(&dW[(((long long)g * M + m) * C + c) * taps + j], tot[i][q]);
        tot[i][q] = 0.f;
      }
    }
  };
  int cur = group ? group[b0] : 0;
  for (int b = b0; b < b1; ++b) {
    const int g = group ? group[b] : 0;
    if (g != cur) {
      flush(cur);
      cur = g;
    }
    const T* Ub = dU + (long long)b * Tn * Ms;
    const T* Xb = Xin + (long long)b * Tn * Cs;
    auto la = [&](int r, int k) -> float {  // A[m][k=t]
      const int m = m0 + r;
      return (m < M) ? to_f(Ub[(long long)k * Ms + m]) : 0.f;
    };
    auto lb = [&](int r, int k) -> float {  // B[c][k=t] = Xin[t+off][c]
      const int c = c0 + r, t = k + off;
      return (c < C && t >= 0 && t < Tn) ? to_f(Xb[(long long)t * Cs + c]) : 0.f;
    };
    float acc[4][4];
    simt_tile<false, false>(la, lb, 0, Tn, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) tot[i][q] += acc[i][q];
  }
  flush(cur);
}

// ------------------------------------------------------------------------------------------ small f32 GEMM
// (16*TM)x(16*TM)x8 tile, 256 threads, TMxTM outputs per thread.  TM = 8 (128x128: two float4 shared loads per operand
// and k, FMA-bound) for problems with enough tiles to fill the machine; TM = 4 (64x64) for the subject-chain products
// whose whole output is a few 128-tiles (270 x 208: 6) -- there the time is one CTA's serial k loop, so the launcher
// trades tile efficiency for parallelism and adds deterministic split-K through the workspace.
constexpr int kSgPad = 4;

template <int TM, int kSgK>
__global__ void __launch_bounds__(256)
simt_sgemm_kernel(int transA, int transB, int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                  long long sA, const float* __restrict__ Bm, int ldb, long long sB, float beta,
                  float* __restrict__ Cm, int ldc, long long sC, int splits, int kchunk, float* __restrict__ part) {
  constexpr int kTile = 16 * TM;                 // rows and columns per CTA
  constexpr int kPer = kTile * kSgK / 256;       // operand elements per thread and k step
  __shared__ __align__(16) float As[kSgK][kTile + kSgPad];
  __shared__ __align__(16) float Bs[kSgK][kTile + kSgPad];
  const int batch = blockIdx.z / splits, split = blockIdx.z % splits;
  const int m0 = blockIdx.x * kTile, n0 = blockIdx.y * kTile;
  const float* Ab = A + batch * sA;
  const float* Bb = Bm + batch * sB;
  float* Cb = Cm + batch * sC;
  const int k_begin = split * kchunk;
  const int k_end = min(K, k_begin + kchunk);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
  // global -> register prefetch of the next k tile while the current one is multiplied (without it every k step
  // exposes a full global-memory latency)
  float ra[kPer], rb[kPer];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = tid + i * 256;
      {
        // transposed A is m-contiguous in memory, plain A is k-contiguous: keep the fast index on the lanes
        const int m = transA ? (e % kTile) : (e / kSgK);
        const int kk = transA ? (e / kTile) : (e % kSgK);
        const int gm = m0 + m, gk = k0 + kk;
        float v = 0.f;
        if (gm < M && gk < k_end) v = transA ? Ab[(long long)gk * lda + gm] : Ab[(long long)gm * lda + gk];
        ra[i] = v;
      }
      {
        const int n = transB ? (e / kSgK) : (e % kTile);
        const int kk = transB ? (e % kSgK) : (e / kTile);
        const int gn = n0 + n, gk = k0 + kk;
        float v = 0.f;
        if (gn < N && gk < k_end) v = transB ? Bb[(long long)gn * ldb + gk] : Bb[(long long)gk * ldb + gn];
        rb[i] = v;
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = tid + i * 256;
      As[transA ? (e / kTile) : (e % kSgK)][transA ? (e % kTile) : (e / kSgK)] = ra[i];
      Bs[transB ? (e % kSgK) : (e / kTile)][transB ? (e / kSgK) : (e % kTile)] = rb[i];
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += kSgK) {
    stash();
    __syncthreads();
    if (k0 + kSgK < k_end) fetch(k0 + kSgK);
#pragma unroll
    for (int kk = 0; kk < kSgK; ++kk) {
      float a[TM], b[TM];
#pragma unroll
      for (int v = 0; v < TM / 4; ++v) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4 * v]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TM + 4 * v]);
        a[4 * v] = av.x; a[4 * v + 1] = av.y; a[4 * v + 2] = av.z; a[4 * v + 3] = av.w;
        b[4 * v] = bv.x; b[4 * v + 1] = bv.y; b[4 * v + 2] = bv.z; b[4 * v + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int n = n0 + tx * TM + j;
      if (n >= N) continue;
      float* dst = Cb + (long long)m * ldc + n;
      if (splits > 1) {
        // deterministic split-K: plain stores of the partial sums, combined in a fixed order by sgemm_reduce_kernel
        part[(((long long)split * (gridDim.z / splits) + batch) * M + m) * N + n] = acc[i][j];
      } else {
        float v = alpha * acc[i][j];
        if (beta != 0.f) v += beta * *dst;
        *dst = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ CLIP contractions
template <typename T>
__global__ void __launch_bounds__(256)
simt_clip_logits_kernel(const T* __restrict__ Yp, const T* __restrict__ Zp, float* __restrict__ raw, int By, int Bz,
                        long long D, long long kchunk) {
  const int i0 = blockIdx.x * kBM, j0 = blockIdx.y * kBN;
  const long long k0 = blockIdx.z * kchunk;
  long long k1 = k0 + kchunk;
  if (k1 > D) k1 = D;
  const T* Ya = Yp + k0;
  const T* Za = Zp + k0;
  auto la = [&](int r, int k) -> float { return (i0 + r < By) ? to_f(Ya[(long long)(i0 + r) * D + k]) : 0.f; };
  auto lb = [&](int r, int k) -> float { return (j0 + r < Bz) ? to_f(Za[(long long)(j0 + r) * D + k]) : 0.f; };
  float acc[4][4];
  simt_tile<true, true>(la, lb, 0, (int)(k1 - k0), acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = i0 + ty * 4 + i, c = j0 + tx * 4 + j;
      if (r < By && c < Bz) atomicAdd(&raw[(long long)r * Bz + c], acc[i][j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
simt_clip_dz_kernel(const T* __restrict__ coef, int ldc, const T* __restrict__ Yp, const T* __restrict__ Z,
                    const float* __restrict__ cvec, const float* __restrict__ gscale, T* __restrict__ dZ, int By,
                    int Bz, long long D) {
  const int j0 = blockIdx.y * kBM;
  const long long d0 = (long long)blockIdx.x * kBN;
  auto la = [&](int r, int k) -> float { return (j0 + r < Bz) ? to_f(coef[(long long)(j0 + r) * ldc + k]) : 0.f; };
  auto lb = [&](int r, int k) -> float { return (d0 + r < D) ? to_f(Yp[(long long)k * D + d0 + r]) : 0.f; };
  float acc[4][4];
  simt_tile<true, false>(la, lb, 0, By, acc);
  const float gs = gscale[0];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= Bz) continue;
    const float cj = cvec[j];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long d = d0 + tx * 4 + q;
      if (d >= D) continue;
      const float z = to_f(Z[(long long)j * D + d]);
      dZ[(long long)j * D + d] = from_f<T>(gs * (acc[i][q] - z * cj));
    }
  }
}

// ------------------------------------------------------------------------------------------ launchers
template <typename T>
int simt_tapgemm(const void* A, int B, int Tn, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps,
                 int dil, const float* bias, const void* Res, int Crs, const int32_t* group, int epi, void* Out,
                 int Cos, void* Out2, int Co2s, cudaStream_t st) {
  const int tps = (Tn + kBM - 1) / kBM;
  const int cover = (epi == MEGB200_EPI_GLU) ? (Cos + kBN / 2 - 1) / (kBN / 2) : (Cos + kBN - 1) / kBN;
  dim3 grid(B * tps, cover);
#define MEGB_SIMT_TAP(E)                                                                                          \
  simt_tapgemm_kernel<T, E><<<grid, 256, 0, st>>>((const T*)A, Tn, Ca, Cas, (const T*)W, N, Np, Kp, taps, dil, bias, \
                                                   (const T*)Res, Crs, group, Out, Cos, (T*)Out2, Co2s, tps)
  switch (epi) {
    case MEGB200_EPI_STORE: MEGB_SIMT_TAP(MEGB200_EPI_STORE); break;
    case MEGB200_EPI_STORE_F32: MEGB_SIMT_TAP(MEGB200_EPI_STORE_F32); break;
    case MEGB200_EPI_GELU: MEGB_SIMT_TAP(MEGB200_EPI_GELU); break;
    case MEGB200_EPI_GLU: MEGB_SIMT_TAP(MEGB200_EPI_GLU); break;
    default: MEGB_CHECK_ARG(false, "tapgemm: unknown epilogue %d", epi);
  }
#undef MEGB_SIMT_TAP
  MEGB_LAUNCH_OK();
  return 0;
}
template int simt_tapgemm<float>(const void*, int, int, int, int, const void*, int, int, int, int, int, const float*,
                                 const void*, int, const int32_t*, int, void*, int, void*, int, cudaStream_t);
template int simt_tapgemm<__nv_bfloat16>(const void*, int, int, int, int, const void*, int, int, int, int, int,
                                         const float*, const void*, int, const int32_t*, int, void*, int, void*, int,
                                         cudaStream_t);

template <typename T>
int simt_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int Tn, int taps, int dil,
               const int32_t* group, float* dW, cudaStream_t st) {
  const int mt = (M + kBM - 1) / kBM, ct = (C + kBN - 1) / kBN;
  int spb = 1;
  if (group == nullptr) {
    // enough CTAs to fill the machine, long enough runs to amortise the atomics
    const long long base = (long long)mt * ct * taps;
    long long want = ((long long)num_sms() * 4 + base - 1) / base;
    if (want < 1) want = 1;
    spb = (int)((B + want - 1) / want);
    if (spb < 1) spb = 1;
  } else {
    spb = 4;
  }
  dim3 grid(mt * ct, taps, (B + spb - 1) / spb);
  MEGB_CHECK_ARG(grid.z <= 65535, "wgrad: too many sample blocks");
  simt_wgrad_kernel<T><<<grid, 256, 0, st>>>((const T*)dU, M, Ms, (const T*)Xin, C, Cs, B, Tn, taps, dil, group, dW,
                                              ct, spb);
  MEGB_LAUNCH_OK();
  return 0;
}
template int simt_wgrad<float>(const void*, int, int, const void*, int, int, int, int, int, int, const int32_t*,
                               float*, cudaStream_t);
template int simt_wgrad<__nv_bfloat16>(const void*, int, int, const void*, int, int, int, int, int, int,
                                       const int32_t*, float*, cudaStream_t);

template <typename T>
int simt_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, cudaStream_t st) {
  const int it = (By + kBM - 1) / kBM, jt = (Bz + kBN - 1) / kBN;
  long long splits = ((long long)num_sms() * 4 + (long long)it * jt - 1) / ((long long)it * jt);
  long long kchunk = (D + splits - 1) / splits;
  kchunk = ((kchunk + 255) / 256) * 256;
  if (kchunk > (1ll << 30)) kchunk = 1ll << 30;
  splits = (D + kchunk - 1) / kchunk;
  MEGB_CHECK_ARG(splits <= 65535, "clip_logits: too many k splits");
  dim3 grid(it, jt, (int)splits);
  simt_clip_logits_kernel<T><<<grid, 256, 0, st>>>((const T*)Yp, (const T*)Zp, raw, By, Bz, D, kchunk);
  MEGB_LAUNCH_OK();
  return 0;
}
template int simt_clip_logits<float>(const void*, const void*, float*, int, int, long long, cudaStream_t);
template int simt_clip_logits<__nv_bfloat16>(const void*, const void*, float*, int, int, long long, cudaStream_t);

template <typename T>
int simt_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
                 void* dZ, int By, int Bz, long long D, cudaStream_t st) {
  const long long dt = (D + kBN - 1) / kBN;
  MEGB_CHECK_ARG(dt <= 2147483647ll && (Bz + kBM - 1) / kBM <= 65535, "clip_dz: grid too large");
  dim3 grid((unsigned)dt, (Bz + kBM - 1) / kBM);
  simt_clip_dz_kernel<T><<<grid, 256, 0, st>>>((const T*)coef, ldc, (const T*)Yp, (const T*)Z, cvec, gscale, (T*)dZ, By,
                                                Bz, D);
  MEGB_LAUNCH_OK();
  return 0;
}
template int simt_clip_dz<float>(const void*, int, const void*, const void*, const float*, const float*, void*, int,
                                 int, long long, cudaStream_t);
template int simt_clip_dz<__nv_bfloat16>(const void*, int, const void*, const void*, const float*, const float*,
                                         void*, int, int, long long, cudaStream_t);

}  // namespace megb200

using namespace megb200;

__global__ void sgemm_reduce_kernel(const float* __restrict__ part, int splits, long long per_split, int M, int N,
                                    float alpha, float beta, float* __restrict__ Cm, int ldc, long long sC) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_split;
       idx += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[k * per_split + idx];
    const int n = (int)(idx % N);
    const long long r = idx / N;
    const int m = (int)(r % M);
    const long long b = r / M;
    float* dst = Cm + b * sC + (long long)m * ldc + n;
    float v = alpha * s;
    if (beta != 0.f) v += beta * *dst;
    *dst = v;
  }
}

extern "C" int megb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                             long long strideA, const float* Bm, int ldb, long long strideB, float beta, float* C,
                             int ldc, long long strideC, int batch, float* workspace, size_t ws_bytes, void* stream) {
  MEGB_CHECK_ARG(A && Bm && C && M > 0 && N > 0 && K > 0 && batch > 0, "sgemm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // tile: 64 unless the problem is large (the hot path's products are all small: at most 7290 x 270 x 208); the 64-tile
  // kernel co-resides 6-8 CTAs per SM, which hides the k loop's global-load latency
  const long long tiles64 = (long long)((M + 63) / 64) * ((N + 63) / 64) * batch;
  const int tile = tiles64 >= 16ll * num_sms() ? 128 : 64;
  const int mt = (M + tile - 1) / tile, nt = (N + tile - 1) / tile;
  MEGB_CHECK_ARG(nt <= 65535, "sgemm: N too large");
  // split-K (through the workspace, deterministic: fixed k ranges, summed in order) until about six CTAs per SM exist
  int splits = 1;
  const long long tiles = (long long)mt * nt * batch;
  const long long per_split = (long long)batch * M * N;
  if (workspace != nullptr && tiles < 6ll * num_sms() && K >= 32) {
    splits = (int)((6ll * num_sms() + tiles - 1) / tiles);
    if (splits > K / 32) splits = K / 32;
    const long long fit = (long long)(ws_bytes / sizeof(float)) / per_split;
    if (splits > fit) splits = (int)fit;
    if (splits < 1) splits = 1;
  }
  int kchunk = (K + splits - 1) / splits;
  const int bk = tile == 128 ? 8 : 32;       // k elements per step: long steps amortise the global-load round trip
  kchunk = (kchunk + bk - 1) / bk * bk;
  splits = (K + kchunk - 1) / kchunk;
  MEGB_CHECK_ARG((long long)batch * splits <= 65535, "sgemm: batch too large");
  dim3 grid(mt, nt, batch * splits);
  if (tile == 128)
    simt_sgemm_kernel<8, 8><<<grid, 256, 0, st>>>(transA, transB, M, N, K, alpha, A, lda, strideA, Bm, ldb, strideB, beta, C,
                                               ldc, strideC, splits, kchunk, workspace);
  else
    simt_sgemm_kernel<4, 32><<<grid, 256, 0, st>>>(transA, transB, M, N, K, alpha, A, lda, strideA, Bm, ldb, strideB, beta, C,
                                               ldc, strideC, splits, kchunk, workspace);
  MEGB_LAUNCH_OK();
  if (splits > 1) {
    long long blocks = (per_split + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    sgemm_reduce_kernel<<<(int)blocks, 256, 0, st>>>(workspace, splits, per_split, M, N, alpha, beta, C, ldc, strideC);
    MEGB_LAUNCH_OK();
  }
  return 0;
}
