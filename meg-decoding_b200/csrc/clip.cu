// CLIPLoss softmax / cross-entropy stage (reference meg_decoding/utils/loss.py:92-107) on the f32 dot-product
// matrix produced by the logits contraction.  Rows = speech segments (x = Y), columns = brain embeddings (y = Z),
// exactly as train.py:203 calls loss_func(Y, Z).  Written in the one-sided data-parallel form of SURVEY appendix A6:
// a rank holds all By speech rows and its Bz local brain columns; its positives sit at row off + j.
#include "common.cuh"

namespace megb200 {

// one warp per speech row: L = raw * s / (|y_i||z_j|), partial (max, sum exp) over the local brains
__global__ void clip_rows_kernel(const float* __restrict__ raw, const float* __restrict__ ysq,
                                 const float* __restrict__ zsq, const float* __restrict__ temp, float* __restrict__ L,
                                 float* __restrict__ row_max, float* __restrict__ row_sumexp, int By, int Bz) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= By) return;
  const float s = expf(temp[0]);
  const float ri = s * rsqrtf(ysq[warp]);
  const float* rr = raw + (long long)warp * Bz;
  float* lr = L + (long long)warp * Bz;
  float mx = -INFINITY;
  for (int j = lane; j < Bz; j += 32) {
    const float v = rr[j] * ri * rsqrtf(zsq[j]);
    lr[j] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float se = 0.f;
  for (int j = lane; j < Bz; j += 32) se += expf(lr[j] - mx);   // own writes, same thread
  se = warp_sum(se);
  if (lane == 0) {
    row_max[warp] = mx;
    row_sumexp[warp] = se;
  }
}

// block = 32 columns x 32 row lanes: complete log-sum-exp over all speech rows for each local brain column.  Every thread
// keeps FOUR independent online (max, sum exp) pairs over interleaved rows so that four loads are in flight per
// iteration -- the first form (8 row lanes, one dependent chain per thread) ran at one L2 round trip per row: 67 us for
// the 2048 x 256 logits of an 8-GPU rank, on the critical path between the two CLIP contractions.
constexpr int kColsLanes = 32;
__global__ void __launch_bounds__(32 * kColsLanes)
clip_cols_kernel(const float* __restrict__ L, float* __restrict__ col_lse, int By, int Bz) {
  __shared__ float smx[kColsLanes][33], sse[kColsLanes][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, se[4] = {0.f, 0.f, 0.f, 0.f};
  if (j < Bz) {
    for (int i0 = threadIdx.y; i0 < By; i0 += 4 * kColsLanes) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kColsLanes;
        v[u] = i < By ? L[(long long)i * Bz + j] : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v[u] > mx[u]) {
          se[u] = se[u] * expf(mx[u] - v[u]) + 1.f;      // mx = -inf: se = 0 and exp(-inf) = 0
          mx[u] = v[u];
        } else if (v[u] > -INFINITY) {
          se[u] += expf(v[u] - mx[u]);
        }
      }
    }
  }
  float M = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  float S = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (mx[u] > -INFINITY) S += se[u] * expf(mx[u] - M);
  smx[threadIdx.y][threadIdx.x] = M;
  sse[threadIdx.y][threadIdx.x] = S;
  __syncthreads();
  if (threadIdx.y == 0 && j < Bz) {
    float Mt = smx[0][threadIdx.x];
    for (int r = 1; r < kColsLanes; ++r) Mt = fmaxf(Mt, smx[r][threadIdx.x]);
    float St = 0.f;
    for (int r = 0; r < kColsLanes; ++r) {
      const float m = smx[r][threadIdx.x];
      if (m > -INFINITY) St += sse[r][threadIdx.x] * expf(m - Mt);
    }
    col_lse[j] = Mt + logf(St);
  }
}

// Complete the per-speech-row log-sum-exp over ALL brains from the ranks' partial (max, sum exp) pairs (SURVEY appendix
// A6: in data parallel a rank only sees its own brain columns).  parts: [W][2][By] -- rank-major, row 0 = max, row 1 =
// sum of exp(L - max) -- as an all-gather of the ranks' [2][By] blocks lays them out; W = 1 is the single-process case.
__global__ void clip_lse_merge_kernel(const float* __restrict__ parts, int W, int By, float* __restrict__ row_lse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= By) return;
  float M = -INFINITY;
  for (int w = 0; w < W; ++w) M = fmaxf(M, parts[((long long)w * 2) * By + i]);
  float S = 0.f;
  for (int w = 0; w < W; ++w) {
    const float m = parts[((long long)w * 2) * By + i];
    if (m > -INFINITY) S += parts[((long long)w * 2 + 1) * By + i] * expf(m - M);
  }
  row_lse[i] = M + logf(S);
}

// 32x32 tiles: dL, loss/dtemp partials, per-column sum dL*L, transposed coefficient tile for the dZ contraction
template <typename T>
__global__ void clip_finish_kernel(const float* __restrict__ L, const float* __restrict__ row_lse,
                                   const float* __restrict__ col_lse, const float* __restrict__ ysq,
                                   const float* __restrict__ zsq, const float* __restrict__ temp, int off,
                                   float wscale, float* __restrict__ out2, T* __restrict__ coef, int ldc,
                                   float* __restrict__ cvec, int By, int Bz) {
  __shared__ float tile[32][33];
  __shared__ float colacc[32];
  __shared__ float red[2][8];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float s = expf(temp[0]);
  if (ty == 0) colacc[tx] = 0.f;
  __syncthreads();
  float loss = 0.f, dtemp = 0.f;
  const int j = j0 + tx;
  float csum = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r;
    float cf = 0.f;
    if (i < By && j < Bz) {
      const float l = L[(long long)i * Bz + j];
      const bool pos = (i == off + j);
      const float dl = (expf(l - row_lse[i]) + expf(l - col_lse[j]) - (pos ? 2.f : 0.f)) * wscale;
      if (pos) loss += (row_lse[i] + col_lse[j] - 2.f * l) * wscale;
      dtemp += dl * l;
      csum += dl * l;
      cf = dl * s * rsqrtf(ysq[i]) * rsqrtf(zsq[j]);
    }
    tile[ty + 8 * r][tx] = cf;
  }
  atomicAdd(&colacc[tx], csum);
  loss = warp_sum(loss);
  dtemp = warp_sum(dtemp);
  if (tx == 0) {
    red[0][ty] = loss;
    red[1][ty] = dtemp;
  }
  __syncthreads();
  // transposed write: coef[j][i]
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int jj = j0 + ty + 8 * r, ii = i0 + tx;
    if (jj < Bz && ii < By) coef[(long long)jj * ldc + ii] = from_f<T>(tile[tx][ty + 8 * r]);
  }
  if (ty == 0 && j < Bz) atomicAdd(&cvec[j], colacc[tx] / zsq[j]);
  if (ty == 0 && tx == 0) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; ++k) {
      a += red[0][k];
      b += red[1][k];
    }
    if (a != 0.f) atomicAdd(&out2[0], a);
    atomicAdd(&out2[1], b);
  }
}

// Retrieval rank of the matching brain embedding for every speech row (the reference's Classifier, models.py:403-460,
// computes the full cosine matrix in an O(B^2) Python loop and then argmax / topk): rank[r] = number of brains j whose
// cosine similarity to speech row off + r exceeds that of the matching brain r.  One warp per row; |y_i| is common to a
// row and drops out of the comparison.
__global__ void clip_rank_kernel(const float* __restrict__ raw, const float* __restrict__ zsq,
                                 int32_t* __restrict__ rank, int Bz, int off, int rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* rr = raw + (long long)(off + warp) * Bz;
  const float own = rr[warp] * rsqrtf(zsq[warp]);
  int cnt = 0;
  for (int j = lane; j < Bz; j += 32)
    if (j != warp && rr[j] * rsqrtf(zsq[j]) > own) ++cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rank[warp] = cnt;
}

}  // namespace megb200

using namespace megb200;

extern "C" {

int megb200_clip_ce_partial(const float* raw, const float* ysq, const float* zsq, const float* temp, float* L,
                            float* row_max, float* row_sumexp, float* col_lse, int By, int Bz, void* stream) {
  MEGB_CHECK_ARG(raw && ysq && zsq && temp && L && row_max && row_sumexp && col_lse && By > 0 && Bz > 0,
                 "clip_ce_partial: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  clip_rows_kernel<<<(By * 32 + 255) / 256, 256, 0, st>>>(raw, ysq, zsq, temp, L, row_max, row_sumexp, By, Bz);
  MEGB_LAUNCH_OK();
  clip_cols_kernel<<<(Bz + 31) / 32, dim3(32, kColsLanes), 0, st>>>(L, col_lse, By, Bz);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_clip_lse_merge(const float* parts, int W, int By, float* row_lse, void* stream) {
  MEGB_CHECK_ARG(parts && row_lse && W > 0 && By > 0, "clip_lse_merge: bad arguments");
  clip_lse_merge_kernel<<<(By + 255) / 256, 256, 0, (cudaStream_t)stream>>>(parts, W, By, row_lse);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_clip_ce_finish(const float* L, const float* row_lse, const float* col_lse, const float* ysq,
                           const float* zsq, const float* temp, int off, float wscale, float* out2, void* coef,
                           int ldc, float* cvec, int By, int Bz, int dtype, void* stream) {
  MEGB_CHECK_ARG(L && row_lse && col_lse && ysq && zsq && temp && out2 && coef && cvec && By > 0 && Bz > 0,
                 "clip_ce_finish: bad arguments");
  MEGB_CHECK_ARG(ldc >= By, "clip_ce_finish: ldc < By");
  MEGB_CHECK_ARG(off >= 0 && off + Bz <= By, "clip_ce_finish: positives [off, off+Bz) outside the %d speech rows", By);
  dim3 grid((Bz + 31) / 32, (By + 31) / 32), block(32, 8);
  MEGB_CHECK_ARG(grid.y <= 65535, "clip_ce_finish: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MEGB200_F32)
    clip_finish_kernel<float><<<grid, block, 0, st>>>(L, row_lse, col_lse, ysq, zsq, temp, off, wscale, out2,
                                                      (float*)coef, ldc, cvec, By, Bz);
  else if (dtype == MEGB200_BF16)
    clip_finish_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(L, row_lse, col_lse, ysq, zsq, temp, off, wscale, out2,
                                                              (__nv_bfloat16*)coef, ldc, cvec, By, Bz);
  else
    MEGB_CHECK_ARG(false, "clip_ce_finish: unknown dtype");
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_retrieval_rank(const float* raw, const float* zsq, int32_t* rank, int By, int Bz, int off, void* stream) {
  MEGB_CHECK_ARG(raw && zsq && rank && By > 0 && Bz > 0 && off >= 0 && off + Bz <= By, "retrieval_rank: bad arguments");
  clip_rank_kernel<<<(Bz * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(raw, zsq, rank, Bz, off, Bz);
  MEGB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
