// Tensor-core contraction engine (MEGB200_TC): tcgen05.mma kind::f16 (bf16 x bf16 -> f32 in TMEM), operands staged
// in shared memory by TMA through an mbarrier ring, warp-specialised CTAs (warp 0 = TMA producer, warp 1 = TMEM
// allocator + single-thread MMA issuer, warps 2-5 = epilogue, one TMEM lane quarter each).
//
//   tc_tapgemm      out[pos][n] = sum_{tap,c} A[pos + off(tap)][c] * W[g][n][tap*Kp + c]   (conv fwd / dgrad / 1x1 /
//                   grouped subject GEMM).  A: 3-D tensor map (c, t, b); the tap shift is a coordinate offset and TMA's
//                   out-of-bounds zero fill is the "same" padding.  Both operands K-major, SWIZZLE_128B.
//   tc_wgrad        dW[m][c][tap] += sum_pos dU[pos][m] * X[pos + off(tap)][c]; positions are the reduction, so both
//                   operands are MN-major straight from the PC layout; split over sample ranges.
//   tc_clip_logits  raw[i][j] += sum_d Y[i][d] Z[j][d]; split-K over D, f32 atomics.
//   tc_clip_dz      dZ[j][d] = gs * (sum_i coef[j][i] Y[i][d] - Z[j][d] cvec[j]); A K-major, B MN-major.
#include <stdlib.h>

#include "tc_common.cuh"

#include <mutex>

namespace megb200 {
namespace tc {

constexpr int kThreads = 192;
constexpr int kHeaderBytes = 1024;  // barriers + tmem slot + small per-CTA tables live in front of the stage ring

struct Header {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_slot;
  uint32_t pad;
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// ================================================================================================ tap GEMM
struct TapGemmParams {
  int T, N, Cos, Crs, Co2s, taps, dil, Kp, kchunks, tiles_per_sample, ntiles, NT, n_last, stages, epi, tma_epi, prefetch, ksub;
  uint32_t stg_bytes;   // CTA-pair kernel: dedicated staging area of the TMA-store epilogue
  float* stats;         // optional BatchNorm statistics sums[2][N] (sum, sum of squares) fused into the f32 epilogue
  long long* dbg;       // MEGB200_DEBUG_TIMING: per-CTA clock64 stamps {start, first data, accumulator ready, end}
  uint32_t tmem_cols, idesc;
  const float* bias;
  const __nv_bfloat16* Res;
  const int32_t* group;
  void* Out;
  __nv_bfloat16* Out2;
};

// Epilogue tail shared by the single-CTA and the CTA-pair kernels: one thread = one output position (TMEM lane);
// `trow` addresses this thread's lane at the first accumulator column of the tile, `n0` is the tile's first channel.
__device__ __forceinline__ void tap_epilogue(const TapGemmParams& p, uint32_t trow, const float* sbias, int n0,
                                             long long row, bool valid) {
  const bool glu = p.epi == MEGB200_EPI_GLU;
  const int half = p.NT / 2;
  const int Ch = p.N / 2;
  if (glu) {
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.Out) + row * p.Cos;
    __nv_bfloat16* vrow = p.Out2 + row * p.Co2s;
    for (int ch = 0; ch < half / 16; ++ch) {
      float va[16], vg[16];
      tmem_ld16(trow + (uint32_t)(ch * 16), va);
      tmem_ld16(trow + (uint32_t)(half + ch * 16), vg);
      if (!valid) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = n0 + ch * 16 + h * 8;
        if (c0 >= p.Cos) continue;
        float o[8], a8[8], g8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool in = (c0 + i) < Ch;
          a8[i] = in ? va[h * 8 + i] + sbias[ch * 16 + h * 8 + i] : 0.f;
          g8[i] = in ? vg[h * 8 + i] + sbias[half + ch * 16 + h * 8 + i] : 0.f;
          o[i] = in ? a8[i] * sigmoid_f(g8[i]) : 0.f;
        }
        store8(orow + c0, o);
        if (c0 + 8 <= Ch) {                  // Ch % 8 == 0 is required by the launcher for this epilogue
          store8(vrow + c0, a8);
          store8(vrow + Ch + c0, g8);
        }
      }
    }
  } else {
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.Out) + row * p.Cos;
    float* frow = reinterpret_cast<float*>(p.Out) + row * p.Cos;
    __nv_bfloat16* prow = p.Out2 ? p.Out2 + row * p.Co2s : nullptr;
    const __nv_bfloat16* rrow = p.Res ? p.Res + row * p.Crs : nullptr;
    for (int ch = 0; ch < p.NT / 16; ++ch) {
      float v[16];
      tmem_ld16(trow + (uint32_t)(ch * 16), v);
      if (!valid) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = n0 + ch * 16 + h * 8;
        if (c0 >= p.Cos) continue;
        float o[8];
        float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (rrow != nullptr && c0 < p.Crs) load8(rrow + c0, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = c0 + i;
          o[i] = (n < p.N) ? v[h * 8 + i] + sbias[ch * 16 + h * 8 + i] + r[i] : 0.f;
        }
        if (p.epi == MEGB200_EPI_STORE_F32) {
          store8(frow + c0, o);
        } else if (p.epi == MEGB200_EPI_GELU) {
          if (c0 < p.Co2s) store8(prow + c0, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = gelu_t<__nv_bfloat16>(o[i]);     // out-of-range channels are 0; gelu(0) = 0
          store8(orow + c0, o);
        } else {
          store8(orow + c0, o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-store epilogue.  The direct-store epilogue above makes every thread write its own row: 32 different 128-byte
// lines per warp instruction and a dependent residual load per chunk; measured, it halves the kernel's throughput.
// Here a thread still owns one accumulator row (TMEM lane), but the tile leaves through shared memory: 32-channel
// slices are written in the TMA swizzle pattern (conflict-free 16-byte stores), one elected thread issues
// cp.async.bulk.tensor stores (coalesced, asynchronous, rows >= T and pad channels clipped by the TMA unit), and the
// residual row is prefetched one slice ahead.  `stg` is a 1024-aligned staging area of `stg_bytes`.
struct TapStoreMaps {
  CUtensorMap out;    // Out:  (Cos, T, B), box (32, 128, 1); bf16 -> SWIZZLE_64B, f32 -> SWIZZLE_128B
  CUtensorMap out2;   // Out2: (Co2s, T, B) bf16, box (32, 128, 1), SWIZZLE_64B
};

__device__ __forceinline__ void stage_bf16_slice(uint32_t base, int rloc, const float (&o)[32]) {
  const uint32_t rowaddr = base + (uint32_t)rloc * 64u;
  const uint32_t sw = (uint32_t)(rloc >> 1) & 3u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    st_shared_v4(rowaddr + (((uint32_t)c ^ sw) << 4), pack_bf16x2(o[8 * c], o[8 * c + 1]),
                 pack_bf16x2(o[8 * c + 2], o[8 * c + 3]), pack_bf16x2(o[8 * c + 4], o[8 * c + 5]),
                 pack_bf16x2(o[8 * c + 6], o[8 * c + 7]));
}
__device__ __forceinline__ void stage_f32_slice(uint32_t base, int rloc, const float (&o)[32]) {
  const uint32_t rowaddr = base + (uint32_t)rloc * 128u;
  const uint32_t sw = (uint32_t)rloc & 7u;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st_shared_v4(rowaddr + (((uint32_t)c ^ sw) << 4), __float_as_uint(o[4 * c]), __float_as_uint(o[4 * c + 1]),
                 __float_as_uint(o[4 * c + 2]), __float_as_uint(o[4 * c + 3]));
}

// kHalves = number of epilogue warps per TMEM lane quarter (1: 4 warps / 128 threads, 2: 8 warps / 256 threads).
// With two halves the 32-channel groups of a tile are dealt alternately to the two warps that share a lane quarter
// (`h` = 0 / 1), which halves the per-tile epilogue time: measured, four warps per SM cannot keep up with the MMA
// stream (their element-wise math, not memory, is the limit).
template <bool kPersistent, int kHalves>
__device__ __forceinline__ void tap_epilogue_tma(const TapGemmParams& p, const TapStoreMaps& maps, uint32_t trow,
                                                 const float* sbias, int n0, int b, int t0, int rloc, bool row_valid,
                                                 bool tile_valid, uint32_t stg, uint32_t stg_bytes, int etid, int h,
                                                 float* sstat = nullptr, int sstat_base = 0, int sstat_stride = 0) {
  // sstat (optional, f32 epilogue only): shared-memory table [2][sstat_stride] of per-channel (sum, sum of squares)
  // accumulators, indexed by channel - sstat_base; the BatchNorm batch statistics are reduced here from the staged
  // f32 tile instead of a separate pass over the 118 MB tensor.
  constexpr int kEpiThreads = 128 * kHalves;
  const bool glu = p.epi == MEGB200_EPI_GLU;
  const bool gelu = p.epi == MEGB200_EPI_GELU;
  const bool f32 = p.epi == MEGB200_EPI_STORE_F32;
  const int half = p.NT / 2;
  const int Ch = p.N / 2;
  const uint32_t slice_bytes = f32 ? 16384u : 8192u;
  const int per_group = glu ? 3 : (gelu ? 2 : 1);
  const int groups = (glu ? half : p.NT) / 32;
  const int iters = (groups + kHalves - 1) / kHalves;
  int cap_iters = (int)(stg_bytes / slice_bytes) / (per_group * kHalves);   // iterations whose slices fit at once
  if (cap_iters < 1) cap_iters = 1;
  // an odd group count leaves the last iteration half empty: when every slice of the tile fits, it is ONE round
  if ((uint32_t)(groups * per_group) * slice_bytes <= stg_bytes) cap_iters = iters;
  const long long row = (long long)b * p.T + t0 + rloc;
  const __nv_bfloat16* rrow = (p.Res != nullptr && row_valid && tile_valid) ? p.Res + row * p.Crs + n0 : nullptr;
  uint4 rnext[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) rnext[c] = make_uint4(0u, 0u, 0u, 0u);
  if (rrow != nullptr && h < groups) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (n0 + 32 * h + 8 * c < p.Crs) rnext[c] = *reinterpret_cast<const uint4*>(rrow + 32 * h + 8 * c);
  }
  // Rounds: all channel groups that fit the staging area are converted and staged back to back; ONE proxy fence and
  // ONE barrier per round, then a single thread issues the round's TMA stores (measured: a fence + barrier per group
  // costs ~2500 cycles per group and made the epilogue as long as the main loop).
  const bool dbgp = kPersistent && p.dbg != nullptr && etid == 0;
  long long dt0 = dbgp ? clock64() : 0;
  auto dbg_phase = [&](int slot) {
    if (dbgp) {
      const long long t = clock64();
      atomicAdd(reinterpret_cast<unsigned long long*>(&p.dbg[16 * blockIdx.x + 8 + slot]), (unsigned long long)(t - dt0));
      dt0 = clock64();
    }
  };
  for (int jj0 = 0; jj0 < iters; jj0 += cap_iters) {
    if (jj0 > 0 || kPersistent) {
      // the staging area is reused: the previous stores must have finished reading it
      if (elect_one()) tma_store_wait_read();
      dbg_phase(0);
      named_bar_sync(2, kEpiThreads);
      dbg_phase(1);
    }
    const int jj1 = (jj0 + cap_iters < iters) ? jj0 + cap_iters : iters;
    for (int jj = jj0; jj < jj1; ++jj) {
      const int j = jj * kHalves + h;
      if (j >= groups) break;
      const uint32_t sbase = stg + (uint32_t)(((jj - jj0) * kHalves + h) * per_group) * slice_bytes;
      float o[32];
      tmem_ld32(trow + (uint32_t)(32 * j), o);
      const int c_first = n0 + 32 * j;                       // first logical output channel of this group
      float bv[32];
      {
        const float4* sb4 = reinterpret_cast<const float4*>(sbias + 32 * j);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v4 = sb4[c];
          bv[4 * c] = v4.x; bv[4 * c + 1] = v4.y; bv[4 * c + 2] = v4.z; bv[4 * c + 3] = v4.w;
        }
      }
      if (glu) {
        float g[32];
        tmem_ld32(trow + (uint32_t)(half + 32 * j), g);
        const float4* sg4 = reinterpret_cast<const float4*>(sbias + half + 32 * j);
        const bool full = c_first + 32 <= Ch;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v4 = sg4[c];
          g[4 * c] += v4.x; g[4 * c + 1] += v4.y; g[4 * c + 2] += v4.z; g[4 * c + 3] += v4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool in = full || (c_first + i) < Ch;
          o[i] = in ? o[i] + bv[i] : 0.f;
          g[i] = in ? g[i] : 0.f;
        }
        stage_bf16_slice(sbase, rloc, o);
        stage_bf16_slice(sbase + slice_bytes, rloc, g);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = o[i] * sigmoid_f(g[i]);      // out-of-range channels: 0 * 0.5 = 0
        stage_bf16_slice(sbase + 2 * slice_bytes, rloc, o);
      } else {
        uint4 rc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rc[c] = rnext[c];
        if (rrow != nullptr && j + kHalves < groups) {
          const int jn = j + kHalves;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            rnext[c] = (n0 + 32 * jn + 8 * c < p.Crs) ? *reinterpret_cast<const uint4*>(rrow + 32 * jn + 8 * c)
                                                       : make_uint4(0u, 0u, 0u, 0u);
        }
        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(rc);
        const bool full = c_first + 32 <= p.N;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 rf = __bfloat1622float2(rh[i]);
          o[2 * i] = (full || c_first + 2 * i < p.N) ? o[2 * i] + bv[2 * i] + rf.x : 0.f;
          o[2 * i + 1] = (full || c_first + 2 * i + 1 < p.N) ? o[2 * i + 1] + bv[2 * i + 1] + rf.y : 0.f;
        }
        if (f32) {
          stage_f32_slice(sbase, rloc, o);
        } else if (gelu) {
          stage_bf16_slice(sbase, rloc, o);                                   // pre-activation P
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = gelu_t<__nv_bfloat16>(o[i]);   // out-of-range channels are already 0 and gelu(0) = 0;
          // a per-element predicate here compiled to 32 serialised branches (2900 cycles per group)
          stage_bf16_slice(sbase + slice_bytes, rloc, o);                     // Q = gelu(P)
        } else {
          stage_bf16_slice(sbase, rloc, o);
        }
      }
    }
    if (!kPersistent && p.dbg != nullptr && etid == 0 && jj0 == 0) p.dbg[16 * blockIdx.x + 4] = clock64();
    dbg_phase(2);
    fence_proxy_async();
    named_bar_sync(2, kEpiThreads);
    dbg_phase(3);
    if (!kPersistent && p.dbg != nullptr && etid == 0 && jj0 == 0) p.dbg[16 * blockIdx.x + 5] = clock64();
    // the stores are issued BEFORE the statistics pass: both only read the staging area, so the TMA engine drains it
    // while the warps reduce it
    if (tile_valid && elect_one()) {
      // one elected lane per warp issues the stores of "its" slices (a single issuing thread spent ~380 cycles per
      // store, 1900 per tile); the bulk group is per thread, so the same lanes wait for it before the area is reused
      const int nslots = (jj1 - jj0) * kHalves;
      bool any = false;
      for (int slot = etid >> 5; slot < nslots; slot += kEpiThreads / 32) {
        const int jh = jj0 * kHalves + slot;
        if (jh >= groups) break;
        const int cf = n0 + 32 * jh;
        const uint32_t sb2 = stg + (uint32_t)(slot * per_group) * slice_bytes;
        if (glu) {
          tma_store_3d(&maps.out2, sb2, cf, t0, b);
          tma_store_3d(&maps.out2, sb2 + slice_bytes, Ch + cf, t0, b);
          tma_store_3d(&maps.out, sb2 + 2 * slice_bytes, cf, t0, b);
        } else if (gelu) {
          tma_store_3d(&maps.out2, sb2, cf, t0, b);
          tma_store_3d(&maps.out, sb2 + slice_bytes, cf, t0, b);
        } else {
          tma_store_3d(&maps.out, sb2, cf, t0, b);
        }
        any = true;
      }
      if (any) tma_store_commit();
    }
    dbg_phase(4);
    if (f32 && sstat != nullptr && tile_valid) {
      // column sums of the staged slices: ONE warp owns a slice (so the table update needs no atomics).  Each LDS.128
      // covers 4 rows x 8 chunks (conflict-free under the 128 B swizzle); a lane keeps 4 channels of every 4th row and
      // the four row-lanes are folded with two shuffle steps.  (Earlier forms: every warp read 16 rows of every slice
      // and met in 8-way shared atomics -- 6000 cycles per tile; two warps per slice with 2-way atomics -- 3500.)
      constexpr int kWarps = kEpiThreads / 32;
      const int l = etid & 31, chunk = l & 7, rsub = l >> 3;
      const int nslots = (jj1 - jj0) * kHalves;
      int rows_valid = p.T - t0;
      if (rows_valid > 128) rows_valid = 128;
      for (int slot = etid >> 5; slot < nslots; slot += kWarps) {
        const int jh = jj0 * kHalves + slot;
        if (jh >= groups) break;
        const uint32_t sb2 = stg + (uint32_t)(slot * per_group) * slice_bytes;
        float su[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int r = rsub; r < rows_valid; r += 4) {
          float4 v;
          asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
              : "r"(sb2 + (uint32_t)r * 128u + ((((uint32_t)chunk) ^ ((uint32_t)r & 7u)) << 4)));
          su[0] += v.x; su[1] += v.y; su[2] += v.z; su[3] += v.w;
          sq[0] = fmaf(v.x, v.x, sq[0]); sq[1] = fmaf(v.y, v.y, sq[1]);
          sq[2] = fmaf(v.z, v.z, sq[2]); sq[3] = fmaf(v.w, v.w, sq[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          su[k] += __shfl_xor_sync(0xffffffffu, su[k], 8);
          sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], 8);
          su[k] += __shfl_xor_sync(0xffffffffu, su[k], 16);
          sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], 16);
        }
        if (rsub == 0) {
          const int ch0 = n0 + 32 * jh + 4 * chunk - sstat_base;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ch0 + k < sstat_stride) {
              sstat[ch0 + k] += su[k];
              sstat[sstat_stride + ch0 + k] += sq[k];
            }
        }
      }
    }
    dbg_phase(5);
    if (!kPersistent && p.dbg != nullptr && etid == 0 && jj0 == 0) p.dbg[16 * blockIdx.x + 6] = clock64();
  }
  if (!kPersistent && elect_one()) tma_store_wait_read();   // shared memory must stay alive until the stores have read it
}

__global__ void __launch_bounds__(kThreads)
tc_tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ TapStoreMaps maps, const TapGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header* hdr = reinterpret_cast<Header*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + 512);  // NT <= 128 floats fit in the header; larger go below
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t a_bytes = kTileM * 128, b_bytes = (uint32_t)p.NT * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* ring = smem + kHeaderBytes;
  float* sbias_big = reinterpret_cast<float*>(ring + (size_t)p.stages * stage_bytes);
  if (p.NT > 128) sbias = sbias_big;
  float* sstat = sbias_big + (p.NT > 128 ? p.NT : 0);      // [2][NT] BatchNorm partial sums of this tile

  const bool glu = p.epi == MEGB200_EPI_GLU;
  const int half = p.NT / 2;          // GLU: the tile is [half "a" channels | the matching half gate channels]
  const int Ch = p.N / 2;
  const int n_tile = blockIdx.x % p.ntiles;
  const int m_tile = blockIdx.x / p.ntiles;
  const int b = m_tile / p.tiles_per_sample;
  const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
  const int n0 = n_tile * (glu ? half : p.NT);   // first output channel of this tile
  const int g = p.group ? p.group[b] : 0;
  const int num_k = p.taps * p.kchunks;
  // ragged N tiling: every tile is NT wide except the last one (n_last): 320 = 256 + 64 keeps the MMA N at the widths
  // the tensor core runs at full rate (N = 160 measured at ~60 %)
  const int wN = (n_tile == p.ntiles - 1) ? p.n_last : p.NT;
  const uint32_t idesc = make_idesc(wN, 0, 0);
  if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[16 * blockIdx.x + 0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < num_k; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
        const uint32_t full = smem_u32(&hdr->full[s]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const int tap = it / p.kchunks, kc = it % p.kchunks;
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        tma_load_3d(sa, &tmA, kc * kBlockK, t0 + (tap - p.taps / 2) * p.dil, b, full);
        const int kw = tap * p.Kp + kc * kBlockK;
        if (glu) {
          // two half-height boxes of the naturally packed weights: rows [n0, n0+half) and [Ch+n0, Ch+n0+half)
          tma_load_3d(sa + a_bytes, &tmW, kw, n0, g, full);
          tma_load_3d(sa + a_bytes + (uint32_t)half * 128, &tmW, kw, Ch + n0, g, full);
        } else {
          tma_load_3d(sa + a_bytes, &tmW, kw, n0, g, full);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int it = 0; it < num_k; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
        if (p.dbg != nullptr && it == 0) p.dbg[16 * blockIdx.x + 1] = clock64();
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc(sa, 16, 1024);
        const uint64_t db = make_smem_desc(sa + a_bytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (it | k) != 0);
        umma_commit(smem_u32(&hdr->empty[s]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> fused element-wise tail -> global (thread = one position)
    const int et = threadIdx.x - 64;
    if (p.stats != nullptr)
      for (int i = et; i < 2 * p.NT; i += 128) sstat[i] = 0.f;
    for (int i = et; i < wN; i += 128) {
      int n;
      bool ok;
      if (glu) {
        const int c = n0 + (i < half ? i : i - half);
        ok = c < Ch;
        n = (i < half) ? c : Ch + c;
      } else {
        n = n0 + i;
        ok = n < p.N;
      }
      sbias[i] = (p.bias != nullptr && ok) ? p.bias[(long long)g * p.N + n] : 0.f;
    }
    named_bar_sync(1, 128);
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    if (p.dbg != nullptr && et == 0) p.dbg[16 * blockIdx.x + 2] = clock64();
    tc_fence_after();
    const int q = warp & 3;
    const int t = t0 + q * 32 + lane;
    const long long row = (long long)b * p.T + t;
    TapGemmParams pe = p;
    pe.NT = wN;
    if (p.tma_epi) {
      // every MMA has completed, so the whole stage ring is free: it becomes the staging area of the stores
      tap_epilogue_tma<false, 1>(pe, maps, tmem_base + ((uint32_t)(q * 32) << 16), sbias, n0, b, t0, q * 32 + lane,
                                 t < p.T, true, smem_u32(ring), (uint32_t)p.stages * stage_bytes, et, 0,
                                 p.stats != nullptr ? sstat : nullptr, n0, p.NT);
      if (p.stats != nullptr) {
        named_bar_sync(2, 128);
        for (int i = et; i < wN; i += 128) {
          const int n = n0 + i;
          if (n < p.N) {
            atomicAdd(&p.stats[n], sstat[i]);
            atomicAdd(&p.stats[p.N + n], sstat[p.NT + i]);
          }
        }
      }
    } else {
      tap_epilogue(pe, tmem_base + ((uint32_t)(q * 32) << 16), sbias, n0, row, t < p.T);
    }
    if (p.dbg != nullptr && et == 0) p.dbg[16 * blockIdx.x + 3] = clock64();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ tap GEMM, halo form
// k > 1 convolutions read every activation row `taps` times.  Here the activation tile is staged ONCE per 64-channel
// chunk with its halo (128 + (taps-1)*dil rows) and each tap's MMA reads it through a shared-memory descriptor whose
// start address is shifted by tap*dil rows (128 bytes each): the swizzle XOR acts on absolute shared-memory address
// bits, so a row-shifted window of a SWIZZLE_128B tile is still a valid K-major operand.  Activation traffic from L2
// drops by `taps`x (the kernel is L2->SM bandwidth bound: 909 MB per 320->320 conv launch at config 2).
struct HeaderH {
  uint64_t fullA[4];
  uint64_t emptyA[4];
  uint64_t fullB[8];
  uint64_t emptyB[8];
  uint64_t tmem_full;
  uint32_t tmem_slot;
  uint32_t pad;
};

__global__ void __launch_bounds__(kThreads)
tc_tapgemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ TapStoreMaps maps, const TapGemmParams p, const int a_rows,
                       const int stagesA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  HeaderH* hdr = reinterpret_cast<HeaderH*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + 512);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t a_slot = ((uint32_t)a_rows * 128u + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.NT * 128;
  uint8_t* ringA = smem + kHeaderBytes;
  uint8_t* ringB = ringA + (size_t)stagesA * a_slot;
  const uint32_t ring_bytes = (uint32_t)stagesA * a_slot + (uint32_t)p.stages * b_bytes;
  float* sbias_big = reinterpret_cast<float*>(ringA + ring_bytes);
  if (p.NT > 128) sbias = sbias_big;

  const bool glu = p.epi == MEGB200_EPI_GLU;
  const int half = p.NT / 2;
  const int Ch = p.N / 2;
  const int n_tile = blockIdx.x % p.ntiles;
  const int m_tile = blockIdx.x / p.ntiles;
  const int b = m_tile / p.tiles_per_sample;
  const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
  const int n0 = n_tile * (glu ? half : p.NT);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < stagesA; ++s) {
      mbar_init(smem_u32(&hdr->fullA[s]), 1);
      mbar_init(smem_u32(&hdr->emptyA[s]), 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->fullB[s]), 1);
      mbar_init(smem_u32(&hdr->emptyB[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int itb = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const int sa = kc % stagesA;
        mbar_wait(smem_u32(&hdr->emptyA[sa]), ((kc / stagesA) & 1) ^ 1);
        const uint32_t fa = smem_u32(&hdr->fullA[sa]);
        mbar_arrive_expect_tx(fa, (uint32_t)a_rows * 128u);
        tma_load_3d(smem_u32(ringA + (size_t)sa * a_slot), &tmA, kc * kBlockK, t0 - (p.taps / 2) * p.dil, b, fa);
        for (int tap = 0; tap < p.taps; ++tap, ++itb) {
          const int sb = itb % p.stages;
          mbar_wait(smem_u32(&hdr->emptyB[sb]), ((itb / p.stages) & 1) ^ 1);
          const uint32_t fb = smem_u32(&hdr->fullB[sb]);
          mbar_arrive_expect_tx(fb, b_bytes);
          const uint32_t dstb = smem_u32(ringB + (size_t)sb * b_bytes);
          const int kw = tap * p.Kp + kc * kBlockK;
          if (glu) {
            tma_load_3d(dstb, &tmW, kw, n0, 0, fb);
            tma_load_3d(dstb + (uint32_t)half * 128, &tmW, kw, Ch + n0, 0, fb);
          } else {
            tma_load_3d(dstb, &tmW, kw, n0, 0, fb);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int itb = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const int sa = kc % stagesA;
        mbar_wait(smem_u32(&hdr->fullA[sa]), (kc / stagesA) & 1);
        const uint32_t abase = smem_u32(ringA + (size_t)sa * a_slot);
        for (int tap = 0; tap < p.taps; ++tap, ++itb) {
          const int sb = itb % p.stages;
          mbar_wait(smem_u32(&hdr->fullB[sb]), (itb / p.stages) & 1);
          tc_fence_after();
          const uint32_t aaddr = abase + (uint32_t)(tap * p.dil) * 128u;      // window shifted by tap*dil rows
          uint64_t da = make_smem_desc(aaddr, 16, 1024);
          const uint64_t db = make_smem_desc(smem_u32(ringB + (size_t)sb * b_bytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * 32), p.idesc, (kc | tap | k) != 0);
          umma_commit(smem_u32(&hdr->emptyB[sb]));
        }
        umma_commit(smem_u32(&hdr->emptyA[sa]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    const int et = threadIdx.x - 64;
    for (int i = et; i < p.NT; i += 128) {
      int n;
      bool ok;
      if (glu) {
        const int c = n0 + (i < half ? i : i - half);
        ok = c < Ch;
        n = (i < half) ? c : Ch + c;
      } else {
        n = n0 + i;
        ok = n < p.N;
      }
      sbias[i] = (p.bias != nullptr && ok) ? p.bias[n] : 0.f;
    }
    named_bar_sync(1, 128);
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    const int q = warp & 3;
    const int t = t0 + q * 32 + lane;
    const long long row = (long long)b * p.T + t;
    if (p.tma_epi) {
      tap_epilogue_tma<false, 1>(p, maps, tmem_base + ((uint32_t)(q * 32) << 16), sbias, n0, b, t0, q * 32 + lane,
                                 t < p.T, true, smem_u32(ringA), ring_bytes, et, 0);
    } else {
      tap_epilogue(p, tmem_base + ((uint32_t)(q * 32) << 16), sbias, n0, row, t < p.T);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ tap GEMM, CTA pair
// Persistent cta_group::2 form of the tap GEMM: a cluster of two CTAs (one TPC) computes a 256-position x NT-channel
// tile with ONE tcgen05.mma.cta_group::2 stream issued by the leader; each CTA stages its own 128 activation rows and
// HALF of the weight tile, which halves the per-SM shared-memory operand traffic (the single-CTA kernel is bound by
// it).  Accumulators are double-buffered in TMEM so that the epilogue of tile i overlaps the main loop of tile i+1.
struct Header2 {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_slot;
  uint32_t pad;
};

constexpr int kThreadsPair = 64 + 256;   // producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)

__global__ void __launch_bounds__(kThreadsPair, 1)
tc_tapgemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ TapStoreMaps maps, const TapGemmParams p, const int total_m_tiles,
                   const int num_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header2* hdr = reinterpret_cast<Header2*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int half = p.NT / 2;                                  // weight rows staged by each CTA
  const int kSub = p.ksub;                                    // 64-wide k chunks per pipeline stage (2, or 1 for a deeper ring)
  const uint32_t a_bytes = kTileM * 128, b_bytes = (uint32_t)half * 128;
  const uint32_t sub_bytes = a_bytes + b_bytes;
  const uint32_t stage_bytes = kSub * sub_bytes;
  uint8_t* ring = smem + kHeaderBytes;
  uint8_t* stg = ring + (size_t)p.stages * stage_bytes;                              // TMA-store staging (1024-aligned)
  float* sbias = reinterpret_cast<float*>(stg + p.stg_bytes);                        // [2][NT]: one table per acc stage
  float* sstat = sbias + 2 * p.NT;                 // [2][Cos] BatchNorm partial sums of all tiles of this CTA
  const bool glu = p.epi == MEGB200_EPI_GLU;
  const int Ch = p.N / 2;
  const int num_k = p.taps * p.kchunks;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);      // the leader's producer arrives once and expects both CTAs' bytes
      mbar_init(smem_u32(&hdr->empty[s]), 1);     // one multicast tcgen05.commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&hdr->tmem_full[a]), 1);
      mbar_init(smem_u32(&hdr->tmem_empty[a]), 16);  // 8 epilogue warps x 2 CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();                              // barriers of both CTAs initialised before any remote traffic
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      long long dbg_prod_wait = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int n_tile = item % p.ntiles;
        const int m_tile = (item / p.ntiles) * 2 + (int)rank;
        // a missing second tile of the last pair reads sample index B: fully out of bounds, zero filled
        const int b = m_tile < total_m_tiles ? m_tile / p.tiles_per_sample : 0x3fffffff;
        const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
        // weight rows of this CTA: plain tiles split [n0, n0+half) | [n0+half, n0+NT); GLU tiles are
        // [half a-channels | their half gate channels], i.e. the leader stages the a rows and the peer the gates
        const int wN = (n_tile == p.ntiles - 1) ? p.n_last : p.NT;
        const int wrow = glu ? (n_tile * half + (rank ? Ch : 0)) : (n_tile * p.NT + (int)rank * (wN / 2));
        if (p.prefetch) {
          const int nitem = item + num_clusters;
          const int nm = (nitem / p.ntiles) * 2 + (int)rank;
          if (nitem < num_items && nm < total_m_tiles && nm != m_tile) {
            const int nb = nm / p.tiles_per_sample, nt0 = (nm % p.tiles_per_sample) * kTileM;
            for (int tap = 0; tap < p.taps; ++tap)
              for (int kc = 0; kc < p.kchunks; ++kc)
                tma_prefetch_3d(&tmA, kc * kBlockK, nt0 + (tap - p.taps / 2) * p.dil, nb);
          }
        }
        // a stage holds kSub consecutive 64-wide k chunks: one barrier round trip and one commit per 8 MMAs (the single
        // MMA-issuing thread of the pair is issue-latency bound, measured ~630 cycles per 4-MMA stage)
        for (int kit = 0; kit < num_k; kit += kSub, ++it) {
          const int s = it % p.stages;
          const long long w0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
          if (p.dbg) dbg_prod_wait += clock64() - w0;
          const uint32_t full = smem_u32(&hdr->full[s]);
          const int nsub = (num_k - kit) < kSub ? (num_k - kit) : kSub;
          if (leader) mbar_arrive_expect_tx(full, 2 * (uint32_t)nsub * sub_bytes);
          for (int u = 0; u < nsub; ++u) {
            const int tap = (kit + u) / p.kchunks, kc = (kit + u) % p.kchunks;
            const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes) + (uint32_t)u * sub_bytes;
            tma_load_3d_2sm(sa, &tmA, kc * kBlockK, t0 + (tap - p.taps / 2) * p.dil, b, full);
            tma_load_3d_2sm(sa + a_bytes, &tmW, tap * p.Kp + kc * kBlockK, wrow, 0, full);
          }
        }
      }
      if (p.dbg) p.dbg[16 * blockIdx.x + 0] = dbg_prod_wait;
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      int it = 0, tile = 0;
      long long dbg_full_wait = 0, dbg_acc_wait = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++tile) {
        const int acc = tile & 1;
        const long long w1 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&hdr->tmem_empty[acc]), ((tile >> 1) & 1) ^ 1);   // epilogues of both CTAs drained it
        if (p.dbg) dbg_acc_wait += clock64() - w1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.NT);
        const int wN = ((item % p.ntiles) == p.ntiles - 1) ? p.n_last : p.NT;
        const uint32_t idesc = make_idesc_m(256, wN, 0, 0);
        for (int kit = 0; kit < num_k; kit += kSub, ++it) {
          const int s = it % p.stages;
          const long long w2 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
          if (p.dbg) dbg_full_wait += clock64() - w2;
          tc_fence_after();
          const int nsub = (num_k - kit) < kSub ? (num_k - kit) : kSub;
          const uint32_t sa0 = smem_u32(ring + (size_t)s * stage_bytes);
          for (int u = 0; u < nsub; ++u) {
            const uint32_t sa = sa0 + (uint32_t)u * sub_bytes;
            const uint64_t da = make_smem_desc(sa, 16, 1024);
            const uint64_t db = make_smem_desc(sa + a_bytes, 16, 1024);
#pragma unroll
            for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
              umma_bf16_2sm(d_tmem, desc_advance(da, kk * 32), desc_advance(db, kk * 32), idesc, (kit | u | kk) != 0);
          }
          umma_commit_2sm(smem_u32(&hdr->empty[s]), 3);                      // frees the slot in BOTH CTAs
        }
        umma_commit_2sm(smem_u32(&hdr->tmem_full[acc]), 3);                  // accumulator ready in BOTH CTAs
      }
      if (p.dbg) {
        p.dbg[16 * blockIdx.x + 1] = clock64() - dbg_t0;
        p.dbg[16 * blockIdx.x + 2] = dbg_full_wait;
        p.dbg[16 * blockIdx.x + 3] = dbg_acc_wait;
        p.dbg[16 * blockIdx.x + 6] = tile;
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): each CTA drains its own 128 rows of the 256-row accumulator with eight
    // warps: warp w covers TMEM lane quarter (w & 3); the two warps of a quarter alternate over the channel groups
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    if (p.stats != nullptr)
      for (int i = et; i < 2 * p.Cos; i += 256) sstat[i] = 0.f;      // visible after the first named barrier below
    int tile = 0;
    long long dbg_epi_wait = 0, dbg_epi_work = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++tile) {
      const int acc = tile & 1;
      const int n_tile = item % p.ntiles;
      const int m_tile = (item / p.ntiles) * 2 + (int)rank;
      const int b = m_tile / p.tiles_per_sample;
      const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
      const int n0 = n_tile * (glu ? half : p.NT);
      const int wN = (n_tile == p.ntiles - 1) ? p.n_last : p.NT;
      float* sb = sbias + acc * p.NT;
      for (int i = et; i < wN; i += 256) {
        int n;
        bool ok;
        if (glu) {
          const int c = n0 + (i < half ? i : i - half);
          ok = c < Ch;
          n = (i < half) ? c : Ch + c;
        } else {
          n = n0 + i;
          ok = n < p.N;
        }
        sb[i] = (p.bias != nullptr && ok) ? p.bias[n] : 0.f;
      }
      named_bar_sync(1, 256);
      const long long w3 = p.dbg ? clock64() : 0;
      mbar_wait_sleep(smem_u32(&hdr->tmem_full[acc]), (tile >> 1) & 1, 64);
      const long long w4 = p.dbg ? clock64() : 0;
      dbg_epi_wait += w4 - w3;
      tc_fence_after();
      const int t = t0 + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      TapGemmParams pe = p;
      pe.NT = wN;
      tap_epilogue_tma<true, 2>(pe, maps, trow, sb, n0, b, t0, q * 32 + lane, t < p.T, m_tile < total_m_tiles,
                                smem_u32(stg), p.stg_bytes, et, h, p.stats != nullptr ? sstat : nullptr, 0, p.Cos);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(smem_u32(&hdr->tmem_empty[acc]), 0);
      if (p.dbg) dbg_epi_work += clock64() - w4;
    }
    if (p.stats != nullptr) {
      named_bar_sync(2, 256);
      for (int i = et; i < p.N; i += 256) {
        atomicAdd(&p.stats[i], sstat[i]);
        atomicAdd(&p.stats[p.N + i], sstat[p.Cos + i]);
      }
    }
    if (p.dbg && et == 0) {
      p.dbg[16 * blockIdx.x + 4] = dbg_epi_wait;
      p.dbg[16 * blockIdx.x + 5] = dbg_epi_work;
    }
  }
  if (p.tma_epi && warp >= 2 && elect_one()) tma_store_wait_read();
  tc_fence_before();
  cluster_sync_all();                              // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, p.tmem_cols);
  }
}

// CTA-pair + halo (the default for k > 1): the activation rows (with halo) of a 64-channel chunk are staged once and
// shared by all taps (see tc_tapgemm_halo_kernel); the same stage carries the weight half-tiles of all taps, so one
// barrier round trip feeds taps*4 MMAs.  L2->SM traffic per 256x160 tile of a 320->320 k=3 conv: 780 KB (pair) -> 460 KB.
struct Header2H {
  uint64_t fullA[4];
  uint64_t emptyA[4];
  uint64_t fullB[8];
  uint64_t emptyB[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_slot;
  uint32_t pad;
};

__global__ void __launch_bounds__(kThreadsPair, 1)
tc_tapgemm2h_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ TapStoreMaps maps, const TapGemmParams p, const int total_m_tiles,
                    const int num_items, const int a_rows, const int stagesA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header2H* hdr = reinterpret_cast<Header2H*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int half = p.NT / 2;                                  // weight rows staged by each CTA
  const uint32_t a_slot = ((uint32_t)a_rows * 128u + 1023u) & ~1023u;              // activation rows incl. halo
  const uint32_t b_bytes = (uint32_t)half * 128;
  // one stage = the activation rows (with halo) of a 64-channel chunk + the weight half-tiles of ALL taps for that
  // chunk: one barrier round trip and one commit per taps*4 MMAs
  const uint32_t stage_bytes = a_slot + (uint32_t)p.taps * b_bytes;
  uint8_t* ringA = smem + kHeaderBytes;
  uint8_t* stg = ringA + (size_t)stagesA * stage_bytes;                              // TMA-store staging (1024-aligned)
  float* sbias = reinterpret_cast<float*>(stg + p.stg_bytes);                        // [2][NT]: one table per acc stage
  float* sstat = sbias + 2 * p.NT;                 // [2][Cos] BatchNorm partial sums of all tiles of this CTA
  const bool glu = p.epi == MEGB200_EPI_GLU;
  const int Ch = p.N / 2;
  const int num_k = p.taps * p.kchunks;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < stagesA; ++s) {
      mbar_init(smem_u32(&hdr->fullA[s]), 1);     // the leader's producer arrives once and expects both CTAs' bytes
      mbar_init(smem_u32(&hdr->emptyA[s]), 1);    // one multicast tcgen05.commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&hdr->tmem_full[a]), 1);
      mbar_init(smem_u32(&hdr->tmem_empty[a]), 16);  // 8 epilogue warps x 2 CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();                              // barriers of both CTAs initialised before any remote traffic
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int ita = 0;
      long long dbg_prod_wait = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int n_tile = item % p.ntiles;
        const int m_tile = (item / p.ntiles) * 2 + (int)rank;
        const int b = m_tile < total_m_tiles ? m_tile / p.tiles_per_sample : 0x3fffffff;
        const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
        const int wrow = glu ? (n_tile * half + (rank ? Ch : 0)) : (n_tile * p.NT + (int)rank * half);
        if (p.prefetch) {
          // activation boxes of this cluster's next tile -> L2 (its weights are shared by all tiles and stay there)
          const int nitem = item + num_clusters;
          const int nm = (nitem / p.ntiles) * 2 + (int)rank;
          if (nitem < num_items && nm < total_m_tiles && nm != m_tile) {
            const int nb = nm / p.tiles_per_sample, nt0 = (nm % p.tiles_per_sample) * kTileM;
            for (int kc = 0; kc < p.kchunks; ++kc) tma_prefetch_3d(&tmA, kc * kBlockK, nt0 - (p.taps / 2) * p.dil, nb);
          }
        }
        for (int kc = 0; kc < p.kchunks; ++kc, ++ita) {
          const int sa = ita % stagesA;
          const long long w0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&hdr->emptyA[sa]), ((ita / stagesA) & 1) ^ 1);
          if (p.dbg) dbg_prod_wait += clock64() - w0;
          const uint32_t fa = smem_u32(&hdr->fullA[sa]);
          if (leader) mbar_arrive_expect_tx(fa, 2u * ((uint32_t)a_rows * 128u + (uint32_t)p.taps * b_bytes));
          const uint32_t sbase = smem_u32(ringA + (size_t)sa * stage_bytes);
          tma_load_3d_2sm(sbase, &tmA, kc * kBlockK, t0 - (p.taps / 2) * p.dil, b, fa);
          for (int tap = 0; tap < p.taps; ++tap)
            tma_load_3d_2sm(sbase + a_slot + (uint32_t)tap * b_bytes, &tmW, tap * p.Kp + kc * kBlockK, wrow, 0, fa);
        }
      }
      if (p.dbg) p.dbg[16 * blockIdx.x + 0] = dbg_prod_wait;
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      int ita = 0, tile = 0;
      long long dbg_full_wait = 0, dbg_acc_wait = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++tile) {
        const int acc = tile & 1;
        const long long w1 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&hdr->tmem_empty[acc]), ((tile >> 1) & 1) ^ 1);   // epilogues of both CTAs drained it
        if (p.dbg) dbg_acc_wait += clock64() - w1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.NT);
        for (int kc = 0; kc < p.kchunks; ++kc, ++ita) {
          const int sa = ita % stagesA;
          const long long w2 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&hdr->fullA[sa]), (ita / stagesA) & 1);
          if (p.dbg) dbg_full_wait += clock64() - w2;
          tc_fence_after();
          const uint32_t abase = smem_u32(ringA + (size_t)sa * stage_bytes);
          for (int tap = 0; tap < p.taps; ++tap) {
            // tap window = the staged rows shifted by tap*dil (absolute-address swizzle keeps it a valid operand)
            const uint64_t da = make_smem_desc(abase + (uint32_t)(tap * p.dil) * 128u, 16, 1024);
            const uint64_t db = make_smem_desc(abase + a_slot + (uint32_t)tap * b_bytes, 16, 1024);
#pragma unroll
            for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
              umma_bf16_2sm(d_tmem, desc_advance(da, kk * 32), desc_advance(db, kk * 32), p.idesc, (kc | tap | kk) != 0);
          }
          umma_commit_2sm(smem_u32(&hdr->emptyA[sa]), 3);
        }
        umma_commit_2sm(smem_u32(&hdr->tmem_full[acc]), 3);                  // accumulator ready in BOTH CTAs
      }
      if (p.dbg) {
        p.dbg[16 * blockIdx.x + 1] = clock64() - dbg_t0;
        p.dbg[16 * blockIdx.x + 2] = dbg_full_wait;
        p.dbg[16 * blockIdx.x + 3] = dbg_acc_wait;
        p.dbg[16 * blockIdx.x + 6] = tile;
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): each CTA drains its own 128 rows of the 256-row accumulator with eight
    // warps: warp w covers TMEM lane quarter (w & 3); the two warps of a quarter alternate over the channel groups
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    int tile = 0;
    long long dbg_epi_wait = 0, dbg_epi_work = 0;
    if (p.stats != nullptr)
      for (int i = et; i < 2 * p.Cos; i += 256) sstat[i] = 0.f;      // visible after the first named barrier below
    for (int item = cluster_id; item < num_items; item += num_clusters, ++tile) {
      const int acc = tile & 1;
      const int n_tile = item % p.ntiles;
      const int m_tile = (item / p.ntiles) * 2 + (int)rank;
      const int b = m_tile / p.tiles_per_sample;
      const int t0 = (m_tile % p.tiles_per_sample) * kTileM;
      const int n0 = n_tile * (glu ? half : p.NT);
      float* sb = sbias + acc * p.NT;
      for (int i = et; i < p.NT; i += 256) {
        int n;
        bool ok;
        if (glu) {
          const int c = n0 + (i < half ? i : i - half);
          ok = c < Ch;
          n = (i < half) ? c : Ch + c;
        } else {
          n = n0 + i;
          ok = n < p.N;
        }
        sb[i] = (p.bias != nullptr && ok) ? p.bias[n] : 0.f;
      }
      named_bar_sync(1, 256);
      const long long w3 = p.dbg ? clock64() : 0;
      mbar_wait_sleep(smem_u32(&hdr->tmem_full[acc]), (tile >> 1) & 1, 64);
      const long long w4 = p.dbg ? clock64() : 0;
      dbg_epi_wait += w4 - w3;
      tc_fence_after();
      const int t = t0 + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      tap_epilogue_tma<true, 2>(p, maps, trow, sb, n0, b, t0, q * 32 + lane, t < p.T, m_tile < total_m_tiles,
                                smem_u32(stg), p.stg_bytes, et, h, p.stats != nullptr ? sstat : nullptr, 0, p.Cos);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(smem_u32(&hdr->tmem_empty[acc]), 0);
      if (p.dbg) dbg_epi_work += clock64() - w4;
    }
    if (p.dbg && et == 0) {
      p.dbg[16 * blockIdx.x + 4] = dbg_epi_wait;
      p.dbg[16 * blockIdx.x + 5] = dbg_epi_work;
    }
    if (p.stats != nullptr) {
      named_bar_sync(2, 256);
      for (int i = et; i < p.N; i += 256) {
        atomicAdd(&p.stats[i], sstat[i]);
        atomicAdd(&p.stats[p.N + i], sstat[p.Cos + i]);
      }
    }
  }
  if (p.tma_epi && warp >= 2 && elect_one()) tma_store_wait_read();
  tc_fence_before();
  cluster_sync_all();                              // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ weight gradient
struct WgradParams {
  int T, M, C, taps, dil, B, kchunks /*per sample*/, ctile /*channels per CTA, multiple of 64*/, c_tiles, m_tiles;
  int samples_per_cta, stages, mode /*0 = partial slot store, 1 = atomic (grouped)*/;
  uint32_t tmem_cols, idesc0, idesc1;
  int n0_cols, n1_cols;  // the c tile is issued as one or two MMAs of n0 (+ n1) columns (multiples of 64)
  const int32_t* group;
  const int32_t* order;
  float* out;  // mode 0: partial[slot][tap][M][C]; mode 1: dW[g][M][C][taps]
};

__global__ void __launch_bounds__(kThreads)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmX,
                const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t atom_bytes = kBlockK * 128;                // 64 k-rows x 128 B
  const uint32_t a_bytes = 2 * atom_bytes;                  // 128 m channels = 2 atoms
  const uint32_t b_bytes = (uint32_t)(p.ctile / 64) * atom_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* ring = smem + kHeaderBytes;

  // blockIdx.x = (m tile, c tile), blockIdx.y = tap, blockIdx.z = sample range
  const int m0 = (blockIdx.x / p.c_tiles) * kTileM;
  const int c0 = (blockIdx.x % p.c_tiles) * p.ctile;
  const int tap = blockIdx.y;
  const int s0 = blockIdx.z * p.samples_per_cta;
  const int s1 = min(s0 + p.samples_per_cta, p.B);
  const int off = (tap - p.taps / 2) * p.dil;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    mbar_init(smem_u32(&hdr->tmem_empty), 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int si = s0; si < s1; ++si) {
        const int b = p.order ? p.order[si] : si;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
          const uint32_t full = smem_u32(&hdr->full[s]);
          mbar_arrive_expect_tx(full, stage_bytes);
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
          tma_load_3d(sa, &tmU, m0, kc * kBlockK, b, full);
          tma_load_3d(sa + atom_bytes, &tmU, m0 + 64, kc * kBlockK, b, full);
          for (int a = 0; a < p.ctile / 64; ++a)
            tma_load_3d(sa + a_bytes + a * atom_bytes, &tmX, c0 + a * 64, kc * kBlockK + off, b, full);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int it = 0, flushes = 0;
      int cur = (p.group && s0 < s1) ? p.group[p.order ? p.order[s0] : s0] : 0;
      bool fresh = true;
      for (int si = s0; si < s1; ++si) {
        const int b = p.order ? p.order[si] : si;
        const int g = p.group ? p.group[b] : 0;
        if (g != cur) {
          // hand the finished group to the epilogue and wait until it has drained the accumulator
          umma_commit(smem_u32(&hdr->tmem_full));
          mbar_wait(smem_u32(&hdr->tmem_empty), flushes & 1);
          tc_fence_after();
          ++flushes;
          cur = g;
          fresh = true;
        }
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
          // MN-major: LBO = distance between 64-channel atoms, SBO = distance between 8-position groups
          const uint64_t da = make_smem_desc(sa, atom_bytes, 1024);
          const uint64_t db0 = make_smem_desc(sa + a_bytes, atom_bytes, 1024);
          const uint64_t db1 = make_smem_desc(sa + a_bytes + (uint32_t)(p.n0_cols / 64) * atom_bytes, atom_bytes, 1024);
          auto issue = [&](int k) {
            const uint32_t acc = (!fresh || k != 0) ? 1u : 0u;
            umma_bf16(tmem_base, desc_advance(da, k * kUmmaK * 128), desc_advance(db0, k * kUmmaK * 128), p.idesc0,
                      acc);
            if (p.n1_cols > 0)
              umma_bf16(tmem_base + (uint32_t)p.n0_cols, desc_advance(da, k * kUmmaK * 128),
                        desc_advance(db1, k * kUmmaK * 128), p.idesc1, acc);
          };
          // (trimming the zero-filled k steps of a sample's last chunk -- T = 360: 3 of 4 -- was measured: the extra
          // control flow in this single issuing thread cost more than the skipped MMAs, 595 -> 578 TFLOP/s)
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) issue(k);
          fresh = false;
          umma_commit(smem_u32(&hdr->empty[s]));
        }
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    // ---------------- epilogue: thread = one output channel m (TMEM lane); columns = input channels c
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    int flushes = 0;
    int si = s0;
    while (si < s1) {
      // the run of samples [si, se) shares one group
      const int g = p.group ? p.group[p.order ? p.order[si] : si] : 0;
      int se = si + 1;
      if (p.group) {
        while (se < s1 && p.group[p.order ? p.order[se] : se] == g) ++se;
      } else {
        se = s1;
      }
      mbar_wait_sleep(smem_u32(&hdr->tmem_full), flushes & 1);
      tc_fence_after();
      const int ncols = p.n0_cols + p.n1_cols;
      for (int ch = 0; ch < ncols / 16; ++ch) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 16), v);
        if (m >= p.M) continue;
        const int cb = c0 + ch * 16;
        if (p.mode == 0) {
          float* dst = p.out + (((long long)blockIdx.z * p.taps + tap) * p.M + m) * p.C;
          if (cb + 16 <= p.C && (p.C & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float4*>(dst + cb + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cb + i < p.C) dst[cb + i] = v[i];
          }
        } else {
          float* dst = p.out + (((long long)g * p.M + m) * p.C) * p.taps + tap;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cb + i < p.C) atomicAdd(dst + (long long)(cb + i) * p.taps, v[i]);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&hdr->tmem_empty));
      ++flushes;
      si = se;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// ================================================================================================ weight gradient, CTA pair
// tc_wgrad_kernel contracts a 128 x 320 output tile per CTA: 56 KB of operands per 64 positions for 2.6 M MACs, i.e. the
// kernel pulls ~10 TB/s from L2 (ncu: 0.8 - 1.4 GB per launch against 0.12 - 0.32 GB from DRAM) and stalls on it.  Here a
// cluster of two CTAs computes a 256-row tile with tcgen05.mma.cta_group::2 and each CTA stages only HALF of the N
// operand: 36-40 KB per CTA for the same MACs.  Roles are swapped with respect to the single-CTA kernel so that the
// 256-row M dimension is filled without padding: M = (tap, input channel) -- taps * ceil(C/64) atoms of 64 channels, four
// atoms per cluster, each with its own tap shift on the TMA coordinate -- and N = output channels in UMMA tiles of 256 /
// 128 (320 = 256 + 128 with the last 64 zero-filled; 640 = [256 + 128] + [256]; 1024 = 2 x [256 + 256]).
// D[(tap, c)][m] = sum_pos X[pos + off(tap)][c] * dU[pos][m]; both operands MN-major straight from the PC layout.
struct Wgrad2Params {
  int T, M, C, taps, dil, B, kchunks, catoms, atoms_total, pair_tiles, n_groups, group_cols, samples_per_cta, stages;
  uint32_t tmem_cols;
  float* out;   // partial[slot][tap][C][M]
};

struct HeaderW2 {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint32_t tmem_slot;
  uint32_t pad;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad2_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmX,
                 const Wgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  HeaderW2* hdr = reinterpret_cast<HeaderW2*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t atom_bytes = kBlockK * 128;                // 64 positions x 64 channels of bf16
  const int cluster = blockIdx.x >> 1;
  const int pt = cluster % p.pair_tiles;
  const int ng = (cluster / p.pair_tiles) % p.n_groups;
  const int split = cluster / (p.pair_tiles * p.n_groups);
  const int s0 = split * p.samples_per_cta;
  const int s1 = min(s0 + p.samples_per_cta, p.B);
  // output-channel tiles of this group: 256-wide while at least 256 columns remain, then one 128-wide tile
  const int n0 = ng * p.group_cols;
  int ncols = p.M - n0;
  ncols = ncols > p.group_cols ? p.group_cols : (ncols + 127) / 128 * 128;
  const int n_tiles = (ncols + 255) / 256;
  const int b_atoms = ncols / 128;                          // per CTA: two atoms per 256-wide tile, one per 128-wide tile
  const uint32_t a_bytes = 2 * atom_bytes;
  const uint32_t stage_bytes = a_bytes + (uint32_t)b_atoms * atom_bytes;
  uint8_t* ring = smem + kHeaderBytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);      // the leader's producer arrives once and expects both CTAs' bytes
      mbar_init(smem_u32(&hdr->empty[s]), 1);     // one multicast tcgen05.commit per use
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;
  const int total_it = (s1 - s0) * p.kchunks;

  if (warp == 0) {
    if (elect_one()) {
      // this CTA's two M atoms: (tap, 64 input channels) each; atoms past the end read a sample index out of range (zero fill)
      int a_tap[2], a_c0[2];
      bool a_ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int a = 4 * pt + 2 * (int)rank + u;
        a_ok[u] = a < p.atoms_total;
        a_tap[u] = a_ok[u] ? a / p.catoms : 0;
        a_c0[u] = a_ok[u] ? (a % p.catoms) * 64 : 0;
      }
      int it = 0;
      for (int b = s0; b < s1; ++b) {
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
          const uint32_t full = smem_u32(&hdr->full[s]);
          if (leader) mbar_arrive_expect_tx(full, 2 * stage_bytes);
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
#pragma unroll
          for (int u = 0; u < 2; ++u)
            tma_load_3d_2sm(sa + (uint32_t)u * atom_bytes, &tmX, a_c0[u],
                            kc * kBlockK + (a_tap[u] - p.taps / 2) * p.dil, a_ok[u] ? b : 0x3fffffff, full);
          // N operand: tile t of width w: this CTA holds channels n0 + 256 t + rank * (w / 2) + [0, w / 2)
          int atom = 0;
          for (int t = 0; t < n_tiles; ++t) {
            const int w = (ncols - 256 * t) >= 256 ? 256 : 128;
            const int ch0 = n0 + 256 * t + (int)rank * (w / 2);
            for (int u = 0; u < w / 128; ++u, ++atom)
              tma_load_3d_2sm(sa + a_bytes + (uint32_t)atom * atom_bytes, &tmU, ch0 + 64 * u, kc * kBlockK, b, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      for (int it = 0; it < total_it; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        // MN-major: LBO = distance between 64-channel atoms, SBO = distance between 8-position groups
        const uint64_t da = make_smem_desc(sa, atom_bytes, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          int atom = 0;
          for (int t = 0; t < n_tiles; ++t) {
            const int w = (ncols - 256 * t) >= 256 ? 256 : 128;
            const uint64_t db = make_smem_desc(sa + a_bytes + (uint32_t)atom * atom_bytes, atom_bytes, 1024);
            umma_bf16_2sm(tmem_base + (uint32_t)(256 * t), desc_advance(da, k * kUmmaK * 128),
                          desc_advance(db, k * kUmmaK * 128), make_idesc_m(256, w, 1, 1), (it | k) != 0);
            atom += w / 128;
          }
        }
        umma_commit_2sm(smem_u32(&hdr->empty[s]), 3);
      }
      umma_commit_2sm(smem_u32(&hdr->tmem_full), 3);
    }
  } else {
    // ---------------- epilogue (both CTAs): thread = one accumulator row = (tap, input channel); columns = output channels
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int a = 4 * pt + 2 * (int)rank + (r >> 6);
    const bool ok = a < p.atoms_total;
    const int tap = ok ? a / p.catoms : 0;
    const int c = ok ? (a % p.catoms) * 64 + (r & 63) : 0;
    const bool valid = ok && c < p.C && total_it > 0;
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    // partial layout of this kernel: [slot][tap][C][M] -- the thread's accumulator row (one input channel) is contiguous
    // in the output channels it holds, so it leaves as 16-byte stores (a [..][M][C] layout would need one 4-byte store
    // per element: 49 k store instructions per CTA, as long as the main loop)
    float* dst = p.out + (((long long)split * p.taps + tap) * p.C + c) * p.M;
    const bool vec = (p.M & 3) == 0;
    for (int ch = 0; ch < ncols / 16; ++ch) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 16), v);
      if (!valid) continue;
      const int nb = n0 + ch * 16;
      if (vec && nb + 16 <= p.M) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(dst + nb + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (nb + i < p.M) dst[nb + i] = v[i];
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  cluster_sync_all();                              // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, p.tmem_cols);
  }
}

// partial[slot][tap][M][C] -> dW[m][c][tap] += sum_slot.  Thread = one (m, c): coalesced slot reads (all independent,
// unrolled for memory-level parallelism) and one contiguous run of `taps` floats written per thread.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, int slots, int taps, int M,
                                    int C, int cmajor) {
  const long long mc = (long long)M * C;
  const long long n = (long long)taps * mc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < mc;
       idx += (long long)gridDim.x * blockDim.x) {
    // cmajor = 0: partial[slot][tap][M][C] (single-CTA kernel), idx = m * C + c
    // cmajor = 1: partial[slot][tap][C][M] (CTA-pair kernel),   idx = c * M + m: coalesced reads either way
    const long long o = cmajor ? ((idx % M) * C + idx / M) : idx;      // m * C + c
    for (int tap = 0; tap < taps; ++tap) {
      const float* src = part + (long long)tap * mc + idx;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int k = 0;
      for (; k + 4 <= slots; k += 4) {
        s0 += src[(long long)k * n];
        s1 += src[(long long)(k + 1) * n];
        s2 += src[(long long)(k + 2) * n];
        s3 += src[(long long)(k + 3) * n];
      }
      for (; k < slots; ++k) s0 += src[(long long)k * n];
      dW[o * taps + tap] += (s0 + s1) + (s2 + s3);
    }
  }
}

// ================================================================================================ CLIP logits
struct LogitsParams {
  int By, Bz, NT, stages, kchunks_total, kchunks_per_cta, i_tiles;
  uint32_t tmem_cols, idesc;
  float* raw;
};

__global__ void __launch_bounds__(kThreads)
tc_clip_logits_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmZ,
                      const LogitsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t a_bytes = kTileM * 128, b_bytes = (uint32_t)p.NT * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* ring = smem + kHeaderBytes;
  const int i0 = (blockIdx.x % p.i_tiles) * kTileM;
  const int j0 = (blockIdx.x / p.i_tiles) * p.NT;
  const int kc0 = blockIdx.y * p.kchunks_per_cta;
  const int kc1 = min(kc0 + p.kchunks_per_cta, p.kchunks_total);
  const int num_k = kc1 - kc0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmZ);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (num_k > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < num_k; ++it) {
          const int s = it % p.stages;
          mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
          const uint32_t full = smem_u32(&hdr->full[s]);
          mbar_arrive_expect_tx(full, stage_bytes);
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
          tma_load_2d(sa, &tmY, (kc0 + it) * kBlockK, i0, full);
          tma_load_2d(sa + a_bytes, &tmZ, (kc0 + it) * kBlockK, j0, full);
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        for (int it = 0; it < num_k; ++it) {
          const int s = it % p.stages;
          mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
          const uint64_t da = make_smem_desc(sa, 16, 1024);
          const uint64_t db = make_smem_desc(sa + a_bytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * 32), p.idesc, (it | k) != 0);
          umma_commit(smem_u32(&hdr->empty[s]));
        }
        umma_commit(smem_u32(&hdr->tmem_full));
      }
    } else {
      mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
      tc_fence_after();
      const int q = warp & 3;
      const int i = i0 + q * 32 + lane;
      for (int ch = 0; ch < p.NT / 16; ++ch) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 16), v);
        if (i >= p.By) continue;
        float* dst = p.raw + (long long)i * p.Bz + j0 + ch * 16;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (j0 + ch * 16 + e < p.Bz) atomicAdd(dst + e, v[e]);
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ CLIP dZ
struct DzParams {
  int By, Bz, NT /*d columns per CTA, multiple of 64*/, stages, kchunks, swap_grid, jt /*j tiles of 128 rows per CTA*/;
  long long D;
  uint32_t tmem_cols, idesc;
  const __nv_bfloat16* Z;
  const float* cvec;
  const float* gscale;
  __nv_bfloat16* dZ;
};

// Large-K form (By > 512: every >= 4-GPU job, and the GOD shards).  A CTA owns NT d-columns for `jt` tiles of 128 brain
// rows: the Y tile of a k step (the big operand: By x D is read once per j tile) is staged ONCE and contracted against
// every j tile's coefficient block into its own TMEM accumulator (jt * NT <= 512 columns).  With jt = 2 at the 8-GPU
// speech shape (256 local brains) Y is fetched once instead of twice: the kernel sits close to both floors there
// (1.9 GB of DRAM traffic ~ 290 us, 386 GFLOP ~ 290 us), so operand re-reads show directly in its time.
__global__ void __launch_bounds__(kThreads)
tc_clip_dz_kernel(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY,
                  const DzParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t atom_bytes = kBlockK * 128;
  const uint32_t a_bytes = kTileM * 128;
  const uint32_t b_bytes = (uint32_t)(p.NT / 64) * atom_bytes;
  const uint32_t stage_bytes = (uint32_t)p.jt * a_bytes + b_bytes;
  uint8_t* ring = smem + kHeaderBytes;
  // blockIdx.x = j group (fastest) when the launcher could swap: CTAs sharing a Y tile run together (L2 hit on the re-read)
  const long long d0 = (long long)(p.swap_grid ? blockIdx.y : blockIdx.x) * p.NT;
  const int j0 = (p.swap_grid ? blockIdx.x : blockIdx.y) * kTileM * p.jt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
        const uint32_t full = smem_u32(&hdr->full[s]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        for (int t = 0; t < p.jt; ++t)                                                    // coef[j][i]: K-major
          tma_load_2d(sa + (uint32_t)t * a_bytes, &tmC, it * kBlockK, j0 + t * kTileM, full);   // rows >= Bz: zero filled
        for (int a = 0; a < p.NT / 64; ++a)                                               // Y[i][d]: MN-major
          tma_load_2d(sa + (uint32_t)p.jt * a_bytes + a * atom_bytes, &tmY, (int)(d0 + a * 64), it * kBlockK, full);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        const uint64_t db = make_smem_desc(sa + (uint32_t)p.jt * a_bytes, atom_bytes, 1024);
        for (int t = 0; t < p.jt; ++t) {
          const uint64_t da = make_smem_desc(sa + (uint32_t)t * a_bytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(tmem_base + (uint32_t)(t * p.NT), desc_advance(da, k * 32), desc_advance(db, k * kUmmaK * 128),
                      p.idesc, (it | k) != 0);
        }
        umma_commit(smem_u32(&hdr->empty[s]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    const int q = warp & 3;
    const float gs = p.gscale[0];
    for (int t = 0; t < p.jt; ++t) {
      const int j = j0 + t * kTileM + q * 32 + lane;
      const bool valid = j < p.Bz;
      const float cj = valid ? p.cvec[j] : 0.f;
      for (int ch = 0; ch < p.NT / 16; ++ch) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.NT + ch * 16), v);
        if (!valid) continue;
        const long long d = d0 + ch * 16;
        if (d >= p.D) continue;   // D is a multiple of 16 on this path
        float z[16];
        load8(p.Z + (long long)j * p.D + d, reinterpret_cast<float(&)[8]>(z[0]));
        load8(p.Z + (long long)j * p.D + d + 8, reinterpret_cast<float(&)[8]>(z[8]));
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = gs * (v[e] - z[e] * cj);
        store8(p.dZ + (long long)j * p.D + d, reinterpret_cast<float(&)[8]>(z[0]));
        store8(p.dZ + (long long)j * p.D + d + 8, reinterpret_cast<float(&)[8]>(z[8]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Small-K form of the dZ contraction (K = By <= 512: the single-GPU step and 2-GPU jobs).  There the kernel is bound by
// the Y / Z / dZ streams (566 MB at config 2), not by the MMAs, so everything that touches global memory is a TMA tile:
// the Z tile of the CTA is fetched into swizzled shared memory while the k loop runs, the epilogue rewrites it IN PLACE
// (thread = row: 16-byte swizzled chunks) and one thread stores it back with TMA.  The first form read Z and wrote dZ
// with per-thread 32-byte global accesses strided by D (2.4 TB/s).  128 x 128 tiles, 96 KB: two CTAs per SM overlap
// one's epilogue with the other's loads.
struct DzSmallHeader {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t tmem_full;
  uint64_t zfull;
  uint32_t tmem_slot;
  uint32_t pad;
};

__global__ void __launch_bounds__(kThreads)
tc_clip_dz_small_kernel(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY,
                        const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmDZ,
                        const DzParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  DzSmallHeader* hdr = reinterpret_cast<DzSmallHeader*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  constexpr int kNT = 128;
  const uint32_t atom_bytes = kBlockK * 128;                // 64 k-rows x 128 B (= 64 d columns)
  const uint32_t a_bytes = kTileM * 128;                    // coef tile: 128 j rows x 64 k
  const uint32_t b_bytes = (kNT / 64) * atom_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;           // 32 KB
  uint8_t* ring = smem + kHeaderBytes;
  uint8_t* zt = ring + (size_t)p.stages * stage_bytes;      // Z / dZ tile: 2 slices of [128 rows][64 cols] bf16
  const uint32_t zslice = kTileM * 128;
  // blockIdx.x = j tile (fastest): the CTAs that share a Y tile are scheduled together, so its second read is an L2 hit
  const long long d0 = (long long)blockIdx.y * kNT;
  const int j0 = blockIdx.x * kTileM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmZ);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    mbar_init(smem_u32(&hdr->zfull), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), kNT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t zf = smem_u32(&hdr->zfull);
      mbar_arrive_expect_tx(zf, 2 * zslice);
      tma_load_2d(smem_u32(zt), &tmZ, (int)d0, j0, zf);
      tma_load_2d(smem_u32(zt) + zslice, &tmZ, (int)d0 + 64, j0, zf);
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
        const uint32_t full = smem_u32(&hdr->full[s]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        tma_load_2d(sa, &tmC, it * kBlockK, j0, full);                                    // coef[j][i]: K-major
        for (int a = 0; a < kNT / 64; ++a)                                                // Y[i][d]: MN-major
          tma_load_2d(sa + a_bytes + a * atom_bytes, &tmY, (int)(d0 + a * 64), it * kBlockK, full);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc(sa, 16, 1024);
        const uint64_t db = make_smem_desc(sa + a_bytes, atom_bytes, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * kUmmaK * 128), p.idesc, (it | k) != 0);
        umma_commit(smem_u32(&hdr->empty[s]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    const int q = warp & 3;
    const int rloc = q * 32 + lane;                 // row of the tile = TMEM lane
    const int j = j0 + rloc;
    const float cj = j < p.Bz ? p.cvec[j] : 0.f;
    const float gs = p.gscale[0];
    mbar_wait_sleep(smem_u32(&hdr->zfull), 0);
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    const uint32_t sw = (uint32_t)rloc & 7u;
#pragma unroll
    for (int g = 0; g < kNT / 32; ++g) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), v);
      const uint32_t rowaddr = smem_u32(zt) + (uint32_t)(g >> 1) * zslice + (uint32_t)rloc * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t addr = rowaddr + ((((uint32_t)(g & 1) * 4u + (uint32_t)c) ^ sw) << 4);
        uint4 zz;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(zz.x), "=r"(zz.y), "=r"(zz.z), "=r"(zz.w) : "r"(addr));
        const __nv_bfloat162* zh = reinterpret_cast<const __nv_bfloat162*>(&zz);
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 zf2 = __bfloat1622float2(zh[e]);
          o[e] = pack_bf16x2(gs * (v[8 * c + 2 * e] - zf2.x * cj), gs * (v[8 * c + 2 * e + 1] - zf2.y * cj));
        }
        st_shared_v4(addr, o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (warp == 2 && elect_one()) {
      tma_store_2d(&tmDZ, smem_u32(zt), (int)d0, j0);
      tma_store_2d(&tmDZ, smem_u32(zt) + zslice, (int)d0 + 64, j0);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kNT);
  }
}


// ================================================================================================ CLIP, fused large-batch form
// Static embeddings (D = F <= a few thousand: the GOD / seq2static configs) at large batch: the B x B logits matrix is
// never written.  One CTA contracts a 128 x 128 tile over the WHOLE (short) K = D and its epilogue works on the tile in
// TMEM / registers:
//   mode 0 (forward):  L = s * y_i.z_j / (|y_i||z_j|); per-row online (max, sum exp) over the tile's columns, per-column
//                      (max, sum exp) over its rows (warp shuffles, then the four row quarters meet in shared memory),
//                      the positives L[off+j][j].  Partials go to [j_tiles][2][By] / [i_tiles][2][Bz] tables that
//                      clip_lse_merge_kernel completes (the same kernel that merges the ranks' partials in data parallel).
//   mode 1 (backward): the tile is RECOMPUTED, dL = (exp(L - row_lse) + exp(L - col_lse) - 2 [positive]) * wscale is formed
//                      in registers, and what leaves is the bf16 coefficient tile of the dZ contraction (transposed,
//                      for a chunk of brain columns at a time), sum_i dL*L per column and d loss / d temp.
// loss.py:92-107 with SURVEY appendix A5/A6 (one-sided data-parallel form: rows = all speech targets, columns = local brains).
struct ClipFusedParams {
  int By, Bz, kchunks, stages, i_tiles, mode, off, ldc, j_begin, j_end;
  float wscale;
  uint32_t idesc;
  const float* ysq;
  const float* zsq;
  const float* temp;
  float* row_part;
  float* col_part;
  float* diag;
  const float* row_lse;
  const float* col_lse;
  __nv_bfloat16* coef;
  float* cvec;
  float* out2;
};

constexpr int kFusedNT = 128;

__global__ void __launch_bounds__(kThreads)
tc_clip_fused_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmZ,
                     const ClipFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  Header* hdr = reinterpret_cast<Header*>(smem);
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t a_bytes = kTileM * 128, b_bytes = kFusedNT * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* ring = smem + kHeaderBytes;
  float* srz = reinterpret_cast<float*>(ring + (size_t)p.stages * stage_bytes);   // [NT] 1/|z_j|
  float* scl = srz + kFusedNT;                                                    // [NT] col_lse (mode 1)
  float* scm = scl + kFusedNT;                                                    // [4][NT] per-warp column max / dL*L sums
  float* scs = scm + 4 * kFusedNT;                                                // [4][NT] per-warp column sum exp
  const int i_tile = blockIdx.x % p.i_tiles, j_tile = blockIdx.x / p.i_tiles;
  const int i0 = i_tile * kTileM;
  const int j0 = p.j_begin + j_tile * kFusedNT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmZ);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), 1);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&hdr->tmem_slot), kFusedNT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->empty[s]), ((it / p.stages) & 1) ^ 1);
        const uint32_t full = smem_u32(&hdr->full[s]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        tma_load_2d(sa, &tmY, it * kBlockK, i0, full);
        tma_load_2d(sa + a_bytes, &tmZ, it * kBlockK, j0, full);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int it = 0; it < p.kchunks; ++it) {
        const int s = it % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc(sa, 16, 1024);
        const uint64_t db = make_smem_desc(sa + a_bytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * 32), p.idesc, (it | k) != 0);
        umma_commit(smem_u32(&hdr->empty[s]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else {
    const int et = threadIdx.x - 64;                // 0..127
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;               // this thread's speech row (TMEM lane)
    const bool iv = i < p.By;
    {
      const int j = j0 + et;
      const bool jv = j < p.j_end;
      srz[et] = jv ? rsqrtf(p.zsq[j]) : 0.f;
      scl[et] = (jv && p.mode == 1) ? p.col_lse[j] : 0.f;
      if (p.mode == 1) {
#pragma unroll
        for (int w = 0; w < 4; ++w) scm[w * kFusedNT + et] = 0.f;
      }
    }
    const float s = expf(p.temp[0]);
    const float ri = iv ? s * rsqrtf(p.ysq[i]) : 0.f;
    const float rl = (iv && p.mode == 1) ? p.row_lse[i] : 0.f;
    named_bar_sync(1, 128);
    mbar_wait_sleep(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    float rm = -INFINITY, rs = 0.f, dtemp = 0.f;
    for (int g = 0; g < kFusedNT / 32; ++g) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * g), v);
      if (j0 + 32 * g >= p.j_end) continue;         // whole group outside the column range (uniform across the CTA)
      if (p.mode == 0) {
        float gm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const bool ok = iv && (j0 + 32 * g + e) < p.j_end;
          v[e] = ok ? v[e] * ri * srz[32 * g + e] : -INFINITY;
          gm = fmaxf(gm, v[e]);
        }
        if (gm > -INFINITY) {                        // row-wise online log-sum-exp over this group
          const float nm = fmaxf(rm, gm);
          float acc = rs * __expf(rm - nm);          // rm = -inf: rs = 0 and exp(-inf) = 0
#pragma unroll
          for (int e = 0; e < 32; ++e) acc += __expf(v[e] - nm);      // masked entries: exp(-inf) = 0
          rm = nm;
          rs = acc;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) {               // column-wise over this warp's 32 rows
          const float cm = warp_max(v[e]);
          const float cs = warp_sum(v[e] > -INFINITY ? __expf(v[e] - cm) : 0.f);
          if (elect_one()) {
            scm[q * kFusedNT + 32 * g + e] = cm;
            scs[q * kFusedNT + 32 * g + e] = cs;
          }
          if (iv && i == p.off + j0 + 32 * g + e && v[e] > -INFINITY) p.diag[j0 + 32 * g + e] = v[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + 32 * g + e;
          const bool ok = iv && j < p.j_end;
          float t = 0.f;
          if (ok) {
            const float l = v[e] * ri * srz[32 * g + e];
            const float dl = (__expf(l - rl) + __expf(l - scl[32 * g + e]) - (i == p.off + j ? 2.f : 0.f)) * p.wscale;
            t = dl * l;
            p.coef[(long long)(j - p.j_begin) * p.ldc + i] = __float2bfloat16_rn(dl * ri * srz[32 * g + e]);
          }
          dtemp += t;
          const float cs = warp_sum(t);
          if (lane == 0) scm[q * kFusedNT + 32 * g + e] = cs;
        }
      }
    }
    tc_fence_before();
    if (p.mode == 0) {
      if (iv) {
        p.row_part[((long long)j_tile * 2) * p.By + i] = rm;
        p.row_part[((long long)j_tile * 2 + 1) * p.By + i] = rs;
      }
      named_bar_sync(1, 128);
      const int j = j0 + et;
      if (j < p.j_end) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, scm[w * kFusedNT + et]);
        float S = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float m = scm[w * kFusedNT + et];
          if (m > -INFINITY) S += scs[w * kFusedNT + et] * __expf(m - M);
        }
        p.col_part[((long long)i_tile * 2) * p.Bz + j] = M;
        p.col_part[((long long)i_tile * 2 + 1) * p.Bz + j] = S;
      }
    } else {
      named_bar_sync(1, 128);
      const int j = j0 + et;
      if (j < p.j_end) {
        const float c = scm[et] + scm[kFusedNT + et] + scm[2 * kFusedNT + et] + scm[3 * kFusedNT + et];
        atomicAdd(&p.cvec[j], c * srz[et] * srz[et]);
      }
      dtemp = warp_sum(dtemp);
      if (lane == 0 && dtemp != 0.f) atomicAdd(&p.out2[1], dtemp);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kFusedNT);
  }
}

// loss = wscale * sum_j (row_lse[off + j] + col_lse[j] - 2 L[off + j][j])  (loss.py:107 on the completed log-sum-exps)
__global__ void clip_fused_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                                       const float* __restrict__ diag, int off, int Bz, float wscale,
                                       float* __restrict__ out2) {
  float a = 0.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Bz; j += gridDim.x * blockDim.x)
    a += (row_lse[off + j] + col_lse[j] - 2.f * diag[j]) * wscale;
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0 && a != 0.f) atomicAdd(&out2[0], a);
}

// ================================================================================================ host launchers
static bool halo_mode_env_off() { return options().tapgemm_halo == 0; }

static int pick_stages(uint32_t stage_bytes, uint32_t budget, int max_stages) {
  int s = (int)(budget / stage_bytes);
  if (s > max_stages) s = max_stages;
  return s;
}

static int pick_ntile(int Np) {
  if (Np <= 256) return Np;
  for (int nt = 256; nt >= 128; nt -= 16)
    if (Np % nt == 0) return nt;
  const int tiles = (Np + 255) / 256;
  return (((Np + tiles - 1) / tiles) + 15) / 16 * 16;
}

// Raises the kernel's dynamic shared-memory limit when needed.  The attribute only ever grows and is remembered per
// kernel (and device), so steady-state launches -- and launches recorded into a CUDA graph after a warm-up step -- make
// no attribute call at all.
template <class K>
static int set_smem(K kernel, size_t bytes) {
  struct Entry { const void* fn; int dev; size_t bytes; };
  static Entry table[64];
  static int used = 0;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const void* fn = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  Entry* e = nullptr;
  for (int i = 0; i < used; ++i)
    if (table[i].fn == fn && table[i].dev == dev) e = &table[i];
  if (e != nullptr && bytes <= e->bytes) return 0;
  MEGB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (e == nullptr && used < 64) e = &table[used++], e->fn = fn, e->dev = dev, e->bytes = 0;
  if (e != nullptr) e->bytes = bytes;
  return 0;
}

int tc_tapgemm(const void* A, int B, int T, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps, int dil,
               const float* bias, const void* Res, int Crs, const int32_t* group, int G, int epi, void* Out, int Cos,
               void* Out2, int Co2s, float* stats, int* stats_fused, cudaStream_t st) {
  *stats_fused = 0;
  MEGB_CHECK_ARG(Cas % 8 == 0 && Cos % 8 == 0 && (Res == nullptr || Crs % 8 == 0), "tc_tapgemm: strides must be multiples of 8");
  MEGB_CHECK_ARG(Kp % 64 == 0 && Np % 16 == 0 && Np >= N, "tc_tapgemm: Kp %% 64 and Np %% 16 required (Kp=%d Np=%d)", Kp, Np);
  const bool glu = epi == MEGB200_EPI_GLU;
  TapGemmParams p;
  // Per-phase clock64 accounting is a developer build only (-DMEGB200_DEBUG, then MEGB200_DEBUG_TIMING=1): the
  // production library carries no debug mode that could change results or leak memory.
#ifdef MEGB200_DEBUG
  static const bool dbg_timing = getenv("MEGB200_DEBUG_TIMING") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_timing && dbg_buf == nullptr) cudaMalloc(&dbg_buf, sizeof(long long) * 16 * 65536);
#else
  constexpr bool dbg_timing = false;
  long long* const dbg_buf = nullptr;
#endif
  p.dbg = nullptr;
  p.T = T; p.N = N; p.Cos = Cos; p.Crs = Crs; p.Co2s = Co2s; p.taps = taps; p.dil = dil; p.Kp = Kp; p.epi = epi;
  p.kchunks = (Ca + kBlockK - 1) / kBlockK;
  p.tiles_per_sample = (T + kTileM - 1) / kTileM;
  p.tma_epi = options().tapgemm_direct_epi ? 0 : 1;
  p.prefetch = options().tapgemm_l2_prefetch;
  p.ksub = 2;
  p.stg_bytes = 0;
  if (glu) {
    const int Ch = N / 2;
    MEGB_CHECK_ARG(N % 16 == 0 && Res == nullptr && Out2 != nullptr && Co2s % 8 == 0 && Co2s >= N,
                   "tc_tapgemm: the GLU epilogue needs N %% 16 == 0, no residual and a V output (N=%d)", N);
    // tile = [half a-channels | half gate channels], half a multiple of 16 dividing Ch if possible
    int half = 0;
    if (p.tma_epi) {                       // TMA-store epilogue works on 32-channel slices
      for (int h = 64; h >= 32; h -= 32)
        if (Ch % h == 0) { half = h; break; }
      if (half == 0) p.tma_epi = 0;
    }
    if (half == 0)
      for (int h = 128; h >= 16; h -= 16)
        if (Ch % h == 0 && 2 * h <= 160) { half = h; break; }
    if (half == 0) half = Ch >= 80 ? 80 : (Ch + 15) / 16 * 16;
    p.NT = 2 * half;
    p.n_last = p.NT;
    const int cover = Cos > Ch ? Cos : Ch;
    p.ntiles = (cover + half - 1) / half;
  } else {
    MEGB_CHECK_ARG(epi != MEGB200_EPI_GELU || (Out2 != nullptr && Co2s % 8 == 0 && Co2s >= N),
                   "tc_tapgemm: the GELU epilogue needs a pre-activation output");
    // output channels up to Cos must be covered so that pad channels are written as zero
    const int ncover = Cos > Np ? Cos : Np;
    // MEGB200_TAPGEMM_RAGGED_N=1: 256-wide tiles plus one narrower tail (320 = 256 + 64).  Measured slower than uniform
    // 160-wide tiles at config 2 (576 vs 647 TFLOP/s: the 48 KB stages leave a 2-deep ring), so it is off by default.
    const bool uniform_tiles = options().tapgemm_ragged_n == 0;
    const int cover32 = (ncover + 31) / 32 * 32;
    if (!uniform_tiles && cover32 > 256 && halo_mode_env_off()) {
      p.NT = 256;
      p.ntiles = (cover32 + 255) / 256;
      p.n_last = cover32 - 256 * (p.ntiles - 1);
    } else {
      p.NT = pick_ntile(Np);
      p.ntiles = (ncover + p.NT - 1) / p.NT;
      p.n_last = p.NT;
    }
    if (p.NT % 32 != 0) p.tma_epi = 0;
  }
  TapStoreMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (p.tma_epi) {
    const int eb = epi == MEGB200_EPI_STORE_F32 ? 4 : 2;
    {
      const uint64_t dims[3] = {(uint64_t)Cos, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Cos * eb, (uint64_t)T * Cos * eb};
      const uint32_t box[3] = {32, kTileM, 1};
      int rc = encode_tmap(&maps.out, Out, eb, eb == 4 ? 128 : 64, 3, dims, strides, box);
      if (rc) return rc;
    }
    if (Out2 != nullptr) {
      const uint64_t dims[3] = {(uint64_t)Co2s, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Co2s * 2, (uint64_t)T * Co2s * 2};
      const uint32_t box[3] = {32, kTileM, 1};
      int rc = encode_tmap(&maps.out2, Out2, 2, 64, 3, dims, strides, box);
      if (rc) return rc;
    }
  }
  p.bias = bias; p.Res = (const __nv_bfloat16*)Res; p.group = group; p.Out = Out; p.Out2 = (__nv_bfloat16*)Out2;
  // BatchNorm statistics ride on the f32 TMA-store epilogue (single-CTA and pair kernels, not the halo forms)
  const bool fuse_stats = stats != nullptr && epi == MEGB200_EPI_STORE_F32 && p.tma_epi && halo_mode_env_off();
  p.stats = fuse_stats ? stats : nullptr;

  // Kernel choice.  Measured at config 2 (profiles/r1_tapgemm_variants.md): the persistent CTA-pair kernel (8 epilogue
  // warps, 128-wide k stages) reaches 647 TFLOP/s, the single-CTA kernel with two co-resident CTAs per SM 608; the pair
  // kernel is used for large ungrouped problems.  MEGB200_TAPGEMM_1CTA=1 forces the single-CTA kernel,
  // MEGB200_TAPGEMM_2CTA_MIN_ITEMS=<n> lowers the pair kernel's work threshold (the tests exercise it on small shapes).
  const Options& opt = options();
  const bool force_1cta = opt.tapgemm_force_1cta != 0;
  const long long min_items = opt.tapgemm_pair_min_items >= 0 ? opt.tapgemm_pair_min_items : 2ll * num_sms();
  const long long m_tiles_total = (long long)B * p.tiles_per_sample;
  if (group == nullptr && G == 1 && !force_1cta && p.tma_epi && p.NT % 32 == 0 && m_tiles_total * p.ntiles >= min_items) {
    // ---- CTA-pair persistent kernel
    const int half = p.NT / 2;
    const uint32_t stage2 = kTileM * 128 + half * 128;
    // Staging area of the TMA-store epilogue: 64 KB, or the whole tile (one staging round per tile instead of two: the
    // f32 and GELU epilogues of a 160-channel tile need 80 KB) when that costs no pipeline stage.  kSmemMax = the
    // 227 KB a block may opt into.
    constexpr uint32_t kSmemMax = 232448;
    uint32_t want_stg = 0;
    auto grow_staging = [&](uint32_t stage_bytes, uint32_t tables, int max_stages) {
      const int now = pick_stages(stage_bytes, kSmemMax - 2048 - tables - p.stg_bytes, max_stages);
      if (want_stg > p.stg_bytes && want_stg <= 96u * 1024 &&
          pick_stages(stage_bytes, kSmemMax - 2048 - tables - want_stg, max_stages) == now)
        p.stg_bytes = want_stg;
      return pick_stages(stage_bytes, kSmemMax - 2048 - tables - p.stg_bytes, max_stages);
    };
    if (p.tma_epi) {
      const uint32_t slice = epi == MEGB200_EPI_STORE_F32 ? 16384u : 8192u;
      const int per_group = glu ? 3 : (epi == MEGB200_EPI_GELU ? 2 : 1);
      want_stg = (uint32_t)((glu ? half : p.NT) / 32) * per_group * slice;
      p.stg_bytes = want_stg > 64u * 1024 ? 64u * 1024 : want_stg;
    }
    p.stages = pick_stages(stage2, 229000 - 2048 - 2 * p.NT * 4 - p.stg_bytes, 8);
    p.tmem_cols = tmem_cols_for(2 * p.NT);
    p.idesc = make_idesc_m(256, p.NT, 0, 0);
    CUtensorMap tmA2, tmW2;
    {
      const uint64_t dims[3] = {(uint64_t)Ca, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Cas * 2, (uint64_t)T * Cas * 2};
      const uint32_t box[3] = {kBlockK, kTileM, 1};
      int rc = encode_tmap_bf16(&tmA2, A, 3, dims, strides, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[3] = {(uint64_t)taps * Kp, (uint64_t)Np, 1};
      const uint64_t strides[2] = {(uint64_t)taps * Kp * 2, (uint64_t)Np * taps * Kp * 2};
      const uint32_t box[3] = {kBlockK, (uint32_t)half, 1};
      int rc = encode_tmap_bf16(&tmW2, W, 3, dims, strides, box);
      if (rc) return rc;
    }
    const int m_pairs = (int)((m_tiles_total + 1) / 2);
    const int num_items = m_pairs * p.ntiles;
    int clusters = num_sms() / 2;
    if (clusters > num_items) clusters = num_items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreadsPair);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto report_pair_timing = [&](const char* which) {
      static long long hostp[16 * 512];
      cudaStreamSynchronize(st);
      cudaMemcpy(hostp, dbg_buf, sizeof(long long) * 16 * 2 * clusters, cudaMemcpyDeviceToHost);
      double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < clusters; ++c) {
        const long long* l = hostp + 16 * (2 * c);         // leader CTA of the pair
        const long long* f = hostp + 16 * (2 * c + 1);     // follower
        v[0] += (double)(l[0] + f[0]) / 2;                 // producer wait (empty)
        v[1] += (double)l[1];                              // MMA loop total
        v[2] += (double)l[2];                              // MMA wait full
        v[3] += (double)l[3];                              // MMA wait accumulator
        v[4] += (double)(l[4] + f[4]) / 2;                 // epilogue wait tmem_full
        v[5] += (double)(l[5] + f[5]) / 2;                 // epilogue work
        v[6] += (double)l[6];                              // tiles per cluster
      }
      double ph[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < clusters; ++c)
        for (int i = 0; i < 6; ++i) ph[i] += (double)hostp[16 * (2 * c) + 8 + i];
      for (int i = 0; i < 7; ++i) v[i] /= clusters;
      for (int i = 0; i < 6; ++i) ph[i] /= clusters * v[6];
      fprintf(stderr, "[megb200 timing]   epilogue thread 0 per tile: wait stores-read %.0f, barrier %.0f, tmem->convert->stage %.0f, "
              "fence+barrier %.0f, store issue %.0f, stats %.0f\n", ph[0], ph[1], ph[2], ph[3], ph[4], ph[5]);
      fprintf(stderr, "[megb200 timing] %s tapgemm N=%d K=%d taps=%d epi=%d NT=%d stages=%d: per cluster %.1f tiles; MMA loop %.0f "
              "cycles (waiting for data %.0f, for a free accumulator %.0f); per tile: MMA %.0f, epilogue work %.0f, epilogue "
              "idle %.0f; producer waiting for free slots %.0f\n", which, N, Ca, taps, epi, p.NT, p.stages, v[6], v[1], v[2], v[3],
              v[1] / v[6], v[5] / v[6], v[4] / v[6], v[0]);
    };
    // halo form of the pair kernel, default for k > 1 (653 vs 599 TFLOP/s on the same box, r1_tapgemm_variants.md);
    // MEGB200_TAPGEMM_PAIR_HALO=0 selects the plain pair kernel
    const bool no_halo = opt.tapgemm_pair_halo == 0;
    const int a_rows2 = ((kTileM + (taps - 1) * dil) + 7) / 8 * 8;
    if (dbg_timing) {
      p.dbg = dbg_buf;
      cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 16 * 2 * clusters, st);
    }
    // the halo form needs two stages of (activation rows with halo + ALL taps' weight half-tiles): wide kernels (k = 11 of
    // BrainEncoderSeq2Static) and large dilations take the plain pair kernel, one activation box per tap
    const uint32_t halo_stage = (((uint32_t)a_rows2 * 128u + 1023u) & ~1023u) + (uint32_t)taps * (uint32_t)half * 128u;
    const bool halo_fits = a_rows2 <= 256 &&
                           pick_stages(halo_stage, kSmemMax - 2048 - (2 * p.NT * 4 + 2 * Cos * 4) - p.stg_bytes, 4) >= 2;
    if (taps > 1 && halo_fits && !no_halo && p.n_last == p.NT) {
      // halo form: activation ring (with halo rows) + weight ring
      const uint32_t a_slot = ((uint32_t)a_rows2 * 128u + 1023u) & ~1023u;
      const uint32_t b_half = (uint32_t)half * 128;
      const uint32_t stage_h = a_slot + (uint32_t)taps * b_half;
      const int stagesA = grow_staging(stage_h, 2 * p.NT * 4 + 2 * Cos * 4, 4);
      MEGB_CHECK_ARG(stagesA >= 2, "tc_tapgemm(pair, halo): tile too large");
      p.stages = stagesA;
      CUtensorMap tmAh;
      const uint64_t dims[3] = {(uint64_t)Ca, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Cas * 2, (uint64_t)T * Cas * 2};
      const uint32_t box[3] = {kBlockK, (uint32_t)a_rows2, 1};
      int rc = encode_tmap_bf16(&tmAh, A, 3, dims, strides, box);
      if (rc) return rc;
      const size_t smemh = 1024 + kHeaderBytes + (size_t)stagesA * stage_h + p.stg_bytes + 2 * p.NT * 4 + 2 * Cos * 4;
      rc = set_smem(tc_tapgemm2h_kernel, smemh);
      if (rc) return rc;
      cfg.dynamicSmemBytes = smemh;
      MEGB_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_tapgemm2h_kernel, tmAh, tmW2, maps, p, (int)m_tiles_total, num_items,
                                      a_rows2, stagesA));
      MEGB_LAUNCH_OK();
      *stats_fused = fuse_stats ? 1 : 0;
      options().last_tapgemm_kernel = 2;
      if (p.dbg != nullptr) report_pair_timing("pair+halo");
      return 0;
    }
    // two k chunks per stage (one barrier round trip per 8 MMAs).  One chunk per stage doubles the ring depth (the MMA
    // thread of the 1x1 head convs waits for data 32-57 % of its loop behind a two-stage ring) but measured slower:
    // 6.96 vs 6.90 ms per step -- opt-in ("tapgemm_ksub" = 1)
    p.ksub = 2;
    p.stages = grow_staging(2 * stage2, 2 * p.NT * 4 + 2 * Cos * 4, 8);
    if (options().tapgemm_ksub == 1) {
      p.ksub = 1;
      p.stages = grow_staging(stage2, 2 * p.NT * 4 + 2 * Cos * 4, 8);
    }
    MEGB_CHECK_ARG(p.stages >= 2, "tc_tapgemm(pair): tile too large");
    const size_t smem2 = 1024 + kHeaderBytes + (size_t)p.stages * p.ksub * stage2 + p.stg_bytes + 2 * p.NT * 4 + 2 * Cos * 4;
    int rc = set_smem(tc_tapgemm2_kernel, smem2);
    if (rc) return rc;
    cfg.dynamicSmemBytes = smem2;
    MEGB_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_tapgemm2_kernel, tmA2, tmW2, maps, p, (int)m_tiles_total, num_items));
    MEGB_LAUNCH_OK();
    *stats_fused = fuse_stats ? 1 : 0;
    options().last_tapgemm_kernel = 1;
    if (p.dbg != nullptr) report_pair_timing("pair");
    return 0;
  }

  // ---- single-CTA halo form for k > 1 convolutions, opt-in (MEGB200_TAPGEMM_HALO=1).  The row-shifted descriptor uses
  // base_offset = 0: the swizzle is a function of the absolute shared-memory address (setting the field gave wrong
  // results, profiles/r1_tapgemm_variants.md)
  const int halo_mode = opt.tapgemm_halo;
  const int a_rows = ((kTileM + (taps - 1) * dil) + 7) / 8 * 8;
  if (halo_mode > 0 && taps > 1 && group == nullptr && G == 1 && a_rows <= 256) {
    p.stats = nullptr;
    const uint32_t a_slot = ((uint32_t)a_rows * 128u + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.NT * 128;
    const int stagesA = 2;
    const uint32_t budget = 112000u - (p.NT > 128 ? p.NT * 4 : 0) - stagesA * a_slot;
    p.stages = pick_stages(b_bytes, budget, 8);
    MEGB_CHECK_ARG(p.stages >= 2, "tc_tapgemm(halo): tile too large");
    p.tmem_cols = tmem_cols_for(p.NT);
    p.idesc = make_idesc(p.NT, 0, 0);
    CUtensorMap tmAh, tmWh;
    {
      const uint64_t dims[3] = {(uint64_t)Ca, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Cas * 2, (uint64_t)T * Cas * 2};
      const uint32_t box[3] = {kBlockK, (uint32_t)a_rows, 1};
      int rc = encode_tmap_bf16(&tmAh, A, 3, dims, strides, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[3] = {(uint64_t)taps * Kp, (uint64_t)Np, 1};
      const uint64_t strides[2] = {(uint64_t)taps * Kp * 2, (uint64_t)Np * taps * Kp * 2};
      const uint32_t box[3] = {kBlockK, (uint32_t)(glu ? p.NT / 2 : p.NT), 1};
      int rc = encode_tmap_bf16(&tmWh, W, 3, dims, strides, box);
      if (rc) return rc;
    }
    const size_t smemh = 1024 + kHeaderBytes + (size_t)stagesA * a_slot + (size_t)p.stages * b_bytes +
                         (p.NT > 128 ? p.NT * 4 : 0);
    int rc = set_smem(tc_tapgemm_halo_kernel, smemh);
    if (rc) return rc;
    const long long blocks = (long long)B * p.tiles_per_sample * p.ntiles;
    MEGB_CHECK_ARG(blocks > 0 && blocks < 2147483647ll, "tc_tapgemm: grid too large");
    tc_tapgemm_halo_kernel<<<(unsigned)blocks, kThreads, smemh, st>>>(tmAh, tmWh, maps, p, a_rows, stagesA);
    MEGB_LAUNCH_OK();
    options().last_tapgemm_kernel = 3;
    return 0;
  }

  const uint32_t stage_bytes = kTileM * 128 + p.NT * 128;
  p.stages = pick_stages(stage_bytes, 112000 - (p.NT > 128 ? p.NT * 4 : 0) - 2 * p.NT * 4, 6);
  MEGB_CHECK_ARG(p.stages >= 2, "tc_tapgemm: tile too large for the stage ring");
  p.tmem_cols = tmem_cols_for(p.NT);
  p.idesc = make_idesc(p.NT, 0, 0);
  p.bias = bias; p.Res = (const __nv_bfloat16*)Res; p.group = group; p.Out = Out; p.Out2 = (__nv_bfloat16*)Out2;

  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[3] = {(uint64_t)Ca, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)Cas * 2, (uint64_t)T * Cas * 2};
    const uint32_t box[3] = {kBlockK, kTileM, 1};
    int rc = encode_tmap_bf16(&tmA, A, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)taps * Kp, (uint64_t)Np, (uint64_t)G};
    const uint64_t strides[2] = {(uint64_t)taps * Kp * 2, (uint64_t)Np * taps * Kp * 2};
    const uint32_t box[3] = {kBlockK, (uint32_t)(glu ? p.NT / 2 : p.NT), 1};
    int rc = encode_tmap_bf16(&tmW, W, 3, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * stage_bytes + (p.NT > 128 ? p.NT * 4 : 0) + 2 * p.NT * 4;
  int rc = set_smem(tc_tapgemm_kernel, smem);
  if (rc) return rc;
  const long long blocks = (long long)B * p.tiles_per_sample * p.ntiles;
  MEGB_CHECK_ARG(blocks > 0 && blocks < 2147483647ll, "tc_tapgemm: grid too large");
  if (dbg_timing && blocks <= 65536) p.dbg = dbg_buf;
  tc_tapgemm_kernel<<<(unsigned)blocks, kThreads, smem, st>>>(tmA, tmW, maps, p);
  MEGB_LAUNCH_OK();
  options().last_tapgemm_kernel = 0;
  *stats_fused = p.stats != nullptr ? 1 : 0;
  if (p.dbg != nullptr) {
    // debug only: phase durations per CTA in SM clocks (synchronises!)
    static long long host[16 * 65536];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, dbg_buf, sizeof(long long) * 16 * blocks, cudaMemcpyDeviceToHost);
    double fill_c = 0, main_c = 0, epi_c = 0, total_c = 0, e_stage = 0, e_bar = 0, e_issue = 0, e_wait = 0;
    for (long long i = 0; i < blocks; ++i) {
      fill_c += (double)(host[16 * i + 1] - host[16 * i + 0]);
      main_c += (double)(host[16 * i + 2] - host[16 * i + 1]);
      epi_c += (double)(host[16 * i + 3] - host[16 * i + 2]);
      total_c += (double)(host[16 * i + 3] - host[16 * i + 0]);
      e_stage += (double)(host[16 * i + 4] - host[16 * i + 2]);
      e_bar += (double)(host[16 * i + 5] - host[16 * i + 4]);
      e_issue += (double)(host[16 * i + 6] - host[16 * i + 5]);
      e_wait += (double)(host[16 * i + 3] - host[16 * i + 6]);
    }
    if (p.tma_epi)
      fprintf(stderr, "[megb200 timing]   epilogue round 0: convert+stage %.0f, fence+barrier %.0f, store issue %.0f, rest "
              "(later rounds + wait for stores to read smem) %.0f\n", e_stage / blocks, e_bar / blocks, e_issue / blocks,
              e_wait / blocks);
    fprintf(stderr, "[megb200 timing] tapgemm N=%d K=%d taps=%d epi=%d NT=%d stages=%d blocks=%lld: "
            "start->first data %.0f, main loop %.0f, epilogue %.0f, total %.0f cycles (mean per CTA)\n",
            N, Ca, taps, epi, p.NT, p.stages, blocks, fill_c / blocks, main_c / blocks, epi_c / blocks, total_c / blocks);
  }
  return 0;
}

int tc_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int T, int taps, int dil,
             const int32_t* group, const int32_t* order, int G, float* dW, float* workspace, size_t ws_bytes,
             cudaStream_t st) {
  MEGB_CHECK_ARG(Ms % 8 == 0 && Cs % 8 == 0, "tc_wgrad: strides must be multiples of 8");
  if (group == nullptr && options().wgrad_pair != 0) {
    // ---- CTA-pair form: opt-in (option "wgrad_pair").  Measured on B200 at config 2 it is SLOWER than the single-CTA
    // kernel (320->320 k=3: 92 vs 73 us; profiles/r2_wgrad_ncu.md) although it moves 37 % fewer bytes from L2: ncu shows the
    // tensor pipe busy ~90 % of the time in BOTH kernels at half the K-major MMA rate -- the weight gradient contracts over
    // positions, so both operands are MN-major in the position-major layout, and that operand form is what bounds it.
    Wgrad2Params q;
    q.T = T; q.M = M; q.C = C; q.taps = taps; q.dil = dil; q.B = B;
    q.kchunks = (T + kBlockK - 1) / kBlockK;
    q.catoms = (C + 63) / 64;
    q.atoms_total = taps * q.catoms;
    q.pair_tiles = (q.atoms_total + 3) / 4;
    const int m128 = (M + 127) / 128 * 128;
    // output channels per cluster pass: <= 512 TMEM columns, balanced over the groups (640 -> 384 + 256, 1024 -> 512 + 512)
    q.n_groups = (m128 + 511) / 512;
    q.group_cols = ((m128 / 128 + q.n_groups - 1) / q.n_groups) * 128;
    q.tmem_cols = tmem_cols_for(q.group_cols);
    const uint32_t stage_w2 = (2 + q.group_cols / 128) * kBlockK * 128;
    q.stages = pick_stages(stage_w2, 232448 - 2048, 8);
    MEGB_CHECK_ARG(q.stages >= 2, "tc_wgrad(pair): tile too large");
    const long long base2 = (long long)q.pair_tiles * q.n_groups;
    const long long slot_elems2 = (long long)taps * M * C;
    int splits2 = (int)((num_sms() / 2) / base2);              // ONE wave of clusters
    if (options().wgrad_splits > 0) splits2 = options().wgrad_splits;
    if (splits2 < 1) splits2 = 1;
    if (splits2 > B) splits2 = B;
    const long long max_slots2 = (long long)(ws_bytes / sizeof(float)) / slot_elems2;
    if (splits2 > max_slots2) splits2 = (int)max_slots2;
    MEGB_CHECK_ARG(splits2 >= 1, "tc_wgrad: workspace too small (%zu bytes, need >= %lld)", ws_bytes, slot_elems2 * 4);
    q.samples_per_cta = (B + splits2 - 1) / splits2;
    splits2 = (B + q.samples_per_cta - 1) / q.samples_per_cta;
    q.out = workspace;
    CUtensorMap tmU2, tmX2;
    {
      const uint64_t dims[3] = {(uint64_t)M, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Ms * 2, (uint64_t)T * Ms * 2};
      const uint32_t box[3] = {64, kBlockK, 1};
      int rc = encode_tmap_bf16(&tmU2, dU, 3, dims, strides, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)Cs * 2, (uint64_t)T * Cs * 2};
      const uint32_t box[3] = {64, kBlockK, 1};
      int rc = encode_tmap_bf16(&tmX2, Xin, 3, dims, strides, box);
      if (rc) return rc;
    }
    const size_t smem2 = 1024 + kHeaderBytes + (size_t)q.stages * stage_w2;
    int rc = set_smem(tc_wgrad2_kernel, smem2);
    if (rc) return rc;
    // the partial slots of padded (tap, channel) rows / channels are never written: the reduce only reads [taps][M][C]
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * base2 * splits2));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MEGB_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_wgrad2_kernel, tmU2, tmX2, q));
    MEGB_LAUNCH_OK();
    options().last_wgrad_splits = splits2;
    options().last_wgrad_kernel = 1;
    long long blocks = ((long long)M * C + 127) / 128;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    wgrad_reduce_kernel<<<(int)blocks, 128, 0, st>>>(workspace, dW, splits2, taps, M, C, 1);
    MEGB_LAUNCH_OK();
    return 0;
  }
  options().last_wgrad_kernel = 0;
  WgradParams p;
  p.T = T; p.M = M; p.C = C; p.taps = taps; p.dil = dil; p.B = B;
  p.kchunks = (T + kBlockK - 1) / kBlockK;
  const int c64 = (C + 63) / 64 * 64;
  // c tile <= 320 columns: issued as (192 + 128) / 256 / 192 / 128 / 64
  p.c_tiles = (c64 + 319) / 320;
  p.ctile = ((c64 / 64 + p.c_tiles - 1) / p.c_tiles) * 64;
  if (p.ctile > 256) { p.n0_cols = 192; p.n1_cols = p.ctile - 192; }
  else { p.n0_cols = p.ctile; p.n1_cols = 0; }
  p.idesc0 = make_idesc(p.n0_cols, 1, 1);
  p.idesc1 = p.n1_cols ? make_idesc(p.n1_cols, 1, 1) : 0;
  p.m_tiles = (M + kTileM - 1) / kTileM;
  p.tmem_cols = tmem_cols_for(p.ctile);
  const uint32_t stage_bytes = 2 * kBlockK * 128 + (p.ctile / 64) * kBlockK * 128;
  p.stages = pick_stages(stage_bytes, 232448 - 2048, 6);       // all of the 227 KB a block may opt into (one CTA per SM)
  MEGB_CHECK_ARG(p.stages >= 2, "tc_wgrad: tile too large");
  p.group = group; p.order = order;
  const long long base = (long long)p.m_tiles * p.c_tiles * taps;
  const long long slot_elems = (long long)taps * M * C;
  int splits;
  if (group == nullptr) {
    p.mode = 0;
    // ONE wave: floor, not ceil.  (ceil gave 15 x 10 = 150 CTAs for the 320->640 layers and 16 x 10 = 160 for the 640->1024
    // head on 148 SMs at one CTA per SM: a second wave of 2 / 12 CTAs doubled those launches -- seen in the ncu grid sizes.)
    splits = (int)(num_sms() / base);
    if (options().wgrad_splits > 0) splits = options().wgrad_splits;
    if (splits < 1) splits = 1;
    if (splits > B) splits = B;
    const long long max_slots = (long long)(ws_bytes / sizeof(float)) / slot_elems;
    if (splits > max_slots) splits = (int)max_slots;
    MEGB_CHECK_ARG(splits >= 1, "tc_wgrad: workspace too small (%zu bytes, need >= %lld)", ws_bytes, slot_elems * 4);
    p.out = workspace;
  } else {
    MEGB_CHECK_ARG(order != nullptr, "tc_wgrad: grouped mode needs the group-sorted sample order");
    p.mode = 1;
    splits = (int)(num_sms() / base);
    if (splits > B) splits = B;
    if (splits < 1) splits = 1;
    p.out = dW;
  }
  p.samples_per_cta = (B + splits - 1) / splits;
  splits = (B + p.samples_per_cta - 1) / p.samples_per_cta;

  CUtensorMap tmU, tmX;
  {
    const uint64_t dims[3] = {(uint64_t)M, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)Ms * 2, (uint64_t)T * Ms * 2};
    const uint32_t box[3] = {64, kBlockK, 1};
    int rc = encode_tmap_bf16(&tmU, dU, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)Cs * 2, (uint64_t)T * Cs * 2};
    const uint32_t box[3] = {64, kBlockK, 1};
    int rc = encode_tmap_bf16(&tmX, Xin, 3, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * stage_bytes;
  int rc = set_smem(tc_wgrad_kernel, smem);
  if (rc) return rc;
  dim3 grid(p.m_tiles * p.c_tiles, taps, splits);
  tc_wgrad_kernel<<<grid, kThreads, smem, st>>>(tmU, tmX, p);
  MEGB_LAUNCH_OK();
  options().last_wgrad_splits = splits;
  if (p.mode == 0) {
    long long blocks = ((long long)M * C + 127) / 128;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    wgrad_reduce_kernel<<<(int)blocks, 128, 0, st>>>(workspace, dW, splits, taps, M, C, 0);
    MEGB_LAUNCH_OK();
  }
  return 0;
}

int tc_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, cudaStream_t st) {
  MEGB_CHECK_ARG(D % 8 == 0, "tc_clip_logits: D must be a multiple of 8");
  LogitsParams p;
  p.By = By; p.Bz = Bz;
  const int bz16 = (Bz + 15) / 16 * 16;
  p.NT = bz16 > 256 ? 256 : bz16;
  const int j_tiles = (Bz + p.NT - 1) / p.NT;
  p.i_tiles = (By + kTileM - 1) / kTileM;
  p.kchunks_total = (int)((D + kBlockK - 1) / kBlockK);
  const long long tiles = (long long)p.i_tiles * j_tiles;
  // split-K so that the grid is ONE wave: floor, not ceil -- 16 i-tiles (8 ranks) x ceil(148/16) = 160 CTAs ran as two
  // waves, the second 8 % full
  long long splits = (long long)num_sms() / tiles;
  if (splits > p.kchunks_total) splits = p.kchunks_total;
  if (splits < 1) splits = 1;
  p.kchunks_per_cta = (int)((p.kchunks_total + splits - 1) / splits);
  splits = (p.kchunks_total + p.kchunks_per_cta - 1) / p.kchunks_per_cta;
  MEGB_CHECK_ARG(splits <= 65535, "tc_clip_logits: too many splits");
  const uint32_t stage_bytes = kTileM * 128 + p.NT * 128;
  p.stages = pick_stages(stage_bytes, 200 * 1024, 6);
  p.tmem_cols = tmem_cols_for(p.NT);
  p.idesc = make_idesc(p.NT, 0, 0);
  p.raw = raw;
  CUtensorMap tmY, tmZ;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)By};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {kBlockK, kTileM};
    int rc = encode_tmap_bf16(&tmY, Yp, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)Bz};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)p.NT};
    int rc = encode_tmap_bf16(&tmZ, Zp, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * stage_bytes;
  int rc = set_smem(tc_clip_logits_kernel, smem);
  if (rc) return rc;
  dim3 grid((unsigned)tiles, (unsigned)splits);
  tc_clip_logits_kernel<<<grid, kThreads, smem, st>>>(tmY, tmZ, p);
  MEGB_LAUNCH_OK();
  return 0;
}

int tc_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
               void* dZ, int By, int Bz, long long D, cudaStream_t st) {
  MEGB_CHECK_ARG(D % 16 == 0 && ldc % 8 == 0, "tc_clip_dz: D %% 16 and ldc %% 8 required (D=%lld ldc=%d)", D, ldc);
  DzParams p;
  p.By = By; p.Bz = Bz; p.D = D;
  const bool no_small = options().clip_dz_large != 0;
  if (By <= 512 && D >= 128 && D % 64 == 0 && (D + 127) / 128 <= 65535 && !no_small) {
    // stream-bound regime: TMA-staged Z / dZ tiles, 128 x 128 tiles, two CTAs per SM
    p.NT = 128;
    p.swap_grid = 1;
    p.kchunks = (By + kBlockK - 1) / kBlockK;
    p.stages = 2;
    p.tmem_cols = 128;
    p.idesc = make_idesc(128, 0, 1);
    p.Z = (const __nv_bfloat16*)Z; p.cvec = cvec; p.gscale = gscale; p.dZ = (__nv_bfloat16*)dZ;
    CUtensorMap tmC, tmY, tmZ, tmDZ;
    {
      const uint64_t dims[2] = {(uint64_t)By, (uint64_t)Bz};
      const uint64_t strides[1] = {(uint64_t)ldc * 2};
      const uint32_t box[2] = {kBlockK, kTileM};
      int rc = encode_tmap_bf16(&tmC, coef, 2, dims, strides, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[2] = {(uint64_t)D, (uint64_t)By};
      const uint64_t strides[1] = {(uint64_t)D * 2};
      const uint32_t box[2] = {64, kBlockK};
      int rc = encode_tmap_bf16(&tmY, Yp, 2, dims, strides, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[2] = {(uint64_t)D, (uint64_t)Bz};
      const uint64_t strides[1] = {(uint64_t)D * 2};
      const uint32_t box[2] = {64, kTileM};
      int rc = encode_tmap_bf16(&tmZ, Z, 2, dims, strides, box);
      if (rc) return rc;
      rc = encode_tmap_bf16(&tmDZ, dZ, 2, dims, strides, box);
      if (rc) return rc;
    }
    const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * (kTileM * 128 + 2 * kBlockK * 128) + 2 * kTileM * 128;
    int rc = set_smem(tc_clip_dz_small_kernel, smem);
    if (rc) return rc;
    dim3 grid((Bz + kTileM - 1) / kTileM, (unsigned)((D + 127) / 128));
    tc_clip_dz_small_kernel<<<grid, kThreads, smem, st>>>(tmC, tmY, tmZ, tmDZ, p);
    MEGB_LAUNCH_OK();
    options().last_clip_dz_kernel = 0;
    return 0;
  }
  p.NT = D >= 256 ? 256 : (int)((D + 63) / 64 * 64);
  p.kchunks = (By + kBlockK - 1) / kBlockK;
  const long long dt = (D + p.NT - 1) / p.NT;
  const int j_tiles = (Bz + kTileM - 1) / kTileM;
  // j tiles per CTA: two share one staged Y tile when that still leaves at least two waves of CTAs (speech shapes: D / 256
  // = 1440 d tiles); short embeddings (GOD, D = 512) keep one j tile per CTA for parallelism
  p.jt = 1;
  if (options().clip_dz_jt > 0) p.jt = options().clip_dz_jt;
  else if (j_tiles >= 2 && 2 * p.NT <= 512 && dt * ((j_tiles + 1) / 2) >= 2ll * num_sms()) p.jt = 2;
  MEGB_CHECK_ARG(p.jt >= 1 && p.jt * p.NT <= 512, "tc_clip_dz: %d j tiles of %d columns do not fit TMEM", p.jt, p.NT);
  const uint32_t stage_bytes = p.jt * kTileM * 128 + (p.NT / 64) * kBlockK * 128;
  p.stages = pick_stages(stage_bytes, p.jt > 1 ? 200 * 1024 : 100 * 1024, 4);
  MEGB_CHECK_ARG(p.stages >= 2, "tc_clip_dz: tile too large");
  p.tmem_cols = tmem_cols_for(p.jt * p.NT);
  p.idesc = make_idesc(p.NT, 0, 1);
  p.Z = (const __nv_bfloat16*)Z; p.cvec = cvec; p.gscale = gscale; p.dZ = (__nv_bfloat16*)dZ;
  CUtensorMap tmC, tmY;
  {
    const uint64_t dims[2] = {(uint64_t)By, (uint64_t)Bz};
    const uint64_t strides[1] = {(uint64_t)ldc * 2};
    const uint32_t box[2] = {kBlockK, kTileM};
    int rc = encode_tmap_bf16(&tmC, coef, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)By};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {64, kBlockK};
    int rc = encode_tmap_bf16(&tmY, Yp, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * stage_bytes;
  int rc = set_smem(tc_clip_dz_kernel, smem);
  if (rc) return rc;
  MEGB_CHECK_ARG(dt < 2147483647ll, "tc_clip_dz: grid too large");
  const int j_groups = (j_tiles + p.jt - 1) / p.jt;
  p.swap_grid = dt <= 65535 ? 1 : 0;
  dim3 grid((unsigned)dt, j_groups);
  if (p.swap_grid) grid = dim3(j_groups, (unsigned)dt);
  tc_clip_dz_kernel<<<grid, kThreads, smem, st>>>(tmC, tmY, p);
  MEGB_LAUNCH_OK();
  options().last_clip_dz_kernel = 1;
  return 0;
}

int tc_clip_fused(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp, int By, int Bz,
                  long long D, int off, int mode, float* row_part, float* col_part, float* diag, const float* row_lse,
                  const float* col_lse, float wscale, int j_begin, int j_count, void* coef, int ldc, float* cvec,
                  float* out2, cudaStream_t st) {
  MEGB_CHECK_ARG(D % 8 == 0 && D <= 65536, "tc_clip_fused: D must be a multiple of 8 (and short: %lld)", D);
  MEGB_CHECK_ARG(j_begin >= 0 && j_count > 0 && j_begin + j_count <= Bz, "tc_clip_fused: column range outside [0, Bz)");
  MEGB_CHECK_ARG(off >= 0 && off + Bz <= By, "tc_clip_fused: positives [off, off+Bz) outside the %d speech rows", By);
  ClipFusedParams p;
  p.By = By; p.Bz = Bz; p.mode = mode; p.off = off; p.ldc = ldc; p.j_begin = j_begin; p.j_end = j_begin + j_count;
  p.wscale = wscale;
  p.kchunks = (int)((D + kBlockK - 1) / kBlockK);
  p.i_tiles = (By + kTileM - 1) / kTileM;
  const int j_tiles = (j_count + kFusedNT - 1) / kFusedNT;
  p.stages = p.kchunks < 3 ? (p.kchunks < 2 ? 2 : p.kchunks) : 3;       // 96 KB: two CTAs per SM
  p.idesc = make_idesc(kFusedNT, 0, 0);
  p.ysq = ysq; p.zsq = zsq; p.temp = temp; p.row_part = row_part; p.col_part = col_part; p.diag = diag;
  p.row_lse = row_lse; p.col_lse = col_lse; p.coef = (__nv_bfloat16*)coef; p.cvec = cvec; p.out2 = out2;
  CUtensorMap tmY, tmZ;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)By};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {kBlockK, kTileM};
    int rc = encode_tmap_bf16(&tmY, Yp, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)Bz};
    const uint64_t strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)kFusedNT};
    int rc = encode_tmap_bf16(&tmZ, Zp, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + kHeaderBytes + (size_t)p.stages * (kTileM * 128 + kFusedNT * 128) + 10 * kFusedNT * sizeof(float);
  int rc = set_smem(tc_clip_fused_kernel, smem);
  if (rc) return rc;
  const long long blocks = (long long)p.i_tiles * j_tiles;
  MEGB_CHECK_ARG(blocks < 2147483647ll, "tc_clip_fused: grid too large");
  tc_clip_fused_kernel<<<(unsigned)blocks, kThreads, smem, st>>>(tmY, tmZ, p);
  MEGB_LAUNCH_OK();
  return 0;
}

int tc_clip_fused_loss(const float* row_lse, const float* col_lse, const float* diag, int off, int Bz, float wscale,
                       float* out2, cudaStream_t st) {
  int blocks = (Bz + 255) / 256;
  if (blocks > 64) blocks = 64;
  clip_fused_loss_kernel<<<blocks, 256, 0, st>>>(row_lse, col_lse, diag, off, Bz, wscale, out2);
  MEGB_LAUNCH_OK();
  return 0;
}

}  // namespace tc
}  // namespace megb200
