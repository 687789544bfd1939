// HBM-bound passes of the hot path: layout packing, BatchNorm statistics/apply, GELU, GLU, and their backward
// forms.  All work on the PC layout [rows][Cs] with 8-channel (16 B bf16 / 32 B f32) vector accesses so that a
// warp touches whole 128 B lines; per-channel reductions go registers -> shared atomics -> one global atomic per
// channel per CTA.  Bound: HBM bandwidth; algorithmic bytes are stated per op in DESIGN.md.
#include "common.cuh"

#include <math.h>

namespace megb200 {

static inline int ew_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------ transpose-pack
// in f32 [B][C][T] -> out [B][T][Cs] (+ optional per-sample sum of squares of the unrounded input).
// 64 (channels) x 64 (positions) tile, 256 threads: a warp reads one channel row as 32 x 8 B = 256 contiguous bytes and
// writes one position row as 32 x (2 channels) = 128 B of bf16 (256 B of f32), so both sides move whole 128 B lines.
// (The first form used 32 x 32 tiles and 2-byte stores: half-line writes, 3.4 TB/s on the 566 MB target pack.)
constexpr int kTpTile = 64;

template <typename Tin, typename T>
__global__ void __launch_bounds__(256) transpose_pack_kernel(const Tin* __restrict__ in, T* __restrict__ out,
                                                             float* __restrict__ sq, int B, int C, int Tn, int Cs) {
  __shared__ float tile[kTpTile][kTpTile + 1];
  __shared__ float red[8];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * kTpTile, c0 = blockIdx.y * kTpTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // two-element loads: even rows, base aligned to two elements
  const bool t_vec = (Tn & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & (2 * sizeof(Tin) - 1)) == 0;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cl = warp + 8 * i, c = c0 + cl, t = t0 + 2 * lane;
    float v0 = 0.f, v1 = 0.f;
    if (c < C) {
      const Tin* src = in + ((long long)b * C + c) * Tn + t;
      if (t_vec && t + 1 < Tn) {
        if constexpr (sizeof(Tin) == 2) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src));
          v0 = v.x;
          v1 = v.y;
        } else {
          const float2 v = *reinterpret_cast<const float2*>(src);
          v0 = v.x;
          v1 = v.y;
        }
      } else {
        if (t < Tn) v0 = to_f(src[0]);
        if (t + 1 < Tn) v1 = to_f(src[1]);
      }
    }
    tile[cl][2 * lane] = v0;
    tile[cl][2 * lane + 1] = v1;
    ss = fmaf(v0, v0, fmaf(v1, v1, ss));
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int tl = warp + 8 * i, t = t0 + tl, c = c0 + 2 * lane;
    if (t < Tn && c < Cs) {          // Cs is even (multiple of 8), so c + 1 < Cs too; channels >= C were read as 0
      T* dst = out + ((long long)b * Tn + t) * Cs + c;
      const float a0 = tile[2 * lane][tl], a1 = tile[2 * lane + 1][tl];
      if constexpr (sizeof(T) == 2) {
        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a0, a1);
      } else {
        *reinterpret_cast<float2*>(dst) = make_float2(a0, a1);
      }
    }
  }
  if (sq != nullptr) {
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i];
      atomicAdd(&sq[b], s);
    }
  }
}

template <typename T, typename Tin = float>
static int launch_transpose_pack(const Tin* in, void* out, float* sq, int B, int C, int Tn, int Cs,
                                 cudaStream_t st) {
  dim3 grid((Tn + kTpTile - 1) / kTpTile, (Cs + kTpTile - 1) / kTpTile, B);
  transpose_pack_kernel<Tin, T><<<grid, 256, 0, st>>>(in, (T*)out, sq, B, C, Tn, Cs);
  MEGB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------ weight packing
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int G, int O, int I, int Kt,
                                   int Np, int Kp, int mode, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % ((long long)Kt * Kp));
    const long long r = idx / ((long long)Kt * Kp);
    const int n = (int)(r % Np);
    const int g = (int)(r / Np);
    const int j = kk / Kp, k = kk % Kp;
    float v = 0.f;
    if (mode == 0) {
      if (n < O && k < I) v = w[(((long long)g * O + n) * I + k) * Kt + j];
    } else {
      if (n < I && k < O) v = w[(((long long)g * O + k) * I + n) * Kt + (Kt - 1 - j)];
    }
    out[idx] = from_f<T>(v);
  }
}

// all conv weights of a step in ONE launch: entry e of `table` = {src, dst, G, O, I, Kt, Np, Kp, mode, total} (int64)
// One warp per output row (g, n, tap): the index arithmetic (integer divisions by run-time sizes) is paid once per row,
// the lanes then stream over k with coalesced stores.  (The first form divided 64-bit indices four times per element
// and was ALU-bound: 58 us for 10 M elements.)
template <typename T>
__global__ void pack_weight_multi_kernel(const long long* __restrict__ table) {
  const long long* e = table + 10 * blockIdx.y;
  const float* w = reinterpret_cast<const float*>(e[0]);
  T* out = reinterpret_cast<T*>(e[1]);
  const int G = (int)e[2], O = (int)e[3], I = (int)e[4], Kt = (int)e[5], Np = (int)e[6], Kp = (int)e[7], mode = (int)e[8];
  const unsigned rows = (unsigned)G * Np * Kt;
  const int lane = threadIdx.x & 31;
  const unsigned wpb = blockDim.x >> 5;
  for (unsigned r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int j = (int)(r % Kt);
    const unsigned gn = r / Kt;
    const int n = (int)(gn % Np), g = (int)(gn / Np);
    T* orow = out + ((size_t)gn * Kt + j) * Kp;
    if (mode == 0) {
      const float* src = w + ((size_t)g * O + n) * I * Kt + j;
      const bool row_ok = n < O;
#pragma unroll 8
      for (int k = lane; k < Kp; k += 32) orow[k] = from_f<T>(row_ok && k < I ? src[(size_t)k * Kt] : 0.f);
    } else {
      const float* src = w + ((size_t)g * O * I + n) * Kt + (Kt - 1 - j);
      const bool row_ok = n < I;
#pragma unroll 8
      for (int k = lane; k < Kp; k += 32) orow[k] = from_f<T>(row_ok && k < O ? src[(size_t)k * I * Kt] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------ row-loop passes
// One CTA owns a slab of rows; thread = (channel group of 8, row lane) and keeps its channel group for the whole
// slab, so per-channel parameters are loaded ONCE into registers (Op::init) and the loop body is pure streaming.
// Op::row() consumes one 8-vector per input stream, may write outputs, and may add into up to two per-channel
// accumulators that end in one global atomic per channel per CTA.
//
// An Op names its input streams (base pointer + row pitch in bytes; a row slab of a stream is one contiguous range of
// memory).  The staged kernel moves those slabs HBM -> shared memory with 1-D bulk copies (cp.async.bulk, one elected
// producer thread, mbarrier ring of `stages` slabs), so the depth of the memory pipeline is set by the ring, not by
// registers/occupancy, and the consumer warps spend their issue slots on math, LDS and the output stores only.
// Measured with ncu before this form (direct per-thread LDG, 24 warps/SM): 34-48 % of DRAM peak, top stall
// long_scoreboard.  The direct kernel remains for grouped reductions and very wide rows.
constexpr int kRowConsumers = 256;

template <class Op>
__device__ __forceinline__ void rowloop_flush(const Op&, int C, int Cs, const int32_t* group, int Tn, long long r0,
                                              float* sm0, float* sm1, float* out0, float* out1, int tid, int nthr) {
  long long obase = 0;
  if (group != nullptr) obase = (long long)group[r0 / Tn] * C;
  if (out0 != nullptr)
    for (int c = tid; c < C; c += nthr) atomicAdd(&out0[obase + c], sm0[c]);
  if (Op::kAcc > 1 && out1 != nullptr)
    for (int c = tid; c < C; c += nthr) atomicAdd(&out1[obase + c], sm1[c]);
}

template <class Op>
__global__ void __launch_bounds__(256, 3) rowloop_kernel(Op op, long long rows, long long rows_per_block, int C, int Cs,
                                                      const int32_t* __restrict__ group, int Tn,
                                                      float* __restrict__ out0, float* __restrict__ out1) {
  extern __shared__ float sm[];  // [2][Cs] when the op reduces
  float* sm0 = sm;
  float* sm1 = sm + Cs;
  if (Op::kAcc > 0) {
    for (int i = threadIdx.x; i < 2 * Cs; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  const int cg = Cs >> 3;
  const int W = cg < 256 ? cg : 256;
  const int R = 256 / W;
  const int lane_g = threadIdx.x % W, lane_r = threadIdx.x / W;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const uint8_t* b0 = static_cast<const uint8_t*>(op.in_base(0));
  const uint8_t* b1 = static_cast<const uint8_t*>(op.in_base(Op::kIn - 1));
  const long long p0 = op.in_pitch(0), p1 = op.in_pitch(Op::kIn - 1);
  if (lane_r < R) {
    for (int g = lane_g; g < cg; g += W) {
      float a0[8], a1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
      op.init(g * 8);
      if (Op::kPublish && blockIdx.x == 0 && lane_r == 0) op.publish(g * 8);
#pragma unroll 2
      for (long long r = r0 + lane_r; r < r1; r += R) op.row(b0 + r * p0, b1 + r * p1, r, g * 8, a0, a1);
      op.finish(g * 8, a0, a1);
      if (Op::kAcc > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(&sm0[g * 8 + i], a0[i]);
          if (Op::kAcc > 1) atomicAdd(&sm1[g * 8 + i], a1[i]);
        }
      }
    }
  }
  if (Op::kAcc > 0) {
    __syncthreads();
    rowloop_flush(op, C, Cs, group, Tn, r0, sm0, sm1, out0, out1, threadIdx.x, blockDim.x);
  }
}

// ---- staged form
__device__ __forceinline__ uint32_t rs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded: a lost copy must surface as a launch failure, never as a hung GPU
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void rs_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

constexpr int kRowMaxStages = 8;

template <class Op>
__global__ void __launch_bounds__(kRowConsumers + 32, 2)
rowstage_kernel(Op op, long long rows, long long rows_per_block, int slab_rows, int stages, int C, int Cs,
                float* __restrict__ out0, float* __restrict__ out1) {
  extern __shared__ __align__(128) uint8_t rs_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(rs_smem);
  uint64_t* empty = full + kRowMaxStages;
  uint8_t* ring = rs_smem + 128;
  const uint32_t p0 = (uint32_t)op.in_pitch(0), p1 = Op::kIn > 1 ? (uint32_t)op.in_pitch(1) : 0u;
  const uint32_t stage_bytes = (uint32_t)slab_rows * (p0 + p1);
  float* sm0 = reinterpret_cast<float*>(ring + (size_t)stages * stage_bytes);
  float* sm1 = sm0 + Cs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      rs_mbar_init(rs_smem_u32(&full[s]), 1);
      rs_mbar_init(rs_smem_u32(&empty[s]), kRowConsumers / 32);     // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (Op::kAcc > 0)
    for (int i = threadIdx.x; i < 2 * Cs; i += blockDim.x) sm0[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const int nslabs = r1 > r0 ? (int)((r1 - r0 + slab_rows - 1) / slab_rows) : 0;

  if (warp == kRowConsumers / 32) {
    if (lane == 0) {
      const uint8_t* b0 = static_cast<const uint8_t*>(op.in_base(0));
      const uint8_t* b1 = static_cast<const uint8_t*>(op.in_base(Op::kIn - 1));
      for (int i = 0; i < nslabs; ++i) {
        const int s = i % stages;
        rs_mbar_wait(rs_smem_u32(&empty[s]), ((i / stages) & 1) ^ 1);
        const long long rs = r0 + (long long)i * slab_rows;
        const uint32_t n = (uint32_t)((r1 - rs) < slab_rows ? (r1 - rs) : slab_rows);
        const uint32_t fb = rs_smem_u32(&full[s]);
        const uint32_t dst = rs_smem_u32(ring + (size_t)s * stage_bytes);
        rs_mbar_expect_tx(fb, n * (p0 + p1));
        rs_bulk_load(dst, b0 + rs * p0, n * p0, fb);
        if (Op::kIn > 1) rs_bulk_load(dst + (uint32_t)slab_rows * p0, b1 + rs * p1, n * p1, fb);
      }
    }
  } else {
    const int cg = Cs >> 3;                          // <= 256 on this path (launcher)
    const int R = kRowConsumers / cg;
    const int lane_g = threadIdx.x % cg, lane_r = threadIdx.x / cg;
    const bool active = lane_r < R;
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    op.init(lane_g * 8);
    if (Op::kPublish && blockIdx.x == 0 && lane_r == 0) op.publish(lane_g * 8);
    for (int i = 0; i < nslabs; ++i) {
      const int s = i % stages;
      rs_mbar_wait(rs_smem_u32(&full[s]), (i / stages) & 1);
      const long long rs = r0 + (long long)i * slab_rows;
      const int n = (int)((r1 - rs) < slab_rows ? (r1 - rs) : slab_rows);
      const uint8_t* s0 = ring + (size_t)s * stage_bytes;
      const uint8_t* s1 = Op::kIn > 1 ? s0 + (size_t)slab_rows * p0 : s0;
      if (active) {
        // two rows in flight per thread; three resident CTAs per SM without unrolling measured 3-7 % slower
#pragma unroll 2
        for (int rr = lane_r; rr < n; rr += R) op.row(s0 + (size_t)rr * p0, s1 + (size_t)rr * p1, rs + rr, lane_g * 8, a0, a1);
      }
      __syncwarp();
      if (lane == 0) rs_mbar_arrive(rs_smem_u32(&empty[s]));
    }
    if (active) {
      op.finish(lane_g * 8, a0, a1);
      if (Op::kAcc > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(&sm0[lane_g * 8 + i], a0[i]);
          if (Op::kAcc > 1) atomicAdd(&sm1[lane_g * 8 + i], a1[i]);
        }
      }
    }
  }
  if (Op::kAcc > 0) {
    __syncthreads();
    rowloop_flush(op, C, Cs, nullptr, 1, r0, sm0, sm1, out0, out1, threadIdx.x, blockDim.x);
  }
}

template <typename T>
struct StatsOp {
  static constexpr int kAcc = 2, kIn = 1;
  static constexpr bool kPublish = false;
  const T* u;
  int Cs;
  __host__ __device__ __forceinline__ const void* in_base(int) const { return u; }
  __host__ __device__ __forceinline__ long long in_pitch(int) const { return (long long)Cs * sizeof(T); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pu, const void*, long long, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float x[8];
    load8(static_cast<const T*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a0[i] += x[i];
      a1[i] += x[i] * x[i];
    }
  }
};

template <typename T>
struct ColsumOp {
  static constexpr int kAcc = 1, kIn = 1;
  static constexpr bool kPublish = false;
  const T* u;
  int Cs;
  __host__ __device__ __forceinline__ const void* in_base(int) const { return u; }
  __host__ __device__ __forceinline__ long long in_pitch(int) const { return (long long)Cs * sizeof(T); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pu, const void*, long long, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float x[8];
    load8(static_cast<const T*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] += x[i];
  }
};

// H = gelu(scale*U + shift)
template <typename TU, typename T>
struct BnGeluOp {
  static constexpr int kAcc = 0, kIn = 1;
  static constexpr bool kPublish = false;
  const TU* u;
  T* h;
  const float *scale, *shift;
  int Cs, C;
  float sc[8], sh[8];
  __host__ __device__ __forceinline__ const void* in_base(int) const { return u; }
  __host__ __device__ __forceinline__ long long in_pitch(int) const { return (long long)Cs * sizeof(TU); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool in = c0 + i < C;
      sc[i] = in ? scale[c0 + i] : 0.f;
      sh[i] = in ? shift[c0 + i] : 0.f;
    }
  }
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pu, const void*, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float x[8], o[8];
    load8(static_cast<const TU*>(pu) + c0, x);
    // pad channels have scale = shift = 0 (init): gelu(0) = 0, so no per-element predicate is needed
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = gelu_t<T>(fmaf(x[i], sc[i], sh[i]));
    store8(h + r * Cs + c0, o);
  }
};

// dY = dH * gelu'(y), y = scale*u + shift; accumulates sum dY and sum dY*xhat.  Per element the loop only forms
// sum d and sum d*u; xhat = (u - mean)*rstd is folded in once per thread at the end (finish()).
template <typename TU, typename T>
struct BnGeluBwdOp {
  static constexpr int kAcc = 2, kIn = 2;
  static constexpr bool kPublish = false;
  const T* dh;
  const TU* u;
  T* dy;
  const float *scale, *shift, *mean, *rstd;
  int Cs, C;
  float sc[8], sh[8];
  __host__ __device__ __forceinline__ const void* in_base(int k) const { return k == 0 ? (const void*)dh : (const void*)u; }
  __host__ __device__ __forceinline__ long long in_pitch(int k) const { return (long long)Cs * (k == 0 ? sizeof(T) : sizeof(TU)); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool in = c0 + i < C;
      sc[i] = in ? scale[c0 + i] : 0.f;
      sh[i] = in ? shift[c0 + i] : 0.f;
    }
  }
  __device__ __forceinline__ void row(const void* pd, const void* pu, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float g[8], x[8], o[8];
    load8(static_cast<const T*>(pd) + c0, g);
    load8(static_cast<const TU*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = g[i] * gelu_grad_t<T>(fmaf(x[i], sc[i], sh[i]));
    }
    if (c0 + 8 > C) {                                 // only the last, partial channel group pays for the mask
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i >= C) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a0[i] += o[i];
      a1[i] = fmaf(o[i], x[i], a1[i]);
    }
    if (dy != nullptr) store8(dy + r * Cs + c0, o);
  }
  __device__ __forceinline__ void finish(int c0, float (&a0)[8], float (&a1)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      a1[i] = (c < C) ? (a1[i] - mean[c] * a0[i]) * rstd[c] : 0.f;       // sum d*(u - mean)*rstd
    }
  }
};

// dU = gamma*rstd*(dY - s0/n - xhat*s1/n) = A*dY + Bc*U + Cc per channel; accumulates sum dU (the conv bias gradient)
template <typename TU, typename T>
struct BnBwdApplyOp {
  static constexpr int kAcc = 1, kIn = 2;
  static constexpr bool kPublish = false;
  const T* dy;
  const TU* u;
  T* du;
  const float *gamma, *mean, *rstd, *sums2;
  float inv_n;
  int Cs, C;
  float ca[8], cb[8], cc[8];
  __host__ __device__ __forceinline__ const void* in_base(int k) const { return k == 0 ? (const void*)dy : (const void*)u; }
  __host__ __device__ __forceinline__ long long in_pitch(int k) const { return (long long)Cs * (k == 0 ? sizeof(T) : sizeof(TU)); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c < C) {
        const float rsd = rstd[c], A = gamma[c] * rsd;
        ca[i] = A;
        cb[i] = -A * rsd * sums2[C + c] * inv_n;
        cc[i] = -A * sums2[c] * inv_n - cb[i] * mean[c];
      } else {
        ca[i] = cb[i] = cc[i] = 0.f;
      }
    }
  }
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pd, const void* pu, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float g[8], x[8], o[8];
    load8(static_cast<const T*>(pd) + c0, g);
    load8(static_cast<const TU*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = round_to<T>(fmaf(ca[i], g[i], fmaf(cb[i], x[i], cc[i])));
      a0[i] += o[i];
    }
    store8(du + r * Cs + c0, o);
  }
};

// Training-mode BatchNorm + GELU in ONE launch: every thread derives scale/shift of its 8 channels from the batch sums
// (a handful of flops), block 0 also publishes scale/shift/mean/rstd for the backward pass and updates the running
// statistics (nn.BatchNorm1d: momentum update with the unbiased variance).  Replaces bn_finalize + bn_gelu: the tiny
// finalize launch cost ~8 us of dependent-launch bubble per BatchNorm.
template <typename TU, typename T>
struct BnGeluTrainOp {
  static constexpr int kAcc = 0, kIn = 1;
  static constexpr bool kPublish = true;
  const TU* u;
  T* h;
  const float *sums, *gamma, *beta;
  float *rm, *rv, *scale, *shift, *mean, *rstd;
  float count, eps, momentum;
  int Cs, C;
  float sc[8], sh[8];
  __host__ __device__ __forceinline__ const void* in_base(int) const { return u; }
  __host__ __device__ __forceinline__ long long in_pitch(int) const { return (long long)Cs * sizeof(TU); }
  __device__ __forceinline__ void stats(int c, float& m, float& var, float& rs) const {
    m = sums[c] / count;
    var = fmaxf(sums[C + c] / count - m * m, 0.f);
    rs = rsqrtf(var + eps);
  }
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      sc[i] = sh[i] = 0.f;
      if (c < C) {
        float m, var, rs;
        stats(c, m, var, rs);
        sc[i] = gamma[c] * rs;
        sh[i] = beta[c] - m * sc[i];
      }
    }
  }
  __device__ __forceinline__ void publish(int c0) const {
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c >= C) break;
      float m, var, rs;
      stats(c, m, var, rs);
      const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
      rm[c] = (1.f - momentum) * rm[c] + momentum * m;
      rv[c] = (1.f - momentum) * rv[c] + momentum * unbiased;
      scale[c] = sc[i];
      shift[c] = sh[i];
      mean[c] = m;
      rstd[c] = rs;
    }
  }
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pu, const void*, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float x[8], o[8];
    load8(static_cast<const TU*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = gelu_t<T>(fmaf(x[i], sc[i], sh[i]));
    store8(h + r * Cs + c0, o);
  }
};

// Second backward pass with the GELU derivative recomputed: dU = A*(dH*gelu'(scale*U+shift)) + Bc*U + Cc.  Reads dH
// and U (the same bytes the two-tensor form read) so the reduce pass no longer has to WRITE dY (59 MB per BatchNorm).
template <typename TU, typename T>
struct BnGeluBwdApplyOp {
  static constexpr int kAcc = 1, kIn = 2;
  static constexpr bool kPublish = false;
  const T* dh;
  const TU* u;
  T* du;
  const float *scale, *shift, *gamma, *mean, *rstd, *sums2;
  float inv_n;
  int Cs, C;
  float sc[8], sh[8], ca[8], cb[8], cc[8];
  __host__ __device__ __forceinline__ const void* in_base(int k) const { return k == 0 ? (const void*)dh : (const void*)u; }
  __host__ __device__ __forceinline__ long long in_pitch(int k) const { return (long long)Cs * (k == 0 ? sizeof(T) : sizeof(TU)); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c < C) {
        const float rsd = rstd[c], A = gamma[c] * rsd;
        sc[i] = scale[c];
        sh[i] = shift[c];
        ca[i] = A;
        cb[i] = -A * rsd * sums2[C + c] * inv_n;
        cc[i] = -A * sums2[c] * inv_n - cb[i] * mean[c];
      } else {
        sc[i] = sh[i] = ca[i] = cb[i] = cc[i] = 0.f;
      }
    }
  }
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pd, const void* pu, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float g[8], x[8], o[8];
    load8(static_cast<const T*>(pd) + c0, g);
    load8(static_cast<const TU*>(pu) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dy = g[i] * gelu_grad_t<T>(fmaf(x[i], sc[i], sh[i]));
      o[i] = round_to<T>(fmaf(ca[i], dy, fmaf(cb[i], x[i], cc[i])));     // pad channels: all coefficients 0
      a0[i] += o[i];
    }
    store8(du + r * Cs + c0, o);
  }
};

// dP = dQ * gelu'(P); accumulates sum dP (bias gradient of the 1x1 head convs)
template <typename T>
struct GeluBwdOp {
  static constexpr int kAcc = 1, kIn = 2;
  static constexpr bool kPublish = false;
  const T* dq;
  const T* p;
  T* dp;
  int Cs, C;
  __host__ __device__ __forceinline__ const void* in_base(int k) const { return k == 0 ? dq : p; }
  __host__ __device__ __forceinline__ long long in_pitch(int) const { return (long long)Cs * sizeof(T); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pd, const void* pp, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    float g[8], x[8], o[8];
    load8(static_cast<const T*>(pd) + c0, g);
    load8(static_cast<const T*>(pp) + c0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = g[i] * gelu_grad_t<T>(x[i]);
    }
    if (c0 + 8 > C) {                                 // only the last, partial channel group pays for the mask
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i >= C) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] += o[i];
    store8(dp + r * Cs + c0, o);
  }
};

// GLU backward over 8 output channels: dV[c] = d*s, dV[C+c] = d*a*s*(1-s); accumulates both halves' column sums.
// Requires C % 8 == 0 (the launcher falls back to the scalar kernel otherwise).
template <typename T>
struct GluBwdOp {
  static constexpr int kAcc = 2, kIn = 2;
  static constexpr bool kPublish = false;
  const T* dout;
  const T* V;
  T* dV;
  int Vs, Os, C;
  __host__ __device__ __forceinline__ const void* in_base(int k) const { return k == 0 ? dout : V; }
  __host__ __device__ __forceinline__ long long in_pitch(int k) const { return (long long)(k == 0 ? Os : Vs) * sizeof(T); }
  __device__ __forceinline__ void publish(int) const {}
  __device__ __forceinline__ void init(int) {}
  __device__ __forceinline__ void finish(int, float (&)[8], float (&)[8]) const {}
  __device__ __forceinline__ void row(const void* pd, const void* pv, long long r, int c0, float (&a0)[8],
                                      float (&a1)[8]) const {
    if (c0 >= C) return;
    float d[8], a[8], g[8], da[8], dg[8];
    load8(static_cast<const T*>(pd) + c0, d);
    load8(static_cast<const T*>(pv) + c0, a);
    load8(static_cast<const T*>(pv) + C + c0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = sigmoid_f(g[i]);
      da[i] = round_to<T>(d[i] * s);
      dg[i] = round_to<T>(d[i] * a[i] * s * (1.f - s));
      a0[i] += da[i];
      a1[i] += dg[i];
    }
    store8(dV + r * Vs + c0, da);
    store8(dV + r * Vs + C + c0, dg);
  }
};

static bool rowloop_direct_env() { return options().rowloop_direct != 0; }

template <class Op>
static int launch_rowloop(Op op, long long rows, int C, int Cs, const int32_t* group, int Tn, float* out0,
                          float* out1, cudaStream_t st) {
  MEGB_CHECK_ARG(Cs % 8 == 0 && Cs >= C && Cs <= 4096, "rowloop: bad channel stride %d (C=%d)", Cs, C);
  if (rows < 1) return 0;
  // ---- staged form: ungrouped, row fits the 256 consumer threads, inputs 16-byte aligned (bulk-copy rule)
  const long long pitch_sum = op.in_pitch(0) + (Op::kIn > 1 ? op.in_pitch(1) : 0);
  const bool aligned = (((uintptr_t)op.in_base(0) | (uintptr_t)op.in_base(Op::kIn - 1) | (uintptr_t)op.in_pitch(0) |
                         (uintptr_t)op.in_pitch(Op::kIn - 1)) & 15) == 0;
  if (group == nullptr && Cs <= 8 * kRowConsumers && aligned && pitch_sum <= 24 * 1024 && !rowloop_direct_env()) {
    const int R = kRowConsumers / (Cs >> 3);
    int slab_rows = (int)(24 * 1024 / pitch_sum) / R * R;          // ~24 KB per stage, whole row-lane rounds
    if (slab_rows < R) slab_rows = R;
    const long long stage_bytes = (long long)slab_rows * pitch_sum;
    const size_t acc_bytes = Op::kAcc > 0 ? 2 * Cs * sizeof(float) : 0;
    int stages = (int)((100 * 1024 - 128 - (long long)acc_bytes) / stage_bytes);      // two CTAs per SM
    if (stages > kRowMaxStages) stages = kRowMaxStages;
    if (stages >= 2) {
      long long blocks = (long long)num_sms() * 2;                  // one resident wave (2 CTAs per SM)
      long long rpb = (rows + blocks - 1) / blocks;
      if (rpb < slab_rows) rpb = slab_rows;
      blocks = (rows + rpb - 1) / rpb;
      const size_t smem = 128 + (size_t)stages * stage_bytes + acc_bytes;
      static bool attr_set = false;   // per instantiation
      if (!attr_set) {
        MEGB_CUDA_OK(cudaFuncSetAttribute(rowstage_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, 101 * 1024));
        attr_set = true;
      }
      rowstage_kernel<Op><<<(int)blocks, kRowConsumers + 32, smem, st>>>(op, rows, rpb, slab_rows, stages, C, Cs, out0, out1);
      MEGB_LAUNCH_OK();
      return 0;
    }
  }
  long long rpb;
  if (group != nullptr) {
    rpb = Tn;
  } else {
    long long target_blocks = (long long)num_sms() * 8;
    rpb = (rows + target_blocks - 1) / target_blocks;
    if (rpb < 16) rpb = 16;
  }
  long long blocks = (rows + rpb - 1) / rpb;
  const size_t smem = Op::kAcc > 0 ? 2 * Cs * sizeof(float) : 0;
  rowloop_kernel<Op><<<(int)blocks, 256, smem, st>>>(op, rows, rpb, C, Cs, group, Tn, out0, out1);
  MEGB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------ maps
template <typename T>
__global__ void glu_kernel(const T* __restrict__ V, T* __restrict__ out, long long rows, int C, int Vs, int Os) {
  const int og = Os >> 3;
  const long long nvec = rows * og;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / og;
    const int c0 = (int)(v % og) * 8;
    float o[8];
    if (c0 + 8 <= C && (C % 8) == 0) {
      float a[8], g[8];
      load8(V + r * Vs + c0, a);
      load8(V + r * Vs + C + c0, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = a[i] * sigmoid_f(g[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        o[i] = (c < C) ? to_f(V[r * Vs + c]) * sigmoid_f(to_f(V[r * Vs + C + c])) : 0.f;
      }
    }
    store8(out + r * Os + c0, o);
  }
}

template <typename T>
__global__ void glu_bwd_kernel(const T* __restrict__ dOut, const T* __restrict__ V, T* __restrict__ dV,
                               long long rows, int C, int Vs, int Os) {
  // one thread per 8 output channels; writes both halves of dV (and zeroes the pad of dV once)
  const int og = Os >> 3;
  const long long nvec = rows * og;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / og;
    const int c0 = (int)(v % og) * 8;
    if (c0 + 8 <= C && (C % 8) == 0) {
      float d[8], a[8], g[8], da[8], dg[8];
      load8(dOut + r * Os + c0, d);
      load8(V + r * Vs + c0, a);
      load8(V + r * Vs + C + c0, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s = sigmoid_f(g[i]);
        da[i] = d[i] * s;
        dg[i] = d[i] * a[i] * s * (1.f - s);
      }
      store8(dV + r * Vs + c0, da);
      store8(dV + r * Vs + C + c0, dg);
    } else {
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        if (c < C) {
          const float d = to_f(dOut[r * Os + c]);
          const float a = to_f(V[r * Vs + c]);
          const float s = sigmoid_f(to_f(V[r * Vs + C + c]));
          dV[r * Vs + c] = from_f<T>(d * s);
          dV[r * Vs + C + c] = from_f<T>(d * a * s * (1.f - s));
        }
      }
    }
    if (c0 == 0) {
      for (int c = 2 * C; c < Vs; ++c) dV[r * Vs + c] = from_f<T>(0.f);
    }
  }
}

template <typename T, bool kBwd>
__global__ void gelu_kernel(const T* __restrict__ dq, const T* __restrict__ p, T* __restrict__ o, long long nvec) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    float x[8], r[8];
    load8(p + v * 8, x);
    if (kBwd) {
      float g[8];
      load8(dq + v * 8, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = g[i] * gelu_grad_t<T>(x[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = gelu_t<T>(x[i]);
    }
    store8(o + v * 8, r);
  }
}

template <typename T>
__global__ void time_mean_kernel(const T* __restrict__ Z, T* __restrict__ Zm, int B, int Tn, int Fs) {
  const int fg = Fs >> 3;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < (long long)B * fg;
       v += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(v / fg), f0 = (int)(v % fg) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < Tn; ++t) {
      float x[8];
      load8(Z + ((long long)b * Tn + t) * Fs + f0, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x[i];
    }
    const float inv = 1.f / (float)Tn;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    store8(Zm + (long long)b * Fs + f0, acc);
  }
}

// AvgPool1d(k, stride) over time on the PC layout (BrainEncoderSeq2Static, models.py:483-485: kernel 3, stride 2, no
// padding): Out[b][to][c] = mean_{j<k} X[b][to*stride + j][c].  Thread = 8 channels of one output position.
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ X, T* __restrict__ Out, int B, int Tn, int To, int Cs, int k,
                               int stride) {
  const int cg = Cs >> 3;
  const long long nvec = (long long)B * To * cg;
  const float inv = 1.f / (float)k;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % cg) * 8;
    const long long bt = v / cg;
    const int to = (int)(bt % To), b = (int)(bt / To);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; ++j) {
      float x[8];
      load8(X + ((long long)b * Tn + to * stride + j) * Cs + c0, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    store8(Out + v * 8, acc);
  }
}

// dX[b][t][c] = (1/k) * sum over the windows that contain t of dOut[b][to][c]; positions past the last window get 0
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dOut, T* __restrict__ dX, int B, int Tn, int To, int Cs, int k,
                                   int stride) {
  const int cg = Cs >> 3;
  const long long nvec = (long long)B * Tn * cg;
  const float inv = 1.f / (float)k;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % cg) * 8;
    const long long bt = v / cg;
    const int t = (int)(bt % Tn), b = (int)(bt / Tn);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int lo = t - k + 1;
    lo = lo <= 0 ? 0 : (lo + stride - 1) / stride;           // first window with to*stride + k - 1 >= t
    for (int to = lo; to < To && to * stride <= t; ++to) {
      float x[8];
      load8(dOut + ((long long)b * To + to) * Cs + c0, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    store8(dX + v * 8, acc);
  }
}

template <typename T>
__global__ void time_mean_bwd_kernel(const T* __restrict__ dZm, T* __restrict__ dZ, int B, int Tn, int Fs) {
  const int fg = Fs >> 3;
  const long long nvec = (long long)B * Tn * fg;
  const float inv = 1.f / (float)Tn;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const int f0 = (int)(v % fg) * 8;
    const int b = (int)(v / ((long long)fg * Tn));
    float x[8];
    load8(dZm + (long long)b * Fs + f0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] *= inv;
    store8(dZ + v * 8, x);
  }
}

template <typename T>
__global__ void rowsumsq_kernel(const T* __restrict__ Z, float* __restrict__ sq, long long D, int chunks) {
  // grid = (chunks, B); D % 8 == 0
  __shared__ float red[8];
  const int b = blockIdx.y;
  const long long nvec = D >> 3;
  const long long per = (nvec + chunks - 1) / chunks;
  const long long v0 = blockIdx.x * per;
  long long v1 = v0 + per;
  if (v1 > nvec) v1 = nvec;
  float s = 0.f;
  for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    float x[8];
    load8(Z + b * D + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] * x[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(&sq[b], t);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<D>(to_f(s[i]));
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv,
                                   float count, float eps, float momentum, int training, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ rstd,
                                   int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m, var;
  if (training) {
    m = sums[c] / count;
    var = fmaxf(sums[C + c] / count - m * m, 0.f);
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    rm[c] = (1.f - momentum) * rm[c] + momentum * m;
    rv[c] = (1.f - momentum) * rv[c] + momentum * unbiased;
  } else {
    m = rm[c];
    var = rv[c];
  }
  const float rs = rsqrtf(var + eps);
  const float sc = gamma[c] * rs;
  scale[c] = sc;
  shift[c] = beta[c] - m * sc;
  mean[c] = m;
  rstd[c] = rs;
}

// softmax over sensors + dropout mask folded into the weights (rows = D1 outputs)
__global__ void sa_softmax_kernel(const float* __restrict__ a, const float* __restrict__ mask,
                                  float* __restrict__ W, float* __restrict__ Wm, int C) {
  __shared__ float red[32];
  const int o = blockIdx.x;
  const float* row = a + (long long)o * C;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += expf(row[i] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float w = expf(row[i] - mx) * inv;
    W[(long long)o * C + i] = w;
    Wm[(long long)o * C + i] = mask ? w * mask[i] : w;
  }
}

__global__ void sa_softmax_bwd_kernel(const float* __restrict__ W, const float* __restrict__ dWm,
                                      const float* __restrict__ mask, float* __restrict__ da, int C) {
  __shared__ float red[32];
  const int o = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float g = dWm[(long long)o * C + i] * (mask ? mask[i] : 1.f);
    s += W[(long long)o * C + i] * g;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float g = dWm[(long long)o * C + i] * (mask ? mask[i] : 1.f);
    da[(long long)o * C + i] = W[(long long)o * C + i] * (g - s);
  }
}

}  // namespace megb200

using namespace megb200;

#define DISPATCH_DTYPE(dtype, ...)                                   \
  if ((dtype) == MEGB200_F32) {                                      \
    using DT = float;                                                 \
    __VA_ARGS__                                                      \
  } else if ((dtype) == MEGB200_BF16) {                              \
    using DT = __nv_bfloat16;                                         \
    __VA_ARGS__                                                      \
  } else {                                                           \
    set_last_error("unknown dtype %d", (int)(dtype));                \
    return MEGB200_ERR_ARG;                                          \
  }

#define DISPATCH_U(u_dtype, ...)                                     \
  if ((u_dtype) == MEGB200_F32) {                                    \
    using TU = float;                                                \
    __VA_ARGS__                                                      \
  } else if ((u_dtype) == MEGB200_BF16) {                            \
    using TU = __nv_bfloat16;                                        \
    __VA_ARGS__                                                      \
  } else {                                                           \
    set_last_error("unknown u_dtype %d", (int)(u_dtype));            \
    return MEGB200_ERR_ARG;                                          \
  }

extern "C" {

int megb200_pack_x(const float* X, void* Xp, int B, int C, int T, int Cs, int dtype, void* stream) {
  MEGB_CHECK_ARG(X && Xp && B > 0 && C > 0 && T > 0 && Cs >= C && Cs % 8 == 0, "pack_x: bad arguments");
  MEGB_CHECK_ARG(B <= 65535, "pack_x: B too large");
  DISPATCH_DTYPE(dtype, return launch_transpose_pack<DT>(X, Xp, nullptr, B, C, T, Cs, (cudaStream_t)stream);)
}

int megb200_clip_prep_y(const void* Y, int y_dtype, void* Yp, float* ysq, int B, int F, int T, int dtype, void* stream) {
  MEGB_CHECK_ARG(Y && Yp && ysq && B > 0 && F > 0 && T > 0 && F % 8 == 0, "clip_prep_y: bad arguments (F %% 8)");
  MEGB_CHECK_ARG(B <= 65535, "clip_prep_y: B too large");
  if (y_dtype == MEGB200_BF16) {
    DISPATCH_DTYPE(dtype, return (launch_transpose_pack<DT, __nv_bfloat16>((const __nv_bfloat16*)Y, Yp, ysq, B, F, T, F,
                                                                          (cudaStream_t)stream));)
  }
  MEGB_CHECK_ARG(y_dtype == MEGB200_F32, "clip_prep_y: unknown target dtype %d", y_dtype);
  DISPATCH_DTYPE(dtype, return launch_transpose_pack<DT>((const float*)Y, Yp, ysq, B, F, T, F, (cudaStream_t)stream);)
}

int megb200_pack_weight(const float* w, void* out, int G, int O, int I, int Kt, int Np, int Kp, int mode, int dtype,
                        void* stream) {
  MEGB_CHECK_ARG(w && out && G > 0 && O > 0 && I > 0 && Kt > 0, "pack_weight: bad arguments");
  MEGB_CHECK_ARG(mode == 0 ? (Np >= O && Kp >= I) : (Np >= I && Kp >= O), "pack_weight: padded dims too small");
  const long long total = (long long)G * Np * Kt * Kp;
  DISPATCH_DTYPE(dtype, pack_weight_kernel<DT><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            w, (DT*)out, G, O, I, Kt, Np, Kp, mode, total);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_pack_weight_multi(const long long* table, int n, int dtype, void* stream) {
  MEGB_CHECK_ARG(table && n > 0 && n <= 65535, "pack_weight_multi: bad arguments");
  dim3 grid(192, n);     // 6 warps per block: one row per warp for the largest conv weights (<= 1152 rows)
  DISPATCH_DTYPE(dtype, pack_weight_multi_kernel<DT><<<grid, 192, 0, (cudaStream_t)stream>>>(table);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_colsum(const void* U, int B, int T, int C, int Cs, const int32_t* group, int G, float* out, int dtype,
                   void* stream) {
  MEGB_CHECK_ARG(U && out && B > 0 && T > 0 && G >= 1, "colsum: bad arguments");
  DISPATCH_DTYPE(dtype, ColsumOp<DT> op{(const DT*)U, Cs};
                 return launch_rowloop(op, (long long)B * T, C, Cs, group, T, out, nullptr, (cudaStream_t)stream);)
}

int megb200_bn_stats(const void* U, long long rows, int C, int Cs, float* sums, int u_dtype, void* stream) {
  MEGB_CHECK_ARG(U && sums && rows > 0, "bn_stats: bad arguments");
  DISPATCH_DTYPE(u_dtype, StatsOp<DT> op{(const DT*)U, Cs};
                 return launch_rowloop(op, rows, C, Cs, nullptr, 1, sums, sums + C, (cudaStream_t)stream);)
}

int megb200_bn_finalize(const float* sums, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, long long count, float eps, float momentum, int training, float* scale,
                        float* shift, float* mean, float* rstd, int C, void* stream) {
  MEGB_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && mean && rstd && C > 0,
                 "bn_finalize: null argument");
  MEGB_CHECK_ARG(!training || (sums && count > 0), "bn_finalize: training needs sums and count");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sums, gamma, beta, running_mean, running_var, (float)count, eps, momentum, training, scale, shift, mean, rstd, C);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_bn_gelu(const void* U, const float* scale, const float* shift, void* H, long long rows, int C, int Cs,
                    int u_dtype, int dtype, void* stream) {
  MEGB_CHECK_ARG(U && H && scale && shift && rows > 0 && Cs % 8 == 0 && Cs >= C, "bn_gelu: bad arguments");
  DISPATCH_U(u_dtype, DISPATCH_DTYPE(dtype, BnGeluOp<TU, DT> op{(const TU*)U, (DT*)H, scale, shift, Cs, C};
                                     return launch_rowloop(op, rows, C, Cs, nullptr, 1, nullptr, nullptr,
                                                           (cudaStream_t)stream);))
}

int megb200_bn_gelu_train(const void* U, const float* sums, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long count, float eps, float momentum, float* scale, float* shift,
                          float* mean, float* rstd, void* H, long long rows, int C, int Cs, int u_dtype, int dtype,
                          void* stream) {
  MEGB_CHECK_ARG(U && H && sums && gamma && beta && running_mean && running_var && scale && shift && mean && rstd,
                 "bn_gelu_train: null argument");
  MEGB_CHECK_ARG(rows > 0 && count > 0 && Cs % 8 == 0 && Cs >= C, "bn_gelu_train: bad arguments");
  DISPATCH_U(u_dtype, DISPATCH_DTYPE(dtype, BnGeluTrainOp<TU, DT> op{(const TU*)U, (DT*)H, sums, gamma, beta, running_mean,
                                                                    running_var, scale, shift, mean, rstd, (float)count,
                                                                    eps, momentum, Cs, C};
                                     return launch_rowloop(op, rows, C, Cs, nullptr, 1, nullptr, nullptr,
                                                           (cudaStream_t)stream);))
}

int megb200_bn_gelu_bwd_apply(const void* dH, const void* U, const float* scale, const float* shift, const float* gamma,
                              const float* mean, const float* rstd, const float* sums2, long long count, void* dU,
                              float* dbias, long long rows, int C, int Cs, int u_dtype, int dtype, void* stream) {
  MEGB_CHECK_ARG(dH && U && dU && scale && shift && gamma && mean && rstd && sums2, "bn_gelu_bwd_apply: null argument");
  MEGB_CHECK_ARG(rows > 0 && count > 0 && Cs % 8 == 0, "bn_gelu_bwd_apply: bad arguments");
  DISPATCH_U(u_dtype,
             DISPATCH_DTYPE(dtype, BnGeluBwdApplyOp<TU, DT> op{(const DT*)dH, (const TU*)U, (DT*)dU, scale, shift, gamma, mean,
                                                              rstd, sums2, 1.f / (float)count, Cs, C};
                            return launch_rowloop(op, rows, C, Cs, nullptr, 1, dbias, nullptr, (cudaStream_t)stream);))
}

int megb200_bn_gelu_bwd_reduce(const void* dH, const void* U, const float* scale, const float* shift,
                               const float* mean, const float* rstd, void* dY, float* sums2, long long rows, int C,
                               int Cs, int u_dtype, int dtype, void* stream) {
  MEGB_CHECK_ARG(dH && U && sums2 && rows > 0, "bn_gelu_bwd_reduce: bad arguments");   // dY may be NULL
  DISPATCH_U(u_dtype,
             DISPATCH_DTYPE(dtype, BnGeluBwdOp<TU, DT> op{(const DT*)dH, (const TU*)U, (DT*)dY, scale, shift, mean, rstd, Cs, C};
                            return launch_rowloop(op, rows, C, Cs, nullptr, 1, sums2, sums2 + C, (cudaStream_t)stream);))
}

int megb200_bn_bwd_apply(const void* dY, const void* U, const float* gamma, const float* mean, const float* rstd,
                         const float* sums2, long long count, void* dU, float* dbias, long long rows, int C, int Cs,
                         int u_dtype, int dtype, void* stream) {
  MEGB_CHECK_ARG(dY && U && dU && sums2 && rows > 0 && count > 0 && Cs % 8 == 0, "bn_bwd_apply: bad arguments");
  DISPATCH_U(u_dtype,
             DISPATCH_DTYPE(dtype, BnBwdApplyOp<TU, DT> op{(const DT*)dY, (const TU*)U, (DT*)dU, gamma, mean, rstd, sums2,
                                                          1.f / (float)count, Cs, C};
                            return launch_rowloop(op, rows, C, Cs, nullptr, 1, dbias, nullptr, (cudaStream_t)stream);))
}

int megb200_glu(const void* V, void* Out, long long rows, int C, int Vs, int Os, int dtype, void* stream) {
  MEGB_CHECK_ARG(V && Out && rows > 0 && Vs >= 2 * C && Os >= C && Vs % 8 == 0 && Os % 8 == 0, "glu: bad arguments");
  const long long nvec = rows * (Os / 8);
  DISPATCH_DTYPE(dtype, glu_kernel<DT><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const DT*)V, (DT*)Out,
                                                                                             rows, C, Vs, Os);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_glu_bwd(const void* dOut, const void* V, void* dV, float* dbias, long long rows, int C, int Vs, int Os,
                    int dtype, void* stream) {
  MEGB_CHECK_ARG(dOut && V && dV && rows > 0 && Vs >= 2 * C && Os >= C && Vs % 8 == 0 && Os % 8 == 0,
                 "glu_bwd: bad arguments");
  if (C % 8 == 0 && Vs == 2 * C) {
    // streaming row-loop form with the bias gradient (column sums of both halves) fused in
    DISPATCH_DTYPE(dtype, GluBwdOp<DT> op{(const DT*)dOut, (const DT*)V, (DT*)dV, Vs, Os, C};
                   return launch_rowloop(op, rows, C, Os, nullptr, 1, dbias, dbias ? dbias + C : nullptr,
                                         (cudaStream_t)stream);)
  }
  MEGB_CHECK_ARG(dbias == nullptr, "glu_bwd: the fused bias gradient needs C %% 8 == 0 and an unpadded V");
  const long long nvec = rows * (Os / 8);
  DISPATCH_DTYPE(dtype, glu_bwd_kernel<DT><<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const DT*)dOut, (const DT*)V, (DT*)dV, rows, C, Vs, Os);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_gelu(const void* P, void* Q, long long n, int dtype, void* stream) {
  MEGB_CHECK_ARG(P && Q && n > 0 && n % 8 == 0, "gelu: n must be a positive multiple of 8");
  DISPATCH_DTYPE(dtype, gelu_kernel<DT, false><<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                            nullptr, (const DT*)P, (DT*)Q, n / 8);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_gelu_bwd(const void* dQ, const void* P, void* dP, float* dbias, long long rows, int C, int Cs, int dtype,
                     void* stream) {
  MEGB_CHECK_ARG(dQ && P && dP && rows > 0 && Cs % 8 == 0 && Cs >= C, "gelu_bwd: bad arguments");
  DISPATCH_DTYPE(dtype, GeluBwdOp<DT> op{(const DT*)dQ, (const DT*)P, (DT*)dP, Cs, C};
                 return launch_rowloop(op, rows, C, Cs, nullptr, 1, dbias, nullptr, (cudaStream_t)stream);)
}

int megb200_time_mean(const void* Z, void* Zm, int B, int T, int F, int Fs, int dtype, void* stream) {
  MEGB_CHECK_ARG(Z && Zm && B > 0 && T > 0 && Fs % 8 == 0 && Fs >= F, "time_mean: bad arguments");
  DISPATCH_DTYPE(dtype, time_mean_kernel<DT><<<ew_grid((long long)B * Fs / 8, 128), 128, 0, (cudaStream_t)stream>>>(
                            (const DT*)Z, (DT*)Zm, B, T, Fs);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_time_mean_bwd(const void* dZm, void* dZ, int B, int T, int F, int Fs, int dtype, void* stream) {
  MEGB_CHECK_ARG(dZm && dZ && B > 0 && T > 0 && Fs % 8 == 0 && Fs >= F, "time_mean_bwd: bad arguments");
  DISPATCH_DTYPE(dtype,
                 time_mean_bwd_kernel<DT><<<ew_grid((long long)B * T * Fs / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                     (const DT*)dZm, (DT*)dZ, B, T, Fs);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_avgpool1d(const void* X, void* Out, int B, int T, int C, int Cs, int k, int stride, int dtype, void* stream) {
  MEGB_CHECK_ARG(X && Out && B > 0 && k > 0 && stride > 0 && T >= k && Cs % 8 == 0 && Cs >= C, "avgpool1d: bad arguments");
  const int To = (T - k) / stride + 1;
  DISPATCH_DTYPE(dtype, avgpool_kernel<DT><<<ew_grid((long long)B * To * Cs / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const DT*)X, (DT*)Out, B, T, To, Cs, k, stride);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_avgpool1d_bwd(const void* dOut, void* dX, int B, int T, int C, int Cs, int k, int stride, int dtype,
                          void* stream) {
  MEGB_CHECK_ARG(dOut && dX && B > 0 && k > 0 && stride > 0 && T >= k && Cs % 8 == 0 && Cs >= C, "avgpool1d_bwd: bad arguments");
  const int To = (T - k) / stride + 1;
  DISPATCH_DTYPE(dtype, avgpool_bwd_kernel<DT><<<ew_grid((long long)B * T * Cs / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const DT*)dOut, (DT*)dX, B, T, To, Cs, k, stride);
                 MEGB_LAUNCH_OK(); return 0;)
}

int megb200_rowsumsq(const void* Z, float* sq, int B, long long D, int dtype, void* stream) {
  MEGB_CHECK_ARG(Z && sq && B > 0 && D > 0 && D % 8 == 0 && B <= 65535, "rowsumsq: bad arguments (D %% 8)");
  int chunks = (int)((D / 8 + 1023) / 1024);
  const int cap = (num_sms() * 8 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, B);
  DISPATCH_DTYPE(dtype, rowsumsq_kernel<DT><<<grid, 256, 0, (cudaStream_t)stream>>>((const DT*)Z, sq, D, chunks);
                 MEGB_LAUNCH_OK(); return 0;)
}

// ------------------------------------------------------------------------------------------ segment gather ("next" row N3)
// Gwilliams2022DatasetBase.__getitem__ / segment_to_task (gwilliams2022.py:130-152) for a whole batch: recordings stay
// resident in HBM, a batch is B windows of (C, T) gathered by integer-exact index arithmetic.
//   task      = searchsorted(cumsum(num_segments_foreach_task), i + 1)   (left)
//   i_in_task = i - cumsum[task - 1]
//   onset     = onsets[onset_off[rec * n_tasks + task] + i_in_task]
//   out[b]    = recording(rec, task)[:, onset : onset + T]
// index_out[b] = {task, i_in_task, onset}; *error is set when a segment id or a window falls outside its table /
// recording (the reference would raise or return a short slice): such windows are zero-filled.
__global__ void segment_index_kernel(const long long* __restrict__ nseg_accum, int n_tasks,
                                     const long long* __restrict__ onset_off, const long long* __restrict__ onsets,
                                     const long long* __restrict__ rec_len, const long long* __restrict__ seg,
                                     const int* __restrict__ rec, int n_rec, int B, int T, long long* __restrict__ index_out,
                                     int* __restrict__ error) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long i = seg[b];
  const int r = rec[b];
  long long task = -1, iit = -1, onset = -1;
  if (i >= 0 && i < nseg_accum[n_tasks - 1] && r >= 0 && r < n_rec) {
    int t = 0;
    while (t < n_tasks && nseg_accum[t] < i + 1) ++t;             // np.searchsorted(accum, i + 1), side = "left"
    task = t;
    iit = i - (t > 0 ? nseg_accum[t - 1] : 0);
    const long long slot = (long long)r * n_tasks + t;
    onset = onsets[onset_off[slot] + iit];
    if (onset < 0 || onset + T > rec_len[slot]) {
      atomicOr(error, 2);
      onset = -1;
    }
  } else {
    atomicOr(error, 1);
  }
  index_out[3 * b + 0] = task;
  index_out[3 * b + 1] = iit;
  index_out[3 * b + 2] = onset;
}

__global__ void segment_copy_kernel(const long long* __restrict__ rec_base, const long long* __restrict__ rec_len,
                                    const int* __restrict__ rec, int n_tasks, const long long* __restrict__ index,
                                    int C, int T, float* __restrict__ out) {
  // blockIdx.x = (b, c) row; threads stride over the T samples of the window (contiguous in both source and output)
  const long long row = blockIdx.x;
  const int b = (int)(row / C), c = (int)(row % C);
  const long long task = index[3 * b], onset = index[3 * b + 2];
  float* o = out + row * T;
  if (task < 0 || onset < 0) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) o[t] = 0.f;
    return;
  }
  const long long slot = (long long)rec[b] * n_tasks + task;
  const float* src = reinterpret_cast<const float*>(rec_base[slot]) + (long long)c * rec_len[slot] + onset;
  for (int t = threadIdx.x; t < T; t += blockDim.x) o[t] = src[t];
}

int megb200_gather_segments(const long long* rec_base, const long long* rec_len, const long long* onset_off,
                            const long long* onsets, const long long* nseg_accum, int n_rec, int n_tasks,
                            const long long* seg, const int* rec, int B, int C, int T, float* out, long long* index_out,
                            int* error, void* stream) {
  MEGB_CHECK_ARG(rec_base && rec_len && onset_off && onsets && nseg_accum && seg && rec && out && index_out && error,
                 "gather_segments: null argument");
  MEGB_CHECK_ARG(n_rec > 0 && n_tasks > 0 && B > 0 && C > 0 && T > 0, "gather_segments: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  segment_index_kernel<<<(B + 127) / 128, 128, 0, st>>>(nseg_accum, n_tasks, onset_off, onsets, rec_len, seg, rec, n_rec, B, T,
                                                      index_out, error);
  MEGB_LAUNCH_OK();
  segment_copy_kernel<<<(unsigned)((long long)B * C), 128, 0, st>>>(rec_base, rec_len, rec, n_tasks, index_out, C, T, out);
  MEGB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------ collator ("next" row N2)
// Gwilliams2022Collator (gwilliams2022.py:653-662) on the device: per (sample, channel) row of T time steps
//   x <- x - mean(x[:L])                                  baseline_correction_single, preproc_utils.py:128-142
//   x <- (x - median) / (q75 - q25), IQR < 10 eps -> 1    sklearn RobustScaler via scaleAndClamp, preproc_utils.py:69-90
//   x <- clamp(x, -lim, lim)
// One CTA per row: the row is sorted in shared memory by a bitonic network over P = next power of two >= T slots (+inf
// padding), the three order statistics are read off with numpy's linear interpolation (virtual index q (T-1), lerp in
// double), and the row is rewritten from the unsorted copy.  Integer/compare work on 1.4 KB rows; bound by the
// barrier-separated network stages (45 for T = 360), several CTAs per SM.
__global__ void robust_scale_kernel(const float* __restrict__ X, float* __restrict__ Out, long long rows, int T, int P,
                                    int L, float lim, int clamp) {
  extern __shared__ float rs_buf[];      // [P] sort buffer, [T] unsorted baseline-corrected copy
  float* srt = rs_buf;
  float* raw = rs_buf + P;
  __shared__ float red[32];
  __shared__ float s_med;
  __shared__ double s_q[2];
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* x = X + row * T;
    float part = 0.f;
    for (int i = tid; i < T; i += nthr) {
      const float v = x[i];
      raw[i] = v;
      if (i < L) part += v;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < (nthr + 31) / 32; ++w) base += red[w];
    base /= (float)L;
    for (int i = tid; i < P; i += nthr) {
      float v = INFINITY;
      if (i < T) {
        v = raw[i] - base;
        raw[i] = v;
      }
      srt[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int c = tid; c < (P >> 1); c += nthr) {
          const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1));
          const int ip = i | j;
          const float a = srt[i], b = srt[ip];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            srt[i] = b;
            srt[ip] = a;
          }
        }
        __syncthreads();
      }
    }
    if (tid < 3) {
      if (tid == 1) {
        // np.nanmedian of an f32 column: middle element, or the f32 mean of the middle pair -> center_ stays f32
        s_med = (T & 1) ? srt[(T - 1) >> 1] : __fmul_rn(__fadd_rn(srt[(T >> 1) - 1], srt[T >> 1]), 0.5f);
      } else {
        // np.nanpercentile, method "linear": virtual index q*(T-1); numpy returns f64 for f32 input (its _lerp runs in
        // double: a + (b-a)*g, or b - (b-a)*(1-g) when g >= 0.5), so scale_ = q75 - q25 is a double
        const double vi = (tid == 0 ? 0.25 : 0.75) * (double)(T - 1);
        int lo = (int)floor(vi);
        if (lo > T - 1) lo = T - 1;
        const int hi = lo + 1 < T ? lo + 1 : T - 1;
        const double g = vi - (double)lo;
        const double a = (double)srt[lo], b = (double)srt[hi];
        const double d = b - a;
        s_q[tid >> 1] = g >= 0.5 ? b - d * (1.0 - g) : a + d * g;
      }
    }
    __syncthreads();
    const float med = s_med;
    double scale = s_q[1] - s_q[0];
    if (scale < 10.0 * 2.220446049250313e-16) scale = 1.0;       // sklearn _handle_zeros_in_scale (eps of the f64 scale)
    float* o = Out + row * T;
    for (int i = tid; i < T; i += nthr) {
      // transform(): X -= center_ (f32), X /= scale_ (f32 array / f64 scale, rounded back to f32)
      float v = (float)((double)(raw[i] - med) / scale);
      if (clamp) v = fminf(fmaxf(v, -lim), lim);
      o[i] = v;
    }
    __syncthreads();
  }
}

int megb200_robust_scale(const float* X, float* Out, long long rows, int T, int baseline_len, float clamp_lim, int clamp,
                         void* stream) {
  MEGB_CHECK_ARG(X && Out && rows > 0 && T >= 1 && T <= 4096, "robust_scale: bad arguments (1 <= T <= 4096)");
  MEGB_CHECK_ARG(baseline_len >= 1 && baseline_len <= T, "robust_scale: baseline_len must be in [1, T]");
  int P = 2;
  while (P < T) P <<= 1;
  int threads = P / 2;
  if (threads < 32) threads = 32;
  if (threads > 512) threads = 512;
  const size_t smem = (size_t)(P + T) * sizeof(float);
  long long blocks = rows;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  robust_scale_kernel<<<(int)blocks, threads, smem, (cudaStream_t)stream>>>(X, Out, rows, T, P, baseline_len, clamp_lim, clamp);
  MEGB_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------ multi-tensor Adam ("next" row N4)
// torch.optim.Adam (train.py:160-162) over all parameter tensors in ONE launch (+ a one-block counter bump).  table: n rows
// of 6 int64 = {param, grad, exp_avg, exp_avg_sq (device pointers, f32; complex tensors as 2*numel floats), numel,
// pointer to the tensor's own int64 step counter}.  torch keeps the step count PER PARAMETER (a tensor whose grad was
// None -- a subject absent from the batch -- does not advance), so the bias corrections are formed on the device from
// each row's counter.  blockIdx.y = tensor, the blocks of a row stride over its elements (float4 when the four pointers
// are 16-byte aligned).  HBM-bound: 7 floats of traffic per element (read p, g, m, v; write p, m, v).
struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h, float step_size,
                                            float inv_sqrt_bc2) {
  g = fmaf(h.weight_decay, p, g);                           // L2 form (torch.optim.Adam, not AdamW)
  m = fmaf(1.f - h.beta1, g - m, m);                        // exp_avg.lerp_(grad, 1 - beta1), as torch
  v = fmaf(h.beta2, v, (1.f - h.beta2) * g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + h.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const long long* __restrict__ table, const AdamHyper h) {
  const long long* e = table + 6 * blockIdx.y;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const long long n = e[4];
  const double t = (double)(*reinterpret_cast<const long long*>(e[5]) + 1);      // this update's 1-based step
  const float step_size = (float)((double)h.lr / (1.0 - pow((double)h.beta1, t)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)h.beta2, t)));
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  long long done = 0;
  if (vec) {
    const long long n4 = n >> 2;
#pragma unroll 2
    for (long long i = tid; i < n4; i += nthr) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_update(pp.x, gg.x, mm.x, vv.x, h, step_size, inv_sqrt_bc2);
      adam_update(pp.y, gg.y, mm.y, vv.y, h, step_size, inv_sqrt_bc2);
      adam_update(pp.z, gg.z, mm.z, vv.z, h, step_size, inv_sqrt_bc2);
      adam_update(pp.w, gg.w, mm.w, vv.w, h, step_size, inv_sqrt_bc2);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nthr) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_update(pp, g[i], mm, vv, h, step_size, inv_sqrt_bc2);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

// after the update: advance the counters of the tensors that took part (separate launch = no race with their readers)
__global__ void adam_bump_kernel(const long long* __restrict__ table, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) *reinterpret_cast<long long*>(table[6 * r + 5]) += 1;
}

int megb200_adam_multi(const long long* table, int n, float lr, float beta1, float beta2, float eps, float weight_decay,
                       void* stream) {
  MEGB_CHECK_ARG(table && n > 0 && n <= 65535, "adam_multi: bad arguments");
  MEGB_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_multi: bad hyper-parameters");
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  dim3 grid(32, n);        // largest tensor of the path: 655 k elements = 20 float4 per thread
  adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, h);
  MEGB_LAUNCH_OK();
  adam_bump_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, n);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, void* stream) {
  MEGB_CHECK_ARG(src && dst && n > 0, "cast: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = ew_grid(n, 256);
  if (src_dtype == MEGB200_F32 && dst_dtype == MEGB200_BF16)
    cast_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == MEGB200_BF16 && dst_dtype == MEGB200_F32)
    cast_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == MEGB200_F32 && dst_dtype == MEGB200_F32)
    cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (src_dtype == MEGB200_BF16 && dst_dtype == MEGB200_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else
    MEGB_CHECK_ARG(false, "cast: unknown dtype pair");
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_sa_softmax(const float* a, const float* mask, float* W, float* Wm, int D1, int C, void* stream) {
  MEGB_CHECK_ARG(a && W && Wm && D1 > 0 && C > 0, "sa_softmax: bad arguments");
  sa_softmax_kernel<<<D1, 128, 0, (cudaStream_t)stream>>>(a, mask, W, Wm, C);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_sa_softmax_bwd(const float* W, const float* dWm, const float* mask, float* da, int D1, int C,
                           void* stream) {
  MEGB_CHECK_ARG(W && dWm && da && D1 > 0 && C > 0, "sa_softmax_bwd: bad arguments");
  sa_softmax_bwd_kernel<<<D1, 128, 0, (cudaStream_t)stream>>>(W, dWm, mask, da, C);
  MEGB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
