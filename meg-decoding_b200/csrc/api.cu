// C-ABI entry points that dispatch between the two contraction engines, plus library-level plumbing
// (error string, SM count, TMA descriptor encoding through the runtime-resolved driver entry point).
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace megb200 {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

static long long env_ll(const char* name, long long dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoll(e) : dflt;
}

Options& options() {
  static Options o = {
      env_ll("MEGB200_TAPGEMM_2CTA_MIN_ITEMS", -1),
      getenv("MEGB200_TAPGEMM_1CTA") != nullptr ? 1 : 0,
      (int)env_ll("MEGB200_TAPGEMM_PAIR_HALO", 1),
      (int)env_ll("MEGB200_TAPGEMM_HALO", 0),
      getenv("MEGB200_TAPGEMM_RAGGED_N") != nullptr ? 1 : 0,
      getenv("MEGB200_TAPGEMM_DIRECT_EPI") != nullptr ? 1 : 0,
      (int)env_ll("MEGB200_TAPGEMM_L2_PREFETCH", 0),
      (int)env_ll("MEGB200_TAPGEMM_KSUB", 0),
      getenv("MEGB200_CLIP_DZ_LARGE") != nullptr ? 1 : 0,
      getenv("MEGB200_ROWLOOP_DIRECT") != nullptr ? 1 : 0,
      (int)env_ll("MEGB200_WGRAD_SPLITS", 0),
      (int)env_ll("MEGB200_WGRAD_PAIR", 0),
      (int)env_ll("MEGB200_CLIP_DZ_JT", 0),
      (int)env_ll("MEGB200_CLIP_FUSED", -1),
      -1, -1, -1, -1,
  };
  return o;
}

struct OptionEntry { const char* name; int kind; /* 0 = long long, 1 = int */ size_t off; };
static const OptionEntry kOptionTable[] = {
    {"tapgemm_pair_min_items", 0, offsetof(Options, tapgemm_pair_min_items)},
    {"tapgemm_force_1cta", 1, offsetof(Options, tapgemm_force_1cta)},
    {"tapgemm_pair_halo", 1, offsetof(Options, tapgemm_pair_halo)},
    {"tapgemm_halo", 1, offsetof(Options, tapgemm_halo)},
    {"tapgemm_ragged_n", 1, offsetof(Options, tapgemm_ragged_n)},
    {"tapgemm_direct_epi", 1, offsetof(Options, tapgemm_direct_epi)},
    {"tapgemm_l2_prefetch", 1, offsetof(Options, tapgemm_l2_prefetch)},
    {"tapgemm_ksub", 1, offsetof(Options, tapgemm_ksub)},
    {"clip_dz_large", 1, offsetof(Options, clip_dz_large)},
    {"rowloop_direct", 1, offsetof(Options, rowloop_direct)},
    {"wgrad_splits", 1, offsetof(Options, wgrad_splits)},
    {"wgrad_pair", 1, offsetof(Options, wgrad_pair)},
    {"clip_dz_jt", 1, offsetof(Options, clip_dz_jt)},
    {"clip_fused", 1, offsetof(Options, clip_fused)},
    {"last_tapgemm_kernel", 1, offsetof(Options, last_tapgemm_kernel)},
    {"last_clip_dz_kernel", 1, offsetof(Options, last_clip_dz_kernel)},
    {"last_wgrad_splits", 1, offsetof(Options, last_wgrad_splits)},
    {"last_wgrad_kernel", 1, offsetof(Options, last_wgrad_kernel)},
};

static const OptionEntry* find_option(const char* name) {
  if (name == nullptr) return nullptr;
  for (const OptionEntry& e : kOptionTable)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap(map, base, 2, 128, rank, dims, strides_bytes, box);
}

int encode_tmap(CUtensorMap* map, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box) {
  // Driver entry points need the primary context current on THIS thread; autograd's backward threads may not have
  // issued a runtime call yet (cuTensorMapEncodeTiled then fails with CUDA_ERROR_INVALID_CONTEXT).
  // Once per thread: cudaFree synchronises the device, so doing it per call would serialise host and GPU.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return MEGB200_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA operand base %p is not 16-byte aligned", base);
    return MEGB200_ERR_ARG;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_last_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes", (unsigned long long)gstr[i - 1], i);
        return MEGB200_ERR_ARG;
      }
    }
  }
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0);
    return MEGB200_ERR_DRIVER;
  }
  return 0;
}

int tc_tapgemm(const void* A, int B, int T, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps, int dil,
               const float* bias, const void* Res, int Crs, const int32_t* group, int G, int epi, void* Out, int Cos,
               void* Out2, int Co2s, float* stats, int* stats_fused, cudaStream_t st);
int tc_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int T, int taps, int dil,
             const int32_t* group, const int32_t* order, int G, float* dW, float* workspace, size_t ws_bytes,
             cudaStream_t st);
int tc_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, cudaStream_t st);
int tc_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
               void* dZ, int By, int Bz, long long D, cudaStream_t st);
int tc_clip_fused(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp, int By, int Bz,
                  long long D, int off, int mode, float* row_part, float* col_part, float* diag, const float* row_lse,
                  const float* col_lse, float wscale, int j_begin, int j_count, void* coef, int ldc, float* cvec,
                  float* out2, cudaStream_t st);
int tc_clip_fused_loss(const float* row_lse, const float* col_lse, const float* diag, int off, int Bz, float wscale,
                       float* out2, cudaStream_t st);
}  // namespace tc

template <typename T>
int simt_tapgemm(const void* A, int B, int Tn, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps,
                 int dil, const float* bias, const void* Res, int Crs, const int32_t* group, int epi, void* Out,
                 int Cos, void* Out2, int Co2s, cudaStream_t st);
template <typename T>
int simt_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int Tn, int taps, int dil,
               const int32_t* group, float* dW, cudaStream_t st);
template <typename T>
int simt_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, cudaStream_t st);
template <typename T>
int simt_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
                 void* dZ, int By, int Bz, long long D, cudaStream_t st);

// stable counting sort of the sample indices by group: one CTA, bit-exact integer work
__global__ void group_order_kernel(const int32_t* __restrict__ group, int B, int G, int32_t* __restrict__ order) {
  extern __shared__ int cnt[];  // [G + 1]
  for (int i = threadIdx.x; i <= G; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  // indices outside [0, G) are ignored here (the Python boundary raises IndexError for them before any launch, like the
  // reference's ModuleList; a raw C-ABI caller must not be able to scribble over shared memory)
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int g = group[i];
    if ((unsigned)g < (unsigned)G) atomicAdd(&cnt[g + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int g = 0; g < G; ++g) cnt[g + 1] += cnt[g];  // cnt[g] = first slot of group g
  }
  __syncthreads();
  // stability: thread g walks the samples in order and emits its own group's members
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    int pos = cnt[g];
    for (int i = 0; i < B; ++i)
      if (group[i] == g) order[pos++] = i;
  }
}

}  // namespace megb200

using namespace megb200;

extern "C" {

int megb200_version(void) { return MEGB200_VERSION; }
const char* megb200_last_error(void) { return g_err; }
long long megb200_launch_count(void) { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int megb200_set_option(const char* name, long long value) {
  const OptionEntry* e = find_option(name);
  MEGB_CHECK_ARG(e != nullptr, "set_option: unknown option '%s'", name ? name : "(null)");
  char* base = reinterpret_cast<char*>(&options());
  if (e->kind == 0) *reinterpret_cast<long long*>(base + e->off) = value;
  else *reinterpret_cast<int*>(base + e->off) = (int)value;
  return 0;
}

long long megb200_get_option(const char* name) {
  const OptionEntry* e = find_option(name);
  if (e == nullptr) {
    set_last_error("get_option: unknown option '%s'", name ? name : "(null)");
    return -(1ll << 62);
  }
  const char* base = reinterpret_cast<const char*>(&options());
  return e->kind == 0 ? *reinterpret_cast<const long long*>(base + e->off)
                      : (long long)*reinterpret_cast<const int*>(base + e->off);
}

// ---------------------------------------------------------------------------------------------- peer memory
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

int megb200_ipc_export(const void* dev_ptr, unsigned char* handle64, long long* offset_bytes) {
  MEGB_CHECK_ARG(dev_ptr && handle64 && offset_bytes, "ipc_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  static GetAddressRangeFn get_range = nullptr;
  if (get_range == nullptr) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(nullptr);
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || fp == nullptr) {
      set_last_error("ipc_export: cuMemGetAddressRange is not available from the driver");
      return MEGB200_ERR_DRIVER;
    }
    get_range = reinterpret_cast<GetAddressRangeFn>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = get_range(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr);
  if (r != CUDA_SUCCESS) {
    set_last_error("ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
    return MEGB200_ERR_DRIVER;
  }
  cudaIpcMemHandle_t h;
  MEGB_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle64, &h, 64);
  *offset_bytes = (long long)((uintptr_t)dev_ptr - (uintptr_t)base);
  return 0;
}

int megb200_ipc_import(const unsigned char* handle64, void** base_out) {
  MEGB_CHECK_ARG(handle64 && base_out, "ipc_import: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  MEGB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = p;
  return 0;
}

int megb200_ipc_release(void* base) {
  MEGB_CHECK_ARG(base != nullptr, "ipc_release: null pointer");
  MEGB_CUDA_OK(cudaIpcCloseMemHandle(base));
  return 0;
}

int megb200_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  MEGB_CHECK_ARG(dst && src, "copy_async: null pointer");
  if (bytes == 0) return 0;
  MEGB_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int megb200_group_order(const int32_t* group, int B, int G, int32_t* order, void* stream) {
  MEGB_CHECK_ARG(group && order && B > 0 && G > 0 && G < 8192, "group_order: bad arguments");
  group_order_kernel<<<1, 256, (G + 1) * sizeof(int), (cudaStream_t)stream>>>(group, B, G, order);
  MEGB_LAUNCH_OK();
  return 0;
}

int megb200_tapgemm(const void* A, int B, int T, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps,
                    int dil, const float* bias, const void* Res, int Crs, const int32_t* group, int G, int epilogue,
                    void* Out, int Cos, void* Out2, int Co2s, float* stats, int dtype, int backend, void* stream) {
  MEGB_CHECK_ARG(stats == nullptr || epilogue == MEGB200_EPI_STORE_F32, "tapgemm: statistics need the f32 epilogue");
  MEGB_CHECK_ARG(A && W && Out && B > 0 && T > 0 && Ca > 0 && N > 0 && taps >= 1 && (taps & 1) && dil >= 1 && G >= 1,
                 "tapgemm: bad arguments");
  MEGB_CHECK_ARG(epilogue >= MEGB200_EPI_STORE && epilogue <= MEGB200_EPI_GLU, "tapgemm: unknown epilogue %d", epilogue);
  const int n_out = epilogue == MEGB200_EPI_GLU ? N / 2 : N;
  MEGB_CHECK_ARG(Cas >= Ca && Cos >= n_out && Np >= N && Kp >= Ca, "tapgemm: strides/padding smaller than extents");
  MEGB_CHECK_ARG(Res == nullptr || Crs >= N, "tapgemm: residual stride smaller than N");
  if (epilogue == MEGB200_EPI_GLU || epilogue == MEGB200_EPI_GELU) {
    MEGB_CHECK_ARG(Out2 != nullptr && Co2s >= N, "tapgemm: this epilogue needs the second output (stride >= N)");
    MEGB_CHECK_ARG(epilogue != MEGB200_EPI_GLU || (N % 2 == 0 && Res == nullptr), "tapgemm: GLU needs even N, no residual");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == MEGB200_TC) {
    if (dtype != MEGB200_BF16) {
      set_last_error("tapgemm: the tensor-core engine needs bf16 activations");
      return MEGB200_ERR_UNSUPPORTED;
    }
    int fused = 0;
    int rc = tc::tc_tapgemm(A, B, T, Ca, Cas, W, N, Np, Kp, taps, dil, bias, Res, Crs, group, G, epilogue, Out, Cos,
                            Out2, Co2s, stats, &fused, st);
    if (rc != 0 || stats == nullptr || fused) return rc;
    return megb200_bn_stats(Out, (long long)B * T, N, Cos, stats, MEGB200_F32, stream);   // kernel variant without fusion
  }
  MEGB_CHECK_ARG(backend == MEGB200_SIMT, "tapgemm: unknown backend %d", backend);
  int rc;
  if (dtype == MEGB200_F32)
    rc = simt_tapgemm<float>(A, B, T, Ca, Cas, W, N, Np, Kp, taps, dil, bias, Res, Crs, group, epilogue, Out, Cos, Out2,
                             Co2s, st);
  else if (dtype == MEGB200_BF16)
    rc = simt_tapgemm<__nv_bfloat16>(A, B, T, Ca, Cas, W, N, Np, Kp, taps, dil, bias, Res, Crs, group, epilogue, Out,
                                     Cos, Out2, Co2s, st);
  else
    MEGB_CHECK_ARG(false, "tapgemm: unknown dtype %d", dtype);
  if (rc != 0 || stats == nullptr) return rc;
  return megb200_bn_stats(Out, (long long)B * T, N, Cos, stats, MEGB200_F32, stream);
}

int megb200_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int T, int taps, int dil,
                  const int32_t* group, const int32_t* order, int G, float* dW, float* workspace, size_t ws_bytes,
                  int dtype, int backend, void* stream) {
  MEGB_CHECK_ARG(dU && Xin && dW && M > 0 && C > 0 && B > 0 && T > 0 && taps >= 1 && (taps & 1) && dil >= 1 && G >= 1,
                 "wgrad: bad arguments");
  MEGB_CHECK_ARG(Ms >= M && Cs >= C, "wgrad: strides smaller than extents");
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == MEGB200_TC) {
    if (dtype != MEGB200_BF16) {
      set_last_error("wgrad: the tensor-core engine needs bf16 activations");
      return MEGB200_ERR_UNSUPPORTED;
    }
    return tc::tc_wgrad(dU, M, Ms, Xin, C, Cs, B, T, taps, dil, group, order, G, dW, workspace, ws_bytes, st);
  }
  MEGB_CHECK_ARG(backend == MEGB200_SIMT, "wgrad: unknown backend %d", backend);
  if (dtype == MEGB200_F32) return simt_wgrad<float>(dU, M, Ms, Xin, C, Cs, B, T, taps, dil, group, dW, st);
  if (dtype == MEGB200_BF16) return simt_wgrad<__nv_bfloat16>(dU, M, Ms, Xin, C, Cs, B, T, taps, dil, group, dW, st);
  MEGB_CHECK_ARG(false, "wgrad: unknown dtype %d", dtype);
}

int megb200_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, int dtype,
                        int backend, void* stream) {
  MEGB_CHECK_ARG(Yp && Zp && raw && By > 0 && Bz > 0 && D > 0, "clip_logits: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == MEGB200_TC) {
    if (dtype != MEGB200_BF16) {
      set_last_error("clip_logits: the tensor-core engine needs bf16 operands");
      return MEGB200_ERR_UNSUPPORTED;
    }
    return tc::tc_clip_logits(Yp, Zp, raw, By, Bz, D, st);
  }
  MEGB_CHECK_ARG(backend == MEGB200_SIMT, "clip_logits: unknown backend %d", backend);
  if (dtype == MEGB200_F32) return simt_clip_logits<float>(Yp, Zp, raw, By, Bz, D, st);
  if (dtype == MEGB200_BF16) return simt_clip_logits<__nv_bfloat16>(Yp, Zp, raw, By, Bz, D, st);
  MEGB_CHECK_ARG(false, "clip_logits: unknown dtype %d", dtype);
}

int megb200_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
                    void* dZ, int By, int Bz, long long D, int dtype, int backend, void* stream) {
  MEGB_CHECK_ARG(coef && Yp && Z && cvec && gscale && dZ && By > 0 && Bz > 0 && D > 0 && ldc >= By,
                 "clip_dz: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == MEGB200_TC) {
    if (dtype != MEGB200_BF16) {
      set_last_error("clip_dz: the tensor-core engine needs bf16 operands");
      return MEGB200_ERR_UNSUPPORTED;
    }
    return tc::tc_clip_dz(coef, ldc, Yp, Z, cvec, gscale, dZ, By, Bz, D, st);
  }
  MEGB_CHECK_ARG(backend == MEGB200_SIMT, "clip_dz: unknown backend %d", backend);
  if (dtype == MEGB200_F32) return simt_clip_dz<float>(coef, ldc, Yp, Z, cvec, gscale, dZ, By, Bz, D, st);
  if (dtype == MEGB200_BF16) return simt_clip_dz<__nv_bfloat16>(coef, ldc, Yp, Z, cvec, gscale, dZ, By, Bz, D, st);
  MEGB_CHECK_ARG(false, "clip_dz: unknown dtype %d", dtype);
}

// ---------------------------------------------------------------------------------------------- fused large-batch CLIP
int megb200_clip_fused_stats(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp,
                             int By, int Bz, long long D, int off, float* row_part, float* col_part, float* diag,
                             int dtype, int backend, void* stream) {
  MEGB_CHECK_ARG(Yp && Zp && ysq && zsq && temp && row_part && col_part && diag && By > 0 && Bz > 0 && D > 0,
                 "clip_fused_stats: bad arguments");
  if (backend != MEGB200_TC || dtype != MEGB200_BF16) {
    set_last_error("clip_fused_stats: tensor-core engine / bf16 only (the fp32 check mode materialises the logits)");
    return MEGB200_ERR_UNSUPPORTED;
  }
  return tc::tc_clip_fused(Yp, Zp, ysq, zsq, temp, By, Bz, D, off, 0, row_part, col_part, diag, nullptr, nullptr, 0.f, 0,
                           Bz, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int megb200_clip_fused_loss(const float* row_lse, const float* col_lse, const float* diag, int off, int Bz, float wscale,
                            float* out2, void* stream) {
  MEGB_CHECK_ARG(row_lse && col_lse && diag && out2 && Bz > 0 && off >= 0, "clip_fused_loss: bad arguments");
  return tc::tc_clip_fused_loss(row_lse, col_lse, diag, off, Bz, wscale, out2, (cudaStream_t)stream);
}

int megb200_clip_fused_coef(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp,
                            const float* row_lse, const float* col_lse, int By, int Bz, long long D, int off, float wscale,
                            int j_begin, int j_count, void* coef, int ldc, float* cvec, float* out2, int dtype,
                            int backend, void* stream) {
  MEGB_CHECK_ARG(Yp && Zp && ysq && zsq && temp && row_lse && col_lse && coef && cvec && out2 && By > 0 && Bz > 0 && D > 0,
                 "clip_fused_coef: bad arguments");
  MEGB_CHECK_ARG(ldc >= By && ldc % 8 == 0, "clip_fused_coef: ldc must be >= By and a multiple of 8");
  if (backend != MEGB200_TC || dtype != MEGB200_BF16) {
    set_last_error("clip_fused_coef: tensor-core engine / bf16 only");
    return MEGB200_ERR_UNSUPPORTED;
  }
  return tc::tc_clip_fused(Yp, Zp, ysq, zsq, temp, By, Bz, D, off, 1, nullptr, nullptr, nullptr, row_lse, col_lse, wscale,
                           j_begin, j_count, coef, ldc, cvec, out2, (cudaStream_t)stream);
}

}  // extern "C"
