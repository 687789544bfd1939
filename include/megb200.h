/* megb200 — C-ABI of the B200-native BrainEncoder + CLIPLoss training hot path.
 *
 * The reference (arayabrain/MEG-decoding) is pure Python/PyTorch and has no FFI; every entry point below
 * replaces the ATen calls issued by one reference op site (cited per function, paths relative to the
 * reference tree).  Plain pointers and sizes only; all pointers are DEVICE pointers unless stated.
 * Return value: 0 = ok, <0 = argument/shape error (MEGB200_ERR_*), >0 = cudaError_t.  No exceptions cross
 * the boundary; megb200_last_error() returns a thread-local message.  Every kernel is launched on the
 * `stream` argument (a cudaStream_t passed as void*); the library never synchronises, allocates or frees.
 *
 * Activation layout ("PC"): [B][T][Cs] — position-major, channels contiguous, Cs = channel stride (a multiple
 * of 8, >= the logical channel count C).  `dtype` selects the storage type of activations and packed weights:
 * MEGB200_F32 (check mode) or MEGB200_BF16.  `backend` selects the contraction engine: MEGB200_SIMT (fp32
 * CUDA-core FMA, any dtype) or MEGB200_TC (tcgen05.mma kind::f16 + TMA + TMEM accumulators, bf16 only).
 * Accumulation is always fp32.
 */
#ifndef MEGB200_H_
#define MEGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEGB200_VERSION 100

#define MEGB200_F32 0
#define MEGB200_BF16 1

#define MEGB200_SIMT 0
#define MEGB200_TC 1

/* tap-GEMM epilogues (all evaluated on the fp32 accumulator: exactly one rounding per stored tensor) */
#define MEGB200_EPI_STORE 0     /* Out = acc + bias + Res                       (storage dtype)                  */
#define MEGB200_EPI_STORE_F32 1 /* same, Out stored as f32: the pre-BatchNorm tensors u0/u1 (models.py:312,316)  */
#define MEGB200_EPI_GELU 2      /* Out2 = P = acc + bias, Out = gelu(P)         (models.py:373-374)              */
#define MEGB200_EPI_GLU 3       /* N = 2*Ch: Out2 = V = acc + bias [..,2Ch], Out[..,c] = V[c]*sigmoid(V[Ch+c])   */
                                /*                                              (models.py:319-320)              */

#define MEGB200_ERR_ARG (-1)
#define MEGB200_ERR_UNSUPPORTED (-2)
#define MEGB200_ERR_DRIVER (-3)

int megb200_version(void);
const char* megb200_last_error(void);
/* number of kernels this library has launched in this process (for the bench's gpu_launches claim) */
long long megb200_launch_count(void);

/* Kernel-variant selection (host-side, process-wide; takes effect at the next launch).  The library picks between
 * several hand-written forms of one contraction by problem size (e.g. the persistent CTA-pair tap-GEMM from 2 x #SM work
 * items up); these switches let a caller -- the test-suite -- force each form on any shape.  Names:
 *   "tapgemm_pair_min_items" (-1 = default), "tapgemm_force_1cta", "tapgemm_pair_halo", "tapgemm_halo",
 *   "tapgemm_l2_prefetch" (1 = next tile's activation boxes are prefetched into L2; default 0: measured no gain), "tapgemm_ksub" (1 = one k chunk per pipeline stage in the plain pair kernel; default two), "tapgemm_ragged_n", "tapgemm_direct_epi", "clip_dz_large", "rowloop_direct", "wgrad_splits" (0 = automatic),
 *   "wgrad_pair" (1 = CTA-pair weight-gradient kernel for ungrouped problems; default 0 = single-CTA kernel),
 *   "clip_dz_jt" (0 = automatic: brain-row tiles per CTA of the large-K dZ kernel),
 *   "clip_fused" (policy of the host-side loss: -1 auto, 0 never, 1 whenever supported).
 * Read-only (written by the launchers, for tests): "last_tapgemm_kernel" (0 single CTA, 1 CTA pair, 2 CTA pair + halo,
 * 3 single CTA + halo), "last_clip_dz_kernel" (0 small-K, 1 large-K), "last_wgrad_splits", "last_wgrad_kernel" (0 single
 * CTA, 1 CTA pair).
 * Initial values come from the MEGB200_* environment variables.  Results never depend on the variant beyond f32
 * summation order.  Returns MEGB200_ERR_ARG for an unknown name; get returns -(2^62) then. */
int megb200_set_option(const char* name, long long value);
long long megb200_get_option(const char* name);

/* ---------------------------------------------------------------------------------------------- packing */
/* X f32 [B][C][T] (NCW, as the collator delivers it: gwilliams2022.py:654-662) -> PC [B][T][Cs]; pad
 * channels C..Cs-1 are written as zero. */
int megb200_pack_x(const float* X, void* Xp, int B, int C, int T, int Cs, int dtype, void* stream);

/* Conv weight f32 [G][O][I][Kt] (nn.Conv1d layout, models.py:252-265,284-306,360-361) -> packed operand
 * [G][Np][Kt*Kp] of `dtype`.
 *   mode 0 (forward):  out[g][o][j*Kp + i] = w[g][o][i][j]            N = O, K = I
 *   mode 1 (dgrad):    out[g][i][j*Kp + o] = w[g][o][i][Kt-1-j]       N = I, K = O   (flipped taps)
 * Everything outside the logical extent is zero. */
int megb200_pack_weight(const float* w, void* out, int G, int O, int I, int Kt, int Np, int Kp, int mode,
                        int dtype, void* stream);

/* Same packing for many weights in one launch.  `table` is a DEVICE array of n entries of ten int64:
 * {src pointer, dst pointer, G, O, I, Kt, Np, Kp, mode, G*Np*Kt*Kp}. */
int megb200_pack_weight_multi(const long long* table, int n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------- contractions */
/* "tap GEMM": out[b][t][n] = sum_j sum_c A[b][t + (j - taps/2)*dil][c] * W[g][n][j*Kp + c]
 *                            (+ bias[g][n]) (+ Res[b][t][n]),   g = group[b] (0 when group == NULL),
 * zero outside [0,T) ("same" padding).  One kernel serves: the spatial-attention mix + shared 1x1 + per-subject
 * 1x1 chain pre-multiplied into one grouped GEMM keyed by subject index (models.py:220,269-272), Conv1d k=3
 * forward (models.py:310-319), its data gradient (flipped/transposed pack), and the 1x1 head convs
 * (models.py:373-374).  N = logical output channels; channels N..Cos-1 of Out are written as zero.  `epilogue`
 * is one of MEGB200_EPI_*; Out2/Co2s are used by the GELU and GLU epilogues only.  `stats` (optional, f32 epilogue
 * only): sums[2][N] += (sum, sum of squares) over all positions of Out — the train-mode BatchNorm statistics of
 * models.py:291,299, reduced in the epilogue instead of a separate pass (same result as megb200_bn_stats). */
int megb200_tapgemm(const void* A, int B, int T, int Ca, int Cas, const void* W, int N, int Np, int Kp, int taps,
                    int dil, const float* bias, const void* Res, int Crs, const int32_t* group, int G,
                    int epilogue, void* Out, int Cos, void* Out2, int Co2s, float* stats, int dtype, int backend,
                    void* stream);

/* Weight gradient: dW[g][m][c][j] += sum_{b: group[b]==g} sum_t dU[b][t][m] * Xin[b][t + (j - taps/2)*dil][c]
 * (f32, nn.Conv1d layout [G][M][C][taps]; accumulated — caller zeroes).  Replaces autograd's
 * convolution_backward weight path for every conv on the path, and the per-subject accumulation for the
 * grouped subject GEMM.  `order` (grouped mode only) lists the samples sorted by group
 * (megb200_group_order); `workspace` holds the split-K partial sums of the tensor-core engine (f32, at least
 * taps*M*C*4 bytes, more = more parallelism); both may be NULL for the SIMT engine. */
int megb200_wgrad(const void* dU, int M, int Ms, const void* Xin, int C, int Cs, int B, int T, int taps, int dil,
                  const int32_t* group, const int32_t* order, int G, float* dW, float* workspace, size_t ws_bytes,
                  int dtype, int backend, void* stream);
/* order = stable argsort of group[0..B) (values in [0,G)); bit-exact integer work, one CTA. */
int megb200_group_order(const int32_t* group, int B, int G, int32_t* order, void* stream);

/* out[g][c] += sum_{b: group[b]==g} sum_t U[b][t][c]   (bias gradients; f32 [G][C], accumulated) */
int megb200_colsum(const void* U, int B, int T, int C, int Cs, const int32_t* group, int G, float* out, int dtype,
                   void* stream);

/* Row-major f32 batched GEMM for the small per-step matrices of the subject block / spatial attention
 * (models.py:204-208 and the chain rule of SURVEY appendix A1-A2):
 *   C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],   op = transpose when the flag is set.
 * `workspace` (optional, f32) enables a deterministic split-K for long reductions with few output tiles. */
int megb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                  long long strideA, const float* Bm, int ldb, long long strideB, float beta, float* C, int ldc,
                  long long strideC, int batch, float* workspace, size_t ws_bytes, void* stream);

/* -------------------------------------------------------------------------------------- spatial attention */
/* W[o][:] = softmax(a[o][:]) ; Wm = W * mask[None,:]  (models.py:213 and :235-239 folded into the weights).
 * mask may be NULL (eval mode). a, W, Wm: f32 [D1][C]. */
int megb200_sa_softmax(const float* a, const float* mask, float* W, float* Wm, int D1, int C, void* stream);
/* da = W * (g - rowsum(W * g)),  g = dWm * mask   (softmax Jacobian) */
int megb200_sa_softmax_bwd(const float* W, const float* dWm, const float* mask, float* da, int D1, int C,
                           void* stream);

/* ---------------------------------------------------------------------------------------------- BatchNorm/GELU/GLU */
/* The pre-BatchNorm tensor U may be stored in f32 while the activations around it are bf16 (`u_dtype`): the
 * normalisation then sees the unrounded convolution output, as in the reference.
 * sums[0][c] += sum u, sums[1][c] += sum u^2 over all (b,t)   (models.py:291,299 train-mode statistics) */
int megb200_bn_stats(const void* U, long long rows, int C, int Cs, float* sums, int u_dtype, void* stream);
/* training != 0: mean/var from sums (biased var), running stats updated in place (momentum, unbiased var);
 * training == 0: running stats used.  Writes scale = gamma*rstd, shift = beta - mean*scale, mean, rstd. */
int megb200_bn_finalize(const float* sums, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, long long count, float eps, float momentum, int training, float* scale,
                        float* shift, float* mean, float* rstd, int C, void* stream);
/* H = gelu(scale*U + shift)   (models.py:314,317) */
int megb200_bn_gelu(const void* U, const float* scale, const float* shift, void* H, long long rows, int C, int Cs,
                    int u_dtype, int dtype, void* stream);
/* Training-mode BatchNorm1d + GELU in one launch (models.py:311-317 with nn.BatchNorm1d in train()): scale/shift come
 * from the batch sums (sums[0][c] = sum U, sums[1][c] = sum U^2 over `count` rows); also writes scale/shift/mean/rstd
 * for the backward pass and applies the momentum update of running_mean / running_var (unbiased variance). */
int megb200_bn_gelu_train(const void* U, const float* sums, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long count, float eps, float momentum, float* scale, float* shift,
                          float* mean, float* rstd, void* H, long long rows, int C, int Cs, int u_dtype, int dtype,
                          void* stream);
/* dY = dH * gelu'(scale*U+shift); sums2[0][c] += sum dY, sums2[1][c] += sum dY*xhat, xhat = (U-mean)*rstd.
 * dY may be NULL (sums only) when megb200_bn_gelu_bwd_apply follows. */
int megb200_bn_gelu_bwd_reduce(const void* dH, const void* U, const float* scale, const float* shift,
                               const float* mean, const float* rstd, void* dY, float* sums2, long long rows, int C,
                               int Cs, int u_dtype, int dtype, void* stream);
/* dU = gamma*rstd * (dY - sums2[0]/n - xhat * sums2[1]/n) with dY = dH * gelu'(scale*U+shift) recomputed on the fly
 * (SURVEY appendix A3); dbias as in megb200_bn_bwd_apply (may be NULL). */
int megb200_bn_gelu_bwd_apply(const void* dH, const void* U, const float* scale, const float* shift, const float* gamma,
                              const float* mean, const float* rstd, const float* sums2, long long count, void* dU,
                              float* dbias, long long rows, int C, int Cs, int u_dtype, int dtype, void* stream);
/* dU = gamma*rstd * (dY - sums2[0]/n - xhat * sums2[1]/n)   (SURVEY appendix A3);
 * dbias[c] += sum dU (the gradient of the conv bias in front of the BatchNorm; may be NULL) */
int megb200_bn_bwd_apply(const void* dY, const void* U, const float* gamma, const float* mean, const float* rstd,
                         const float* sums2, long long count, void* dU, float* dbias, long long rows, int C, int Cs,
                         int u_dtype, int dtype, void* stream);
/* eval-mode backward is never needed (no grad in eval). */

/* Out[.., c] = V[.., c] * sigmoid(V[.., C + c]),  V: [rows][Vs] with 2C logical channels (models.py:320) */
int megb200_glu(const void* V, void* Out, long long rows, int C, int Vs, int Os, int dtype, void* stream);
/* dV from dOut; dbias[0..2C) += column sums of dV (gradient of the conv2 bias; may be NULL) */
int megb200_glu_bwd(const void* dOut, const void* V, void* dV, float* dbias, long long rows, int C, int Vs, int Os,
                    int dtype, void* stream);
/* Q = gelu(P) ; dP = dQ * gelu'(P)   (models.py:373-374) */
int megb200_gelu(const void* P, void* Q, long long n, int dtype, void* stream);
/* dP = dQ * gelu'(P) on [rows][Cs]; dbias[c] += sum dP (gradient of the 1x1 head bias; may be NULL) */
int megb200_gelu_bwd(const void* dQ, const void* P, void* dP, float* dbias, long long rows, int C, int Cs, int dtype,
                     void* stream);
/* seq2static head (models.py:382): Zm[b][f] = mean_t Z[b][t][f] ; backward broadcasts dZm/T */
int megb200_time_mean(const void* Z, void* Zm, int B, int T, int F, int Fs, int dtype, void* stream);
int megb200_time_mean_bwd(const void* dZm, void* dZ, int B, int T, int F, int Fs, int dtype, void* stream);

/* nn.AvgPool1d(k, stride) over time, no padding (BrainEncoderSeq2Static puts AvgPool1d(3, 2) between its conv blocks,
 * models.py:483-485; "next" row N4): Out [B][To][Cs], To = (T - k) / stride + 1; backward spreads dOut / k over each window
 * (positions past the last window get zero).  nn.AdaptiveAvgPool1d(1) after the last block (models.py:487) is
 * megb200_time_mean. */
int megb200_avgpool1d(const void* X, void* Out, int B, int T, int C, int Cs, int k, int stride, int dtype, void* stream);
int megb200_avgpool1d_bwd(const void* dOut, void* dX, int B, int T, int C, int Cs, int k, int stride, int dtype,
                          void* stream);

/* ---------------------------------------------------------------------------------------------- CLIP loss */
/* Y [B][F][T] (NCW; `y_dtype` = MEGB200_F32 as the reference's collator delivers it, or MEGB200_BF16 when the data
 * pipeline stores the precomputed wav2vec2 targets in bf16 -- the loss rounds them to bf16 anyway, and the host->device
 * copy halves) -> Yp [B][T][F] of `dtype` and ysq[b] += sum Y^2 (f32, from the values as delivered).
 * T == 1 handles the static [B][F] targets.  loss.py:88-93 (reshape + norm). */
int megb200_clip_prep_y(const void* Y, int y_dtype, void* Yp, float* ysq, int B, int F, int T, int dtype, void* stream);
/* sq[b] += sum_d Z[b][d]^2   (loss.py:93) */
int megb200_rowsumsq(const void* Z, float* sq, int B, long long D, int dtype, void* stream);
/* raw[i][j] += sum_d Yp[i][d] * Zp[j][d]   f32 [By][Bz]  (loss.py:96, split-K over D with f32 atomics) */
int megb200_clip_logits(const void* Yp, const void* Zp, float* raw, int By, int Bz, long long D, int dtype,
                        int backend, void* stream);
/* L[i][j] = raw * exp(temp) / sqrt(ysq[i]*zsq[j])  (loss.py:92-99); per speech row i the partial
 * (max, sum exp) over the local brains; per brain column j the complete log-sum-exp over all speech rows. */
int megb200_clip_ce_partial(const float* raw, const float* ysq, const float* zsq, const float* temp, float* L,
                            float* row_max, float* row_sumexp, float* col_lse, int By, int Bz, void* stream);
/* row_lse[i] = log sum_w parts[w][1][i] * exp(parts[w][0][i])  (max-shifted): completes the per-speech-row log-sum-exp
 * over ALL brains from the W ranks' partial (max, sum exp) pairs; parts is [W][2][By] f32, the layout an all-gather of
 * the ranks' [2][By] blocks (row_max then row_sumexp of megb200_clip_ce_partial, allocated back to back) produces.
 * W = 1: the single-process loss.  One all-gather + this kernel replace a MAX and a SUM all-reduce. */
int megb200_clip_lse_merge(const float* parts, int W, int By, float* row_lse, void* stream);
/* Given the complete row LSEs: out[0] += loss contribution of this shard, out[1] += d loss / d temp;
 * coef[j][i] = dL[i][j] * exp(temp)/(|y_i||z_j|)  (`dtype`), cvec[j] = sum_i dL[i][j]*L[i][j] / zsq[j].
 * dL = (exp(L-row_lse[i]) + exp(L-col_lse[j]) - 2*[i == off + j]) * wscale, wscale = 1/(2*Bglob) for
 * reduction="mean", 1/2 for "sum" (loss.py:60,107).  out2/cvec are accumulated. */
int megb200_clip_ce_finish(const float* L, const float* row_lse, const float* col_lse, const float* ysq,
                           const float* zsq, const float* temp, int off, float wscale, float* out2, void* coef,
                           int ldc, float* cvec, int By, int Bz, int dtype, void* stream);
/* dZ[j][d] = gscale[0] * ( sum_i coef[j][i]*Yp[i][d] - Z[j][d]*cvec[j] )   (normalisation Jacobian folded).
 * coef is [Bz][ldc] with ldc >= By a multiple of 8 (ragged batches keep TMA's 16-byte stride rule). */
int megb200_clip_dz(const void* coef, int ldc, const void* Yp, const void* Z, const float* cvec, const float* gscale,
                    void* dZ, int By, int Bz, long long D, int dtype, int backend, void* stream);

/* ---- CLIPLoss at large batch with short embeddings (static targets: D = F, e.g. 512-d image features; BASELINE config
 * "GOD ... B x B logits sweep to 16k"): the B x B logits matrix is NEVER materialised (north_star: "logits matmul fused
 * with a warp-shuffle online-softmax CE").  bf16 / tensor-core engine only; the fp32 check mode uses the calls above.
 *
 * stats: one tcgen05 tile GEMM per 128 x 128 logits tile over the whole K = D, epilogue from TMEM: per-row online
 *        (max, sum exp) over the tile's columns -> row_part [ceil(Bz/128)][2][By], per-column (max, sum exp) over its rows
 *        (warp shuffles) -> col_part [ceil(By/128)][2][Bz], positives -> diag[j] = L[off+j][j].  megb200_clip_lse_merge
 *        completes both tables (W = number of tiles; in data parallel the ranks' row tables are all-gathered first).
 * loss:  out2[0] += wscale * sum_j (row_lse[off+j] + col_lse[j] - 2 diag[j])                           (loss.py:107)
 * coef:  backward with recompute, for brain columns [j_begin, j_begin + j_count): the tiles are contracted again,
 *        dL = (exp(L-row_lse[i]) + exp(L-col_lse[j]) - 2[i == off+j]) * wscale, and what leaves is
 *        coef[j - j_begin][i] = dL * exp(temp)/(|y_i||z_j|) (bf16, [j_count][ldc]), cvec[j] += sum_i dL*L / zsq[j],
 *        out2[1] += sum dL*L (d loss / d temp).  megb200_clip_dz then runs on that chunk (Z, dZ, cvec offset by j_begin):
 *        peak extra memory is one chunk of coefficients instead of three B x B matrices. */
int megb200_clip_fused_stats(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp,
                             int By, int Bz, long long D, int off, float* row_part, float* col_part, float* diag,
                             int dtype, int backend, void* stream);
int megb200_clip_fused_loss(const float* row_lse, const float* col_lse, const float* diag, int off, int Bz, float wscale,
                            float* out2, void* stream);
int megb200_clip_fused_coef(const void* Yp, const void* Zp, const float* ysq, const float* zsq, const float* temp,
                            const float* row_lse, const float* col_lse, int By, int Bz, long long D, int off, float wscale,
                            int j_begin, int j_count, void* coef, int ldc, float* cvec, float* out2, int dtype,
                            int backend, void* stream);

/* Top-k retrieval metric (the reference's Classifier, models.py:403-460; "next" row N1): given the f32 dot-product
 * matrix raw[By][Bz] (rows speech, columns brain) and the squared brain norms, rank[r] = number of brain embeddings
 * more similar (cosine) to speech row off + r than the matching one.  top-1 = mean(rank == 0), top-10 = mean(rank < 10).
 * Integer output, bit-exact given raw. */
int megb200_retrieval_rank(const float* raw, const float* zsq, int32_t* rank, int By, int Bz, int off, void* stream);

/* ---------------------------------------------------------------------------------------------- misc */
/* "Next" row N3: Gwilliams2022DatasetBase.__getitem__ / segment_to_task (gwilliams2022.py:130-152) for a batch, recordings
 * resident in HBM.  n_rec recordings (session keys) x n_tasks tasks; slot = rec*n_tasks + task.
 *   rec_base[slot]  device pointer of the f32 [C][rec_len[slot]] recording;  onsets[onset_off[slot] + k] = onset (samples) of
 *   the k-th segment of that task in that recording;  nseg_accum = cumsum(num_segments_foreach_task) (int64 [n_tasks]);
 *   seg[b] = speech-segment id, rec[b] = recording drawn for sample b (the reference draws it with np.random.choice, :134).
 * out: f32 [B][C][T]; index_out: int64 [B][3] = {task, i_in_task, onset} (bit-exact with the reference's integers);
 * *error (device int, OR-ed): 1 = segment id / recording index out of range, 2 = window outside its recording (zero-filled). */
int megb200_gather_segments(const long long* rec_base, const long long* rec_len, const long long* onset_off,
                            const long long* onsets, const long long* nseg_accum, int n_rec, int n_tasks,
                            const long long* seg, const int* rec, int B, int C, int T, float* out, long long* index_out,
                            int* error, void* stream);
/* "Next" row N2: the batch collator of the reference (Gwilliams2022Collator.forward, gwilliams2022.py:653-662) on the device.
 * X, Out: f32 [rows][T] with rows = B*C (Out may alias X).  Per row: subtract the mean of the first baseline_len samples
 * (baseline_correction_single, preproc_utils.py:128-142), then sklearn RobustScaler over time -- (x - median) / (q75 - q25),
 * numpy linear-interpolated quantiles, a scale below 10 eps becomes 1 (scaleAndClamp, preproc_utils.py:69-90) -- then, if
 * clamp != 0, clamp to [-clamp_lim, clamp_lim].  1 <= baseline_len <= T <= 4096. */
int megb200_robust_scale(const float* X, float* Out, long long rows, int T, int baseline_len, float clamp_lim, int clamp,
                         void* stream);
/* "Next" row N4: torch.optim.Adam (train.py:160-162: Adam over the encoder's and the loss's parameters) for all tensors in one
 * launch.  table (device): n rows of 6 int64 = {param, grad, exp_avg, exp_avg_sq pointers (f32; a complex tensor counts as
 * 2*numel floats), numel, pointer to the tensor's int64 step counter (device; read for the bias corrections 1-beta^(count+1),
 * then incremented)}.  Per-tensor counters because torch does not advance a parameter whose grad is None.  weight_decay is
 * the L2 form (added to the gradient), amsgrad is not supported -- the reference uses neither. */
int megb200_adam_multi(const long long* table, int n, float lr, float beta1, float beta2, float eps, float weight_decay,
                       void* stream);
/* ---------------------------------------------------------------------------------------------- peer memory (data parallel)
 * The all-gather of the packed speech targets (189 MB per rank, 1.3 GB received per rank at 8 GPUs) runs on
 * the COPY ENGINES over NVLink peer memory, so that no SM is taken from the persistent one-CTA-per-SM contraction
 * kernels that execute meanwhile: every rank exports its slab once (CUDA IPC), maps the peers' slabs, and each step
 * pulls the peers' slots with plain asynchronous copies on a side stream.
 *   export: handle (64 bytes, HOST memory) of the cudaMalloc allocation that contains dev_ptr, and dev_ptr's byte offset
 *           inside it (the caching allocator sub-allocates);
 *   import: maps a peer process's allocation into this process (peer access enabled lazily) -> *base_out; the caller
 *           adds the offset.  One import per (process, peer allocation); release unmaps it.
 *   copy:   cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `stream` -- device/peer/pinned-host pointers (UVA). */
int megb200_ipc_export(const void* dev_ptr, unsigned char* handle64, long long* offset_bytes);
int megb200_ipc_import(const unsigned char* handle64, void** base_out);
int megb200_ipc_release(void* base);
int megb200_copy_async(void* dst, const void* src, size_t bytes, void* stream);

/* d[i] = s[i] cast between f32 and the activation dtype (n elements) */
int megb200_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, void* stream);
#ifdef __cplusplus
}
#endif
#endif /* MEGB200_H_ */
