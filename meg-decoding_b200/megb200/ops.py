"""Tensor-level wrappers over the C-ABI: every method launches hand-written kernels on torch's current CUDA
stream and writes into caller-provided tensors.  torch supplies device memory and streams only.

``CudaOps(dtype, backend)`` fixes the activation storage type (torch.bfloat16 | torch.float32) and the contraction
engine (TC = tcgen05, SIMT = fp32 CUDA cores).  The engine code (engine.py) is written against this interface; the
test-suite's oracle implements the same interface in plain torch to check the orchestration on CPU.
"""
from __future__ import annotations

import torch

from . import lib as L


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def pad16(n: int) -> int:
    return (n + 15) // 16 * 16


def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _ptr(t):
    return None if t is None else t.data_ptr()


class CudaOps:
    def __init__(self, dtype: torch.dtype = torch.bfloat16, backend: str = "tc"):
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("activation dtype must be bfloat16 or float32")
        if backend not in ("tc", "simt"):
            raise ValueError("backend must be 'tc' or 'simt'")
        if backend == "tc" and dtype != torch.bfloat16:
            raise ValueError("the tensor-core engine computes in bf16; use backend='simt' for the fp32 check mode")
        L.load()
        self.act_dtype = dtype
        self.dt = L.BF16 if dtype == torch.bfloat16 else L.F32
        self.be = L.TC if backend == "tc" else L.SIMT
        self.backend = backend
        self._ws = {}
        self.prof = None      # set to a list to time individual contraction launches with CUDA events (bench.py)

    # ------------------------------------------------------------------ helpers
    def _prof_begin(self):
        if self.prof is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def _prof_end(self, ev, name, flops):
        if ev is None:
            return
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        self.prof.append((name, flops, ev, end))

    @staticmethod
    def _stream():
        # raw handle of torch's current stream on the current device (one C call; torch.cuda.current_stream() builds a
        # Stream object and costs ~6 us, 160 times per step)
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())

    def _chk_act(self, *ts):
        for t in ts:
            if t is not None:
                assert t.is_cuda and t.dtype == self.act_dtype and t.is_contiguous(), (t.dtype, t.shape, t.is_contiguous())

    @staticmethod
    def _chk_f32(*ts):
        for t in ts:
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape)

    def workspace(self, device, nbytes: int = 96 << 20):
        key = (device.index, torch._C._cuda_getCurrentRawStream(device.index))
        ws = self._ws.get(key)
        if ws is None or ws.numel() * 4 < nbytes:
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
            self._ws[key] = ws
        return ws

    def empty(self, *shape, device):
        return torch.empty(*shape, dtype=self.act_dtype, device=device)

    # ------------------------------------------------------------------ packing
    def pack_x(self, X, Xp):
        self._chk_f32(X)
        self._chk_act(Xp)
        B, C, T = X.shape
        assert Xp.shape[0] == B and Xp.shape[1] == T
        L.call("megb200_pack_x", X.data_ptr(), Xp.data_ptr(), B, C, T, Xp.shape[2], self.dt, self._stream())

    def pack_weight(self, w, out, mode: int):
        """w f32 [G,O,I,Kt] -> out [G,Np,Kt*Kp]; mode 0 forward (N=O,K=I), 1 dgrad (N=I,K=O, flipped taps)."""
        self._chk_f32(w)
        self._chk_act(out)
        G, O, I, Kt = w.shape
        Np = out.shape[1]
        Kp = out.shape[2] // Kt
        assert out.shape[0] == G and out.shape[2] == Kt * Kp
        L.call("megb200_pack_weight", w.data_ptr(), out.data_ptr(), G, O, I, Kt, Np, Kp, mode, self.dt, self._stream())

    def pack_weights(self, items):
        """items: list of (w f32 [G,O,I,Kt], out [G,Np,Kt*Kp], mode).  One launch for all of them; the descriptor table
        is cached while the tensors' addresses stay the same (parameters are updated in place by the optimizer)."""
        key = tuple((w.data_ptr(), o.data_ptr(), m) for w, o, m in items)
        cached = getattr(self, "_pack_tables", None)
        if cached is None:
            cached = self._pack_tables = {}
        tab = cached.get(key)
        if tab is None:
            rows = []
            for w, o, m in items:
                self._chk_f32(w)
                self._chk_act(o)
                G, O, I, Kt = w.shape
                Np, Kp = o.shape[1], o.shape[2] // Kt
                assert o.shape[0] == G and o.shape[2] == Kt * Kp
                rows.append([w.data_ptr(), o.data_ptr(), G, O, I, Kt, Np, Kp, m, G * Np * Kt * Kp])
            if len(cached) > 64:
                cached.clear()
            # pinned source: legal (and replay-safe: the cache keeps it alive) even if the first use is under graph capture
            host = torch.tensor(rows, dtype=torch.int64).pin_memory()
            tab = cached[key] = host.to(items[0][0].device, non_blocking=True)
            cached[("host",) + key] = host
        L.call("megb200_pack_weight_multi", tab.data_ptr(), len(items), self.dt, self._stream())

    # ------------------------------------------------------------------ contractions
    def tapgemm(self, A, Ca, W, N, taps, dil, bias, Res, group, Out, epi="store", Out2=None, stats=None):
        """epi: "store" | "store_f32" (Out is f32: pre-BatchNorm tensors) | "gelu" (Out2 = pre-activation, Out =
        gelu) | "glu" (N = 2*Ch; Out2 = V [.., 2Ch], Out = GLU(V) [.., Ch])."""
        code = {"store": L.EPI_STORE, "store_f32": L.EPI_STORE_F32, "gelu": L.EPI_GELU, "glu": L.EPI_GLU}[epi]
        self._chk_act(A, W, Res, Out2)
        if epi == "store_f32":
            self._chk_f32(Out)
        else:
            self._chk_act(Out)
        self._chk_f32(bias, stats)
        assert stats is None or (epi == "store_f32" and stats.numel() == 2 * N)
        B, T, Cas = A.shape
        G, Np, Ktot = W.shape
        Kp = Ktot // taps
        assert Out.shape[0] == B and Out.shape[1] == T
        assert (Out2 is not None) == (epi in ("gelu", "glu"))
        if group is not None:
            assert group.dtype == torch.int32 and group.is_cuda and group.numel() == B
        else:
            assert G == 1
        if bias is not None:
            assert bias.numel() == G * N
        ev = self._prof_begin()
        L.call("megb200_tapgemm", A.data_ptr(), B, T, Ca, Cas, W.data_ptr(), N, Np, Kp, taps, dil, _ptr(bias),
               _ptr(Res), 0 if Res is None else Res.shape[2], _ptr(group), G, code, Out.data_ptr(), Out.shape[2],
               _ptr(Out2), 0 if Out2 is None else Out2.shape[2], _ptr(stats), self.dt, self.be, self._stream())
        self._prof_end(ev, "tapgemm", 2.0 * B * T * N * taps * Ca)

    def group_order(self, group, G, order):
        assert group.dtype == torch.int32 and order.dtype == torch.int32
        L.call("megb200_group_order", group.data_ptr(), group.numel(), G, order.data_ptr(), self._stream())

    def wgrad(self, dU, M, Xin, C, taps, dil, group, order, dW):
        """dW f32 [G,M,C,taps] += ...  (caller zeroes)."""
        self._chk_act(dU, Xin)
        self._chk_f32(dW)
        B, T, Ms = dU.shape
        G = dW.shape[0] if dW.dim() == 4 else 1
        assert dW.numel() == G * M * C * taps
        ws = self.workspace(dU.device) if self.be == L.TC else None
        ev = self._prof_begin()
        L.call("megb200_wgrad", dU.data_ptr(), M, Ms, Xin.data_ptr(), C, Xin.shape[2], B, T, taps, dil, _ptr(group),
               _ptr(order), G, dW.data_ptr(), _ptr(ws), 0 if ws is None else ws.numel() * 4, self.dt, self.be,
               self._stream())
        self._prof_end(ev, "wgrad", 2.0 * B * T * M * C * taps)

    def colsum(self, U, C, group, out):
        self._chk_act(U)
        self._chk_f32(out)
        B, T, Cs = U.shape
        G = out.numel() // C
        L.call("megb200_colsum", U.data_ptr(), B, T, C, Cs, _ptr(group), G, out.data_ptr(), self.dt, self._stream())

    def matmul(self, A, Bm, out, transA=False, transB=False, alpha=1.0, beta=0.0):
        """out = alpha * op(A) @ op(B) + beta * out for small f32 matrices; 2-D operands broadcast over the batch."""
        self._chk_f32(A, Bm, out)
        batch = out.shape[0] if out.dim() == 3 else 1
        M, N = out.shape[-2], out.shape[-1]
        K = A.shape[-2] if transA else A.shape[-1]
        assert (A.shape[-1] if transA else A.shape[-2]) == M, (A.shape, out.shape, transA)
        assert (Bm.shape[-2] if transB else Bm.shape[-1]) == N and (Bm.shape[-1] if transB else Bm.shape[-2]) == K
        sA = A.shape[-2] * A.shape[-1] if A.dim() == 3 else 0
        sB = Bm.shape[-2] * Bm.shape[-1] if Bm.dim() == 3 else 0
        sC = M * N if out.dim() == 3 else 0
        ws = self.workspace(out.device)
        L.call("megb200_sgemm", int(transA), int(transB), M, N, K, float(alpha), A.data_ptr(), A.shape[-1], sA,
               Bm.data_ptr(), Bm.shape[-1], sB, float(beta), out.data_ptr(), N, sC, batch, ws.data_ptr(),
               ws.numel() * 4, self._stream())

    # ------------------------------------------------------------------ spatial attention
    def sa_softmax(self, a, mask, W, Wm):
        self._chk_f32(a, mask, W, Wm)
        L.call("megb200_sa_softmax", a.data_ptr(), _ptr(mask), W.data_ptr(), Wm.data_ptr(), a.shape[0], a.shape[1],
               self._stream())

    def sa_softmax_bwd(self, W, dWm, mask, da):
        self._chk_f32(W, dWm, mask, da)
        L.call("megb200_sa_softmax_bwd", W.data_ptr(), dWm.data_ptr(), _ptr(mask), da.data_ptr(), W.shape[0],
               W.shape[1], self._stream())

    # ------------------------------------------------------------------ BN / GELU / GLU
    @staticmethod
    def _udt(U):
        """Storage code of a pre-BatchNorm tensor (f32 in bf16 mode: models.py:312,316 normalise the unrounded conv)."""
        assert U.is_cuda and U.is_contiguous() and U.dtype in (torch.float32, torch.bfloat16)
        return L.F32 if U.dtype == torch.float32 else L.BF16

    def bn_stats(self, U, C, sums):
        self._chk_f32(sums)
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_stats", U.data_ptr(), rows, C, U.shape[2], sums.data_ptr(), self._udt(U), self._stream())

    def bn_finalize(self, sums, gamma, beta, rm, rv, count, eps, momentum, training, scale, shift, mean, rstd):
        self._chk_f32(sums, gamma, beta, rm, rv, scale, shift, mean, rstd)
        L.call("megb200_bn_finalize", _ptr(sums), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
               int(count), float(eps), float(momentum), int(training), scale.data_ptr(), shift.data_ptr(),
               mean.data_ptr(), rstd.data_ptr(), gamma.numel(), self._stream())

    def bn_gelu(self, U, scale, shift, H, C):
        self._chk_act(H)
        assert H.shape == U.shape
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_gelu", U.data_ptr(), scale.data_ptr(), shift.data_ptr(), H.data_ptr(), rows, C, U.shape[2],
               self._udt(U), self.dt, self._stream())

    def bn_gelu_train(self, U, sums, gamma, beta, rm, rv, count, eps, momentum, scale, shift, mean, rstd, H, C):
        """Training-mode BatchNorm + GELU in one launch (bn_finalize + bn_gelu; also updates rm / rv)."""
        self._chk_f32(sums, gamma, beta, rm, rv, scale, shift, mean, rstd)
        self._chk_act(H)
        assert H.shape == U.shape
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_gelu_train", U.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
               rv.data_ptr(), int(count), float(eps), float(momentum), scale.data_ptr(), shift.data_ptr(),
               mean.data_ptr(), rstd.data_ptr(), H.data_ptr(), rows, C, U.shape[2], self._udt(U), self.dt, self._stream())

    def bn_gelu_bwd_reduce(self, dH, U, scale, shift, mean, rstd, dY, sums2, C):
        """dY may be None: only the two per-channel sums are formed (bn_gelu_bwd_apply recomputes dY)."""
        self._chk_act(dH, dY)
        self._chk_f32(sums2)
        assert dH.shape == U.shape and (dY is None or dY.shape == U.shape)
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_gelu_bwd_reduce", dH.data_ptr(), U.data_ptr(), scale.data_ptr(), shift.data_ptr(),
               mean.data_ptr(), rstd.data_ptr(), _ptr(dY), sums2.data_ptr(), rows, C, U.shape[2], self._udt(U),
               self.dt, self._stream())

    def bn_gelu_bwd_apply(self, dH, U, scale, shift, gamma, mean, rstd, sums2, count, dU, C, dbias=None):
        """bn_bwd_apply with dY = dH * gelu'(scale*U + shift) recomputed instead of read."""
        self._chk_act(dH, dU)
        self._chk_f32(dbias)
        assert dH.shape == U.shape == dU.shape
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_gelu_bwd_apply", dH.data_ptr(), U.data_ptr(), scale.data_ptr(), shift.data_ptr(),
               gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums2.data_ptr(), int(count), dU.data_ptr(),
               _ptr(dbias), rows, C, U.shape[2], self._udt(U), self.dt, self._stream())

    def bn_bwd_apply(self, dY, U, gamma, mean, rstd, sums2, count, dU, C, dbias=None):
        """dbias (f32 [C], accumulated): column sums of dU = gradient of the conv bias in front of the BatchNorm."""
        self._chk_act(dY, dU)
        self._chk_f32(dbias)
        assert dY.shape == U.shape == dU.shape
        rows = U.shape[0] * U.shape[1]
        L.call("megb200_bn_bwd_apply", dY.data_ptr(), U.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
               rstd.data_ptr(), sums2.data_ptr(), int(count), dU.data_ptr(), _ptr(dbias), rows, C, U.shape[2],
               self._udt(U), self.dt, self._stream())

    def glu(self, V, Out, C):
        self._chk_act(V, Out)
        rows = V.shape[0] * V.shape[1]
        L.call("megb200_glu", V.data_ptr(), Out.data_ptr(), rows, C, V.shape[2], Out.shape[2], self.dt, self._stream())

    def glu_bwd(self, dOut, V, dV, C, dbias=None):
        """dbias (f32 [2C], accumulated): column sums of dV (needs C % 8 == 0)."""
        self._chk_act(dOut, V, dV)
        self._chk_f32(dbias)
        rows = V.shape[0] * V.shape[1]
        L.call("megb200_glu_bwd", dOut.data_ptr(), V.data_ptr(), dV.data_ptr(), _ptr(dbias), rows, C, V.shape[2],
               dOut.shape[2], self.dt, self._stream())

    def gelu(self, P, Q):
        self._chk_act(P, Q)
        L.call("megb200_gelu", P.data_ptr(), Q.data_ptr(), P.numel(), self.dt, self._stream())

    def gelu_bwd(self, dQ, P, dP, C, dbias=None):
        """dP = dQ * gelu'(P) on [B,T,Cs]; dbias (f32 [C], accumulated): column sums of dP."""
        self._chk_act(dQ, P, dP)
        self._chk_f32(dbias)
        assert dQ.shape == P.shape == dP.shape
        rows = P.shape[0] * P.shape[1]
        L.call("megb200_gelu_bwd", dQ.data_ptr(), P.data_ptr(), dP.data_ptr(), _ptr(dbias), rows, C, P.shape[2],
               self.dt, self._stream())

    def time_mean(self, Z, Zm, F):
        self._chk_act(Z, Zm)
        B, T, Fs = Z.shape
        L.call("megb200_time_mean", Z.data_ptr(), Zm.data_ptr(), B, T, F, Fs, self.dt, self._stream())

    def time_mean_bwd(self, dZm, dZ, F):
        self._chk_act(dZm, dZ)
        B, T, Fs = dZ.shape
        L.call("megb200_time_mean_bwd", dZm.data_ptr(), dZ.data_ptr(), B, T, F, Fs, self.dt, self._stream())

    def avgpool(self, X, Out, C, k=3, stride=2):
        """nn.AvgPool1d(k, stride) over time: X [B,T,Cs] -> Out [B,(T-k)//stride+1,Cs]."""
        self._chk_act(X, Out)
        B, T, Cs = X.shape
        assert Out.shape == (B, (T - k) // stride + 1, Cs)
        L.call("megb200_avgpool1d", X.data_ptr(), Out.data_ptr(), B, T, C, Cs, k, stride, self.dt, self._stream())

    def avgpool_bwd(self, dOut, dX, C, k=3, stride=2):
        self._chk_act(dOut, dX)
        B, T, Cs = dX.shape
        assert dOut.shape == (B, (T - k) // stride + 1, Cs)
        L.call("megb200_avgpool1d_bwd", dOut.data_ptr(), dX.data_ptr(), B, T, C, Cs, k, stride, self.dt, self._stream())

    # ------------------------------------------------------------------ CLIP
    def clip_prep_y(self, Y, Yp, ysq):
        """Y f32 | bf16 [B,F,T] or [B,F] -> Yp [B, T*F] (position-major) and ysq[b] += |Y_b|^2."""
        assert Y.is_cuda and Y.is_contiguous() and Y.dtype in (torch.float32, torch.bfloat16), (Y.dtype, Y.shape)
        self._chk_f32(ysq)
        self._chk_act(Yp)
        B, F = Y.shape[0], Y.shape[1]
        T = Y.shape[2] if Y.dim() == 3 else 1
        assert Yp.numel() == B * F * T
        L.call("megb200_clip_prep_y", Y.data_ptr(), L.F32 if Y.dtype == torch.float32 else L.BF16, Yp.data_ptr(),
               ysq.data_ptr(), B, F, T, self.dt, self._stream())

    def rowsumsq(self, Z, sq):
        self._chk_act(Z)
        self._chk_f32(sq)
        B = Z.shape[0]
        L.call("megb200_rowsumsq", Z.data_ptr(), sq.data_ptr(), B, Z.numel() // B, self.dt, self._stream())

    def clip_logits(self, Yp, Zp, raw):
        self._chk_act(Yp, Zp)
        self._chk_f32(raw)
        By, Bz = raw.shape
        L.call("megb200_clip_logits", Yp.data_ptr(), Zp.data_ptr(), raw.data_ptr(), By, Bz, Yp.numel() // By, self.dt,
               self.be, self._stream())

    def clip_ce_partial(self, raw, ysq, zsq, temp, Lg, row_max, row_sumexp, col_lse):
        self._chk_f32(raw, ysq, zsq, temp, Lg, row_max, row_sumexp, col_lse)
        By, Bz = raw.shape
        L.call("megb200_clip_ce_partial", raw.data_ptr(), ysq.data_ptr(), zsq.data_ptr(), temp.data_ptr(),
               Lg.data_ptr(), row_max.data_ptr(), row_sumexp.data_ptr(), col_lse.data_ptr(), By, Bz, self._stream())

    def clip_lse_merge(self, parts, row_lse):
        """parts f32 [W, 2, By] ((max, sum exp) pairs of the W ranks) -> row_lse [By]."""
        self._chk_f32(parts, row_lse)
        W, two, By = parts.shape
        assert two == 2 and row_lse.numel() == By
        L.call("megb200_clip_lse_merge", parts.data_ptr(), W, By, row_lse.data_ptr(), self._stream())

    # ---- fused large-batch form (short embeddings): the B x B logits are never written (bf16 / tensor cores only)
    @property
    def has_fused_clip(self):
        return self.be == L.TC

    def clip_fused_stats(self, Yp, Zp, ysq, zsq, temp, off, row_part, col_part, diag):
        """row_part f32 [ceil(Bz/128), 2, By], col_part f32 [ceil(By/128), 2, Bz], diag f32 [Bz]."""
        self._chk_act(Yp, Zp)
        self._chk_f32(ysq, zsq, temp, row_part, col_part, diag)
        By, Bz = Yp.shape[0], Zp.shape[0]
        assert row_part.shape == ((Bz + 127) // 128, 2, By) and col_part.shape == ((By + 127) // 128, 2, Bz)
        L.call("megb200_clip_fused_stats", Yp.data_ptr(), Zp.data_ptr(), ysq.data_ptr(), zsq.data_ptr(), temp.data_ptr(),
               By, Bz, Yp.numel() // By, int(off), row_part.data_ptr(), col_part.data_ptr(), diag.data_ptr(), self.dt,
               self.be, self._stream())

    def clip_fused_loss(self, row_lse, col_lse, diag, off, wscale, out2):
        self._chk_f32(row_lse, col_lse, diag, out2)
        L.call("megb200_clip_fused_loss", row_lse.data_ptr(), col_lse.data_ptr(), diag.data_ptr(), int(off),
               col_lse.numel(), float(wscale), out2.data_ptr(), self._stream())

    def clip_fused_coef(self, Yp, Zp, ysq, zsq, temp, row_lse, col_lse, off, wscale, j_begin, coef, cvec, out2):
        """coef [j_count, ldc] (activation dtype) for brain columns [j_begin, j_begin + j_count); cvec / out2[1] accumulated."""
        self._chk_act(Yp, Zp, coef)
        self._chk_f32(ysq, zsq, temp, row_lse, col_lse, cvec, out2)
        By, Bz = Yp.shape[0], Zp.shape[0]
        L.call("megb200_clip_fused_coef", Yp.data_ptr(), Zp.data_ptr(), ysq.data_ptr(), zsq.data_ptr(), temp.data_ptr(),
               row_lse.data_ptr(), col_lse.data_ptr(), By, Bz, Yp.numel() // By, int(off), float(wscale), int(j_begin),
               coef.shape[0], coef.data_ptr(), coef.shape[1], cvec.data_ptr(), out2.data_ptr(), self.dt, self.be,
               self._stream())

    def clip_ce_finish(self, Lg, row_lse, col_lse, ysq, zsq, temp, off, wscale, out2, coef, cvec):
        self._chk_f32(Lg, row_lse, col_lse, ysq, zsq, temp, out2, cvec)
        self._chk_act(coef)
        By, Bz = Lg.shape
        assert coef.shape[0] == Bz and coef.shape[1] >= By
        L.call("megb200_clip_ce_finish", Lg.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(), ysq.data_ptr(),
               zsq.data_ptr(), temp.data_ptr(), int(off), float(wscale), out2.data_ptr(), coef.data_ptr(),
               coef.shape[1], cvec.data_ptr(), By, Bz, self.dt, self._stream())

    def clip_dz(self, coef, Yp, Z, cvec, gscale, dZ):
        """coef: [Bz, ldc] (ldc = By rounded up to 8); Yp: [By, D]; Z, dZ: [Bz, D]."""
        self._chk_act(coef, Yp, Z, dZ)
        self._chk_f32(cvec, gscale)
        Bz, By = coef.shape[0], Yp.shape[0]
        L.call("megb200_clip_dz", coef.data_ptr(), coef.shape[1], Yp.data_ptr(), Z.data_ptr(), cvec.data_ptr(), gscale.data_ptr(),
               dZ.data_ptr(), By, Bz, Z.numel() // Bz, self.dt, self.be, self._stream())

    def retrieval_rank(self, raw, zsq, rank, off=0):
        self._chk_f32(raw, zsq)
        assert rank.dtype == torch.int32 and rank.is_cuda
        By, Bz = raw.shape
        L.call("megb200_retrieval_rank", raw.data_ptr(), zsq.data_ptr(), rank.data_ptr(), By, Bz, int(off),
               self._stream())

    def cast(self, src, dst):
        code = {torch.float32: L.F32, torch.bfloat16: L.BF16}
        assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
        L.call("megb200_cast", src.data_ptr(), dst.data_ptr(), src.numel(), code[src.dtype], code[dst.dtype],
               self._stream())
