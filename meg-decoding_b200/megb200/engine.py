"""Step orchestration of the hot path on top of the op interface (ops.CudaOps): BrainEncoder forward/backward
and CLIPLoss forward/backward as explicit kernel sequences (no autograd inside, no tracing compiler).

Math follows SURVEY appendix A (verified against the reference): reference op sites are cited inline
(paths relative to the reference tree).  Design points:

* The linear chain spatial-attention mix -> shared 1x1 conv -> per-subject 1x1 conv (models.py:220,269-272) is
  pre-multiplied in fp32 per step into one effective (D1 x C) matrix and bias per subject and executed as ONE grouped
  GEMM keyed by the subject index; the spatial-dropout mask (models.py:235-239) is a column mask of that matrix.
  X needs no gradient, so the backward of the whole block is one grouped weight-gradient GEMM plus small fp32 algebra.
* Activations live in the position-major "PC" layout [B][T][Cs]; one rounding to the storage dtype per stored tensor;
  BN statistics, residual adds and GELU/GLU are evaluated in fp32.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from .ops import pad8, pad16, pad64


def _f32(dev, *shape):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


def _zeros(dev, *shape):
    return torch.zeros(*shape, dtype=torch.float32, device=dev)


def validate_subject_idxs(idx: torch.Tensor, S: int) -> torch.Tensor:
    """Range-check (IndexError like ``nn.ModuleList.__getitem__``) and wrap negative subject indices; CPU or CUDA."""
    if idx.numel() == 0:
        return idx
    if idx.dtype not in (torch.int32, torch.int64, torch.int16, torch.int8, torch.uint8):
        raise TypeError("subject_idxs must be an integer tensor, got {}".format(idx.dtype))
    lo, hi = torch.aminmax(idx)
    lo, hi = int(lo), int(hi)                      # a device tensor: one synchronising read of two scalars
    if hi >= S or lo < -S:
        raise IndexError("index {} is out of range".format(hi if hi >= S else lo))
    return torch.where(idx < 0, idx + S, idx) if lo < 0 else idx


class _ZeroPool:
    """One memset per pass instead of ~60 tiny ones: f32 accumulators / gradient buffers are carved out of a single
    zero-filled allocation (256-byte aligned views).  Falls back to torch.zeros when exhausted."""

    def __init__(self, dev, nfloats):
        self.dev = dev
        self.buf = torch.zeros(int(nfloats), dtype=torch.float32, device=dev)
        self.off = 0
        self.spilled = False       # a take() did not fit: that tensor is NOT part of the flat buffer

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_al = (n + 63) // 64 * 64
        if self.off + n_al > self.buf.numel():
            self.spilled = True
            return torch.zeros(*shape, dtype=torch.float32, device=self.dev)
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += n_al
        return v


class _SideStream:
    """Fork / join of a second CUDA stream for the small fp32 launches that do not depend on each other (the per-step
    algebra of the pre-multiplied subject block is ~25 latency-bound launches, 0.43 ms when issued back to back).  All
    tensors are allocated by the caller on the main stream BEFORE the side block; the side block only launches kernels
    and the main stream joins before anything is freed or consumed, so the caching allocator needs no extra care.  Under
    CUDA-graph capture the fork/join becomes two parallel branches of the graph.  No-op on CPU (op restatement).

    Opt-in (MEGB200_SIDE_STREAM=1).  Measured at config 2: 7.242 vs 7.257 ms per replayed step -- the sequential chain, not
    the launches beside it, is the critical path -- while eager launches get SLOWER (8.9 vs 7.9 ms: the stream switches
    cost host time in a host-bound loop), so the default is one stream."""

    _streams = {}

    def __init__(self, dev):
        self.on = dev.type == "cuda" and os.environ.get("MEGB200_SIDE_STREAM", "0") == "1"
        self.dev = dev
        if self.on:
            key = (dev.index if dev.index is not None else torch.cuda.current_device())
            if key not in _SideStream._streams:
                _SideStream._streams[key] = torch.cuda.Stream(device=dev)
            self.side = _SideStream._streams[key]

    def __enter__(self):
        if self.on:
            self.side.wait_stream(torch.cuda.current_stream(self.dev))
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self._ctx.__exit__(*exc)
        return False

    def join(self):
        if self.on:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)


class EncoderEngine:
    """BrainEncoder.forward / backward (models.py:367-383 and the modules it owns)."""

    def __init__(self, ops, *, C: int, D1: int, D2: int, F: int, S: int, seq2seq: bool, dilations=None,
                 eps: float = 1e-5, momentum: float = 0.1, ks=None, pooled: bool = False):
        """``ks`` (five odd kernel sizes) and ``pooled`` select BrainEncoderSeq2Static (models.py:465-512): block k convolves
        with kernel size ks[k], AvgPool1d(3, 2) follows blocks 0..3 and a global average over time block 4, so the 1x1 head
        runs on one position per sample.  ``dilations`` (5 x 3) is the paper's dilated block
        (tests/test_modules/models.py:275-297); the reference code as it runs has 1 everywhere."""
        self.ops, self.C, self.D1, self.D2, self.F, self.S = ops, C, D1, D2, F, S
        self.seq2seq = seq2seq
        self.dil = [tuple(d) for d in dilations] if dilations is not None else [(1, 1, 1)] * 5
        self.ks = [int(k) for k in ks] if ks is not None else [3] * 5
        if any(k < 1 or k % 2 == 0 for k in self.ks):
            raise ValueError("megb200: Conv1d kernel sizes must be odd (padding='same' is symmetric only then)")
        self.pooled = bool(pooled)
        self.eps, self.momentum = eps, momentum
        self._packed = {}        # persistent packed-weight buffers, keyed by (name, mode, device)
        self._zp = None          # zero pool of the pass in flight

    # ------------------------------------------------------------------------------------------ weight packing
    def _conv_list(self, P):
        """(name, weight, N, cin, taps) of every convolution on the path, in execution order."""
        D1, D2, Fo = self.D1, self.D2, self.F
        out = []
        for k in range(5):
            pk = P["blocks"][k]
            out.append((f"b{k}.w0", pk["w0"], D2, D1 if k == 0 else D2, self.ks[k]))
            out.append((f"b{k}.w1", pk["w1"], D2, D2, self.ks[k]))
            out.append((f"b{k}.w2", pk["w2"], 2 * D2, D2, self.ks[k]))
        out.append(("wf1", P["wf1"], 2 * D2, D2, 1))
        out.append(("wf2", P["wf2"], Fo, 2 * D2, 1))
        return out

    def _prepare_packs(self, P, mode, dev):
        """Allocate (once) the persistent packed-weight buffers of `mode` on the calling stream."""
        for name, w, N, cin, taps in self._conv_list(P):
            key = (name, mode, str(dev))
            buf = self._packed.get(key)
            shape = (1, pad16(N), taps * pad64(cin)) if mode == 0 else (1, pad16(cin), taps * pad64(N))
            if buf is None or tuple(buf.shape) != shape or buf.dtype != self.ops.act_dtype:
                self._packed[key] = self.ops.empty(*shape, device=dev)

    def _pack_all(self, P, mode, dev):
        """Pack every conv weight for the forward (mode 0) or the data-gradient (mode 1) tap-GEMMs in one launch."""
        items, bufs = [], {}
        for name, w, N, cin, taps in self._conv_list(P):
            key = (name, mode, str(dev))
            buf = self._packed.get(key)
            shape = (1, pad16(N), taps * pad64(cin)) if mode == 0 else (1, pad16(cin), taps * pad64(N))
            if buf is None or tuple(buf.shape) != shape or buf.dtype != self.ops.act_dtype:
                buf = self._packed[key] = self.ops.empty(*shape, device=dev)
            items.append((w.view(1, N, cin, taps), buf, mode))
            bufs[name] = buf
        self.ops.pack_weights(items)
        return bufs

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, P: Dict, X: torch.Tensor, subj: torch.Tensor, mask: Optional[torch.Tensor], training: bool):
        ops, dev = self.ops, X.device
        B, C, T = X.shape
        D1, D2, Fo, S = self.D1, self.D2, self.F, self.S
        assert C == self.C, f"expected {self.C} sensors, got {C}"                       # models.py:233
        ctx = {"B": B, "T": T, "subj": subj, "mask": mask}
        self._zp = _ZeroPool(dev, 10 * 2 * pad64(D2) + 1024)

        # ---- spatial attention weights + subject chain, fp32 (models.py:204-213, 252, 269-272).  The chain
        # a -> softmax -> W_c. -> W_s[s]. -> pack is sequential; the launches that do not depend on it (operand packs of
        # all conv weights and of X, the effective bias) run beside it on a second stream.
        a, Wt, Wm = _f32(dev, D1, C), _f32(dev, D1, C), _f32(dev, D1, C)
        M1, Weff, beff = _f32(dev, D1, C), _f32(dev, S, D1, C), _f32(dev, S, D1, 1)
        Wp = ops.empty(S, pad16(D1), pad64(C), device=dev)
        Xp = ops.empty(B, T, pad8(C), device=dev)
        self._prepare_packs(P, 0, dev)
        side = _SideStream(dev)
        with side:
            self._wf = self._pack_all(P, 0, dev)
            ops.pack_x(X, Xp)
            ops.matmul(P["Ws"], P["bc"].view(D1, 1), beff)
        ops.matmul(P["z_re"], P["cos"], a)
        ops.matmul(P["z_im"], P["sin"], a, beta=1.0)
        ops.sa_softmax(a, mask, Wt, Wm)
        ops.matmul(P["Wc"], Wm, M1)
        ops.matmul(P["Ws"], M1, Weff)
        ops.pack_weight(Weff.view(S, D1, C, 1), Wp, 0)
        side.join()
        H2 = ops.empty(B, T, pad8(D1), device=dev)
        ops.tapgemm(Xp, C, Wp, D1, 1, 1, beff.view(S, D1), None, subj, H2)
        ctx.update(Wt=Wt, Wm=Wm, M1=M1, Xp=Xp, H2=H2)

        # ---- five residual conv blocks (models.py:308-322)
        x, cin = H2, D1
        blocks = []
        for k in range(5):
            pk, d, kt = P["blocks"][k], self.dil[k], self.ks[k]
            bc = {"x": x, "cin": cin}
            # pre-BatchNorm tensors stay in f32: the normalisation sees the unrounded convolution (+ residual)
            # (the train-mode batch statistics ride on the conv epilogue: sums[2][D2])
            s0 = self._zp.take(2, D2) if training else None
            u0 = self._conv(x, cin, f"b{k}.w0", pk["b0"], D2, d[0], x if k > 0 else None, taps=kt, epi="store_f32", stats=s0)
            h0, bc["bn0"] = self._bn_gelu(u0, pk, "0", training, s0)
            s1 = self._zp.take(2, D2) if training else None
            u1 = self._conv(h0, D2, f"b{k}.w1", pk["b1"], D2, d[1], h0, taps=kt, epi="store_f32", stats=s1)
            h1, bc["bn1"] = self._bn_gelu(u1, pk, "1", training, s1)
            out, v = self._conv(h1, D2, f"b{k}.w2", pk["b2"], 2 * D2, d[2], None, taps=kt, epi="glu")
            bc.update(u0=u0, h0=h0, u1=u1, h1=h1, v=v)
            blocks.append(bc)
            if self.pooled:                                                             # models.py:483-487
                Tk = out.shape[1]
                if k < 4:
                    assert Tk >= 3, "sequence too short for AvgPool1d(3, 2)"
                    pooled = ops.empty(B, (Tk - 3) // 2 + 1, out.shape[2], device=dev)
                    ops.avgpool(out, pooled, D2, 3, 2)
                else:
                    pooled = ops.empty(B, 1, out.shape[2], device=dev)
                    ops.time_mean(out, pooled, D2)
                out = pooled
            x, cin = out, D2
        ctx["blocks"] = blocks
        ctx["x5"] = x

        # ---- 1x1 head (models.py:373-374): GELU after both convs.  After the global average there is one position per
        # sample: a 1x1 conv is position-wise, so the batch is laid out as the positions of ONE sample (full tiles)
        if self.pooled:
            x = x.view(1, B, x.shape[2])
            ctx["x5"] = x
        q1, p1 = self._conv(x, D2, "wf1", P["bf1"], 2 * D2, 1, None, taps=1, epi="gelu")
        Z, p2 = self._conv(q1, 2 * D2, "wf2", P["bf2"], Fo, 1, None, taps=1, epi="gelu")
        ctx.update(p1=p1, q1=q1, p2=p2)
        if self.pooled:
            return Z.view(B, 1, Z.shape[2]), ctx        # mean over the single time step (models.py:510) is the identity
        if not self.seq2seq:                                                            # models.py:382
            Zm = ops.empty(B, 1, pad8(Fo), device=dev)
            ops.time_mean(Z, Zm, Fo)
            return Zm, ctx
        return Z, ctx

    def _conv(self, x, cin, wname, b, N, dil, res, taps=3, epi="store", stats=None):
        """Conv1d (k = taps, "same") as a tap-GEMM with a fused tail on the pre-packed weight `wname`.  Returns Out, or
        (Out, Out2) for the GELU / GLU epilogues (Out2 = the pre-activation kept for backward)."""
        ops, dev = self.ops, x.device
        B, T = x.shape[0], x.shape[1]
        Wpk = self._wf[wname]
        if epi == "store_f32":
            out = _f32(dev, B, T, pad8(N))
            ops.tapgemm(x, cin, Wpk, N, taps, dil, b, res, None, out, epi=epi, stats=stats)
            return out
        if epi == "glu":
            out, v = ops.empty(B, T, pad8(N // 2), device=dev), ops.empty(B, T, pad8(N), device=dev)
            ops.tapgemm(x, cin, Wpk, N, taps, dil, b, None, None, out, epi=epi, Out2=v)
            return out, v
        out = ops.empty(B, T, pad8(N), device=dev)
        if epi == "gelu":
            pre = ops.empty(B, T, pad8(N), device=dev)
            if taps == 1 and res is None:
                # a 1x1 conv is position-wise: B*T positions as ONE sample fill every 128-row tile (T = 360 alone leaves
                # the third tile of each sample 81 % full)
                ops.tapgemm(x.view(1, B * T, x.shape[2]), cin, Wpk, N, 1, 1, b, None, None, out.view(1, B * T, out.shape[2]),
                            epi=epi, Out2=pre.view(1, B * T, pre.shape[2]))
            else:
                ops.tapgemm(x, cin, Wpk, N, taps, dil, b, res, None, out, epi=epi, Out2=pre)
            return out, pre
        ops.tapgemm(x, cin, Wpk, N, taps, dil, b, res, None, out)
        return out

    def _bn_gelu(self, u, pk, i, training, sums):
        ops, dev, D2 = self.ops, u.device, self.D2
        n = u.shape[0] * u.shape[1]
        st = {k: _f32(dev, D2) for k in ("scale", "shift", "mean", "rstd")}
        h = ops.empty(*u.shape, device=dev)
        if training:       # one launch: statistics -> scale/shift, running-stat update, normalise + GELU
            ops.bn_gelu_train(u, sums, pk["g" + i], pk["be" + i], pk["rm" + i], pk["rv" + i], n, self.eps,
                              self.momentum, st["scale"], st["shift"], st["mean"], st["rstd"], h, D2)
        else:
            ops.bn_finalize(sums, pk["g" + i], pk["be" + i], pk["rm" + i], pk["rv" + i], n, self.eps, self.momentum,
                            training, st["scale"], st["shift"], st["mean"], st["rstd"])
            ops.bn_gelu(u, st["scale"], st["shift"], h, D2)
        return h, st

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, P: Dict, ctx: Dict, dZ: torch.Tensor, sync=None):
        """dZ: gradient w.r.t. the returned embedding, PC layout ([B,T,Fs] or [B,1,Fs]).  Returns a dict of f32
        gradients keyed like P.

        Every parameter gradient is a view of ONE flat zero-initialised f32 buffer, laid out in the order the backward
        pass completes them: [head | block 4 .. block 1 | block 0 | subject block, spatial attention | 2 spare floats];
        ``G["_flat"]`` is that range, so data parallel all-reduces it in place (no gather / scatter copies; the caller may
        put the loss and the temperature gradient into ``G["_extra"]`` first).  ``sync`` (optional): ``sync.early(t)`` is
        called with the [head .. block 1] sub-range as soon as those gradients are final -- the caller starts their
        all-reduce on a side stream while block 0 and the subject block are still running -- and ``G["_flat_late"]`` is
        the rest."""
        ops, dev = self.ops, dZ.device
        B, T = ctx["B"], ctx["T"]
        D1, D2, Fo, S, C = self.D1, self.D2, self.F, self.S, self.C
        n = B * T
        G: Dict = {"blocks": [dict() for _ in range(5)]}
        self._wd = self._pack_all(P, 1, dev)
        nparam = sum(int(w.numel()) + N for _, w, N, _, _ in self._conv_list(P))
        ntail = S * D1 * D1 + D1 * D1 + D1 + 2 * int(P["z_re"].numel())
        self._zp = zp = _ZeroPool(dev, nparam + ntail + S * D1 * C + S * D1 + 20 * 2 * pad64(D2) + 64 * 96)

        if self.pooled:
            dZ = dZ.reshape(1, B, dZ.shape[-1])          # the head ran with the batch as positions (see forward)
        elif not self.seq2seq:
            full = ops.empty(B, T, pad8(Fo), device=dev)
            ops.time_mean_bwd(dZ, full, Fo)
            dZ = full
        # ---- head
        # bias gradients are column sums of the tensors the element-wise backward passes write: fused there
        dp2, G["bf2"] = ops.empty(*dZ.shape, device=dev), self._zp.take(Fo)
        ops.gelu_bwd(dZ, ctx["p2"], dp2, Fo, G["bf2"])
        G["wf2"] = self._wgrad(dp2, Fo, ctx["q1"], 2 * D2, 1, 1)
        dq1 = self._dgrad(dp2, Fo, "wf2", 2 * D2, 1, None, taps=1)
        dp1, G["bf1"] = ops.empty(*dq1.shape, device=dev), self._zp.take(2 * D2)
        ops.gelu_bwd(dq1, ctx["p1"], dp1, 2 * D2, G["bf1"])
        G["wf1"] = self._wgrad(dp1, 2 * D2, ctx["x5"], D2, 1, 1)
        dx = self._dgrad(dp1, 2 * D2, "wf1", D2, 1, None, taps=1)

        # ---- conv blocks, last to first
        early_end = 0
        for k in range(4, -1, -1):
            pk, bc, d, g, kt = P["blocks"][k], ctx["blocks"][k], self.dil[k], G["blocks"][k], self.ks[k]
            n = B * bc["v"].shape[1]                     # positions this block's BatchNorms averaged over
            if self.pooled:                              # undo the pooling that followed this block (models.py:483-487)
                full = ops.empty(B, bc["v"].shape[1], dx.shape[-1], device=dev)
                if k == 4:
                    ops.time_mean_bwd(dx.reshape(B, 1, dx.shape[-1]), full, D2)
                else:
                    ops.avgpool_bwd(dx, full, D2, 3, 2)
                dx = full
            dv = ops.empty(*bc["v"].shape, device=dev)
            g["b2"] = self._zp.take(2 * D2)
            if D2 % 8 == 0 and bc["v"].shape[2] == 2 * D2:
                ops.glu_bwd(dx, bc["v"], dv, D2, g["b2"])
            else:
                ops.glu_bwd(dx, bc["v"], dv, D2)
                ops.colsum(dv, 2 * D2, None, g["b2"])
            g["w2"] = self._wgrad(dv, 2 * D2, bc["h1"], D2, kt, d[2])
            dh1 = self._dgrad(dv, 2 * D2, f"b{k}.w2", D2, d[2], None, taps=kt)
            du1, g["g1"], g["be1"], g["b1"] = self._bn_gelu_bwd(dh1, bc["u1"], bc["bn1"], pk["g1"], n)
            g["w1"] = self._wgrad(du1, D2, bc["h0"], D2, kt, d[1])
            dh0 = self._dgrad(du1, D2, f"b{k}.w1", D2, d[1], du1, taps=kt)
            du0, g["g0"], g["be0"], g["b0"] = self._bn_gelu_bwd(dh0, bc["u0"], bc["bn0"], pk["g0"], n)
            g["w0"] = self._wgrad(du0, D2, bc["x"], bc["cin"], kt, d[0])
            dx = self._dgrad(du0, D2, f"b{k}.w0", bc["cin"], d[0], du0 if k > 0 else None, taps=kt)
            if k == 1:
                early_end = zp.off
                if sync is not None and not zp.spilled:
                    sync.early(zp.buf[:early_end])
        dH2 = dx

        # ---- subject block + spatial attention, fp32 small algebra
        subj = ctx["subj"]
        order = torch.empty_like(subj)
        ops.group_order(subj, S, order)
        dWs, dWc, dbc = zp.take(S, D1, D1), zp.take(D1, D1), zp.take(D1, 1)
        dzr, dzi = zp.take(*P["z_re"].shape), zp.take(*P["z_im"].shape)
        G["_extra"] = zp.take(2)              # spare floats inside the flat range (loss, temperature gradient)
        grad_end = zp.off
        dWeff = self._zp.take(S, D1, C, 1)
        ops.wgrad(dH2, D1, ctx["Xp"], C, 1, 1, subj, order, dWeff)
        dWeff = dWeff.view(S, D1, C)
        dbeff = self._zp.take(S, D1)
        ops.colsum(dH2, D1, subj, dbeff)
        # two independent strands of small fp32 GEMMs: the subject-weight / bias gradients on a second stream, the chain
        # dM1 -> dWm -> softmax Jacobian -> dz on the main one
        Ws2d = P["Ws"].view(S * D1, D1)
        dM1, dWm, da = _f32(dev, D1, C), _f32(dev, D1, C), _f32(dev, D1, C)
        side = _SideStream(dev)
        with side:
            ops.matmul(dWeff, ctx["M1"], dWs, transB=True)
            ops.matmul(dbeff.view(S, D1, 1), P["bc"].view(1, D1), dWs, beta=1.0)
            ops.matmul(Ws2d, dbeff.view(S * D1, 1), dbc, transA=True)
        ops.matmul(Ws2d, dWeff.view(S * D1, C), dM1, transA=True)
        ops.matmul(P["Wc"], dM1, dWm, transA=True)
        ops.sa_softmax_bwd(ctx["Wt"], dWm, ctx["mask"], da)
        ops.matmul(da, P["cos"], dzr, transB=True)
        ops.matmul(da, P["sin"], dzi, transB=True)
        ops.matmul(dM1, ctx["Wm"], dWc, transB=True)
        side.join()
        G.update(Ws=dWs, Wc=dWc, bc=dbc.view(D1), z_re=dzr, z_im=dzi)
        if not zp.spilled:
            G["_flat"] = zp.buf[:grad_end]
            G["_flat_late"] = zp.buf[early_end:grad_end] if (sync is not None and early_end) else G["_flat"]
        return G

    def _wgrad(self, dU, M, Xin, C, taps, dil):
        ops, dev = self.ops, dU.device
        dW = self._zp.take(1, M, C, taps)
        ops.wgrad(dU, M, Xin, C, taps, dil, None, None, dW)
        return dW.view(M, C, taps)

    def _dgrad(self, dU, M, wname, cin, dil, res, taps=3):
        """dX = conv_transpose(dU): same tap-GEMM with the transposed, tap-flipped weight pack `wname`."""
        ops = self.ops
        Wd = self._wd[wname]
        B, T = dU.shape[0], dU.shape[1]
        out = ops.empty(B, T, pad8(cin), device=dU.device)
        if taps == 1 and res is None:          # position-wise: flatten (see _conv)
            ops.tapgemm(dU.view(1, B * T, dU.shape[2]), M, Wd, cin, 1, 1, None, None, None, out.view(1, B * T, out.shape[2]))
        else:
            ops.tapgemm(dU, M, Wd, cin, taps, dil, None, res, None, out)
        return out

    def _bn_gelu_bwd(self, dh, u, st, gamma, n):
        ops, dev, D2 = self.ops, dh.device, self.D2
        # Two passes with dY materialised.  The recompute form (bn_gelu_bwd_reduce without dY + bn_gelu_bwd_apply) moves
        # 59 MB less per BatchNorm but measured slower (47 + 52 us vs 48 + 46 us at config 2): the reduce pass is bound by
        # the GELU-derivative arithmetic, not by its store, and the apply pass then pays for that arithmetic a second time.
        dy = ops.empty(*u.shape, device=dev)
        sums2 = self._zp.take(2, D2)
        ops.bn_gelu_bwd_reduce(dh, u, st["scale"], st["shift"], st["mean"], st["rstd"], dy, sums2, D2)
        du = ops.empty(*u.shape, device=dev)
        db = self._zp.take(D2)
        ops.bn_bwd_apply(dy, u, gamma, st["mean"], st["rstd"], sums2, n, du, D2, db)
        return du, sums2[1], sums2[0], db


class ClipEngine:
    """CLIPLoss.forward / backward (loss.py:66-112), one-sided data-parallel form of SURVEY appendix A6.

    Rows of the logits are the speech targets (x = Y), columns the brain embeddings (y = Z), exactly as
    ``train.py:203`` calls ``loss_func(Y, Z)``.  With ``world > 1`` every rank gathers the (constant) normalised
    targets, contracts them against its own brain shard, and the per-row log-sum-exp over all brains is completed
    by a MAX / SUM all-reduce of B_glob floats; no gradient crosses ranks except through the parameter all-reduce.
    """

    def __init__(self, ops):
        self.ops = ops

    def prepare_targets(self, Y, dist_ctx=None):
        """Pack the (constant) targets: f32 [B,F,T] | [B,F] -> position-major rows in the activation dtype plus their
        squared norms, all-gathered over the ranks in data parallel.  Independent of the encoder, so callers may run
        it on a side stream while the encoder forward executes (CLIPLoss.prefetch_targets).

        Data parallel on CUDA: the pack kernel writes this rank's slot of a peer-mapped slab and the other ranks' slots
        are pulled by the copy engines over NVLink (parallel.PeerGather) -- no SM is used, so the gather overlaps the
        encoder forward without slowing its persistent kernels.  The returned tensors are views of that slab: they stay
        valid until the next prepare_targets call with the same shape."""
        ops, dev = self.ops, Y.device
        B = Y.shape[0]
        D = Y.numel() // B
        packed = hasattr(Y, "rows")            # megb200.data.PackedTargets: already position-major in the activation dtype
        if packed and (Y.rows.dtype != ops.act_dtype or not Y.rows.is_contiguous() or Y.sumsq.dtype != torch.float32):
            raise RuntimeError(f"PackedTargets must hold contiguous {ops.act_dtype} rows and f32 norms for this precision")
        if dist_ctx is not None and dist_ctx.world > 1:
            pair = dist_ctx.target_gather(B, D, ops.act_dtype, dev)
            if pair is not None:
                gy, gs = pair
                Yp, ysq = gy.slot().view(B, D), gs.slot()
                if packed:
                    Yp.copy_(Y.rows)
                    ysq.copy_(Y.sumsq)
                else:
                    ysq.zero_()
                    ops.clip_prep_y(Y, Yp, ysq)
                gy.fence()
                gy.pull()
                gs.pull()
                gy.fence()
                return gy.buf.view(dist_ctx.world * B, D), gs.buf.view(dist_ctx.world * B)
        if packed:
            Yp, ysq = Y.rows, Y.sumsq
        else:
            Yp = ops.empty(B, D, device=dev)
            ysq = _zeros(dev, B)
            ops.clip_prep_y(Y, Yp, ysq)
        if dist_ctx is not None and dist_ctx.world > 1:
            Yp, ysq = dist_ctx.all_gather_rows(Yp), dist_ctx.all_gather_rows(ysq)
        return Yp, ysq

    def forward(self, Y, Zi, temp, reduction="mean", dist_ctx=None, want_logits=False, prepared=None,
                reduce_loss=True):
        """Y: f32 targets [B,F,T] | [B,F]; Zi: embeddings in PC layout, contiguous [B, T, F] | [B, 1, F].
        reduce_loss=False (data parallel): return this rank's loss CONTRIBUTION; the caller sums it over the ranks
        (GraphedTrainStep lets it ride in the gradient all-reduce)."""
        ops, dev = self.ops, Zi.device
        B = Zi.shape[0]
        D = Zi.numel() // B
        assert Y.numel() == B * D, f"targets {tuple(Y.shape)} and embeddings {tuple(Zi.shape)} differ in size"
        Yp, ysq = prepared if prepared is not None else self.prepare_targets(Y, dist_ctx)
        zsq = _zeros(dev, B)
        Zf = Zi.view(B, D)
        ops.rowsumsq(Zf, zsq)
        world, rank = (dist_ctx.world, dist_ctx.rank) if dist_ctx is not None else (1, 0)
        Bg, off = B * world, rank * B
        if reduction == "mean":
            wscale = 0.5 / Bg
        elif reduction == "sum":
            wscale = 0.5
        else:
            raise ValueError(f"unsupported reduction {reduction!r}")
        if not want_logits and self._use_fused(Bg, B, D):
            return self._forward_fused(Yp, ysq, Zf, zsq, temp, wscale, world, off, dist_ctx, reduce_loss, Zi.shape)
        raw = _zeros(dev, Bg, B)
        ops.clip_logits(Yp, Zf, raw)
        Lg = _f32(dev, Bg, B)
        parts, col_lse = _f32(dev, 1, 2, Bg), _f32(dev, B)         # per speech row: (max, sum exp) over the LOCAL brains
        ops.clip_ce_partial(raw, ysq, zsq, temp, Lg, parts[0, 0], parts[0, 1], col_lse)
        if world > 1:
            parts = dist_ctx.all_gather_rows(parts)                 # [world, 2, Bg]: one collective, merged on the device
        row_lse = _f32(dev, Bg)
        ops.clip_lse_merge(parts, row_lse)
        out2 = _zeros(dev, 2)
        coef = ops.empty(B, pad8(Bg), device=dev)
        cvec = _zeros(dev, B)
        ops.clip_ce_finish(Lg, row_lse, col_lse, ysq, zsq, temp, off, wscale, out2, coef, cvec)
        loss = out2[0:1].clone()
        if world > 1 and reduce_loss:
            loss = dist_ctx.all_reduce_sum(loss)
        ctx = dict(coef=coef, Yp=Yp, Zf=Zf, cvec=cvec, dtemp=out2[1:2], shape=Zi.shape)
        return loss.view(()), (Lg if want_logits else None), ctx

    # ------------------------------------------------------------------ large batch, short embeddings
    FUSED_MAX_D = 4096                 # the whole contraction runs inside one CTA's k loop
    FUSED_MIN_LOGITS = 1 << 22         # auto: from 2048 x 2048 logits up (below that three small matrices are cheaper)

    def _use_fused(self, Bg, B, D):
        """The never-materialised form (loss.py:86-107 at the GOD sweep's batch sizes)?  Policy option
        "clip_fused": -1 auto, 0 never, 1 whenever the engine supports it."""
        from . import lib
        if not getattr(self.ops, "has_fused_clip", False) or D > self.FUSED_MAX_D or D % 8 != 0:
            return False
        mode = lib.get_option("clip_fused")
        return mode == 1 or (mode != 0 and Bg * B >= self.FUSED_MIN_LOGITS)

    def _forward_fused(self, Yp, ysq, Zf, zsq, temp, wscale, world, off, dist_ctx, reduce_loss, shape):
        """Logits tiles live in TMEM only: per-tile online (max, sum exp) partials for rows and columns + the positives;
        two merges give the log-sum-exps; the backward pass recomputes the tiles (FlashAttention-style)."""
        ops, dev = self.ops, Zf.device
        B, Bg = Zf.shape[0], Yp.shape[0]
        row_part = _f32(dev, (B + 127) // 128, 2, Bg)
        col_part = _f32(dev, (Bg + 127) // 128, 2, B)
        diag = _zeros(dev, B)
        ops.clip_fused_stats(Yp, Zf, ysq, zsq, temp, off, row_part, col_part, diag)
        if world > 1:
            row_part = dist_ctx.all_gather_rows(row_part)           # the other ranks' brain columns
        row_lse, col_lse = _f32(dev, Bg), _f32(dev, B)
        ops.clip_lse_merge(row_part, row_lse)
        ops.clip_lse_merge(col_part, col_lse)
        out2 = _zeros(dev, 2)
        ops.clip_fused_loss(row_lse, col_lse, diag, off, wscale, out2)
        loss = out2[0:1].clone()
        if world > 1 and reduce_loss:
            loss = dist_ctx.all_reduce_sum(loss)
        ctx = dict(fused=True, Yp=Yp, Zf=Zf, ysq=ysq, zsq=zsq, temp=temp, row_lse=row_lse, col_lse=col_lse, off=off,
                   wscale=wscale, out2=out2, shape=shape)
        return loss.view(()), None, ctx

    def _backward_fused(self, ctx, gscale):
        ops = self.ops
        Yp, Zf = ctx["Yp"], ctx["Zf"]
        B, Bg, dev = Zf.shape[0], Yp.shape[0], Zf.device
        ldc = pad8(Bg)
        # brain columns in chunks of <= 64 M coefficients (128 MB of bf16): what exists of dL at any time
        chunk = max(128, min((B + 127) // 128 * 128, ((64 << 20) // ldc) // 128 * 128))
        coef = ops.empty(min(chunk, B), ldc, device=dev)
        cvec = _zeros(dev, B)
        dZ = ops.empty(*Zf.shape, device=dev)
        for j0 in range(0, B, chunk):
            n = min(chunk, B - j0)
            ops.clip_fused_coef(Yp, Zf, ctx["ysq"], ctx["zsq"], ctx["temp"], ctx["row_lse"], ctx["col_lse"], ctx["off"],
                                ctx["wscale"], j0, coef[:n], cvec, ctx["out2"])
            ops.clip_dz(coef[:n], Yp, Zf[j0:j0 + n], cvec[j0:j0 + n], gscale, dZ[j0:j0 + n])
        return dZ.view(ctx["shape"]), ctx["out2"][1:2] * gscale

    def retrieval_ranks(self, Y, Zi):
        """Rank of the matching brain embedding among all brains of the batch for every speech segment (cosine
        similarity; the reference's Classifier, models.py:403-460).  Returns int32 [B]."""
        ops, dev = self.ops, Zi.device
        B = Zi.shape[0]
        D = Zi.numel() // B
        Yp, _ = self.prepare_targets(Y, None)
        zsq = _zeros(dev, B)
        Zf = Zi.view(B, D)
        ops.rowsumsq(Zf, zsq)
        raw = _zeros(dev, B, B)
        ops.clip_logits(Yp, Zf, raw)
        rank = torch.empty(B, dtype=torch.int32, device=dev)
        ops.retrieval_rank(raw, zsq, rank, 0)
        return rank

    def backward(self, ctx, gscale):
        """gscale: f32 [1] upstream gradient of the (scalar) loss.  Returns (dZ in PC layout, dtemp f32 [1])."""
        if ctx.get("fused"):
            return self._backward_fused(ctx, gscale)
        ops = self.ops
        dZ = ops.empty(*ctx["Zf"].shape, device=gscale.device)
        ops.clip_dz(ctx["coef"], ctx["Yp"], ctx["Zf"], ctx["cvec"], gscale, dZ)
        return dZ.view(ctx["shape"]), ctx["dtemp"] * gscale
