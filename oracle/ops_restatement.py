"""ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-torch restatement of every op of the C-ABI (include/megb200.h), with
the same Python interface as ``megb200.ops.CudaOps``.

Two uses, both in tests only: (1) each CUDA kernel is compared against its restatement here on identical inputs;
(2) ``megb200.engine`` (the step orchestration) is run on CPU on top of this class and compared with autograd of
``oracle/restatement.py`` — that checks the hand-derived backward chain without a GPU.  The product never imports it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gelu_grad(x):
    """d gelu / dx by autograd of torch's own erf-GELU (independent of the closed form the kernels use)."""
    with torch.enable_grad():
        y = x.detach().clone().requires_grad_(True)
        g, = torch.autograd.grad(F.gelu(y).sum(), y)
    return g


class TorchOps:
    backend = "oracle"

    def __init__(self, dtype=torch.float32):
        self.act_dtype = dtype

    def empty(self, *shape, device):
        # NaN-fill so that reads of never-written elements show up in the comparisons
        return torch.full(shape, float("nan"), dtype=self.act_dtype, device=device)

    def _r(self, t):            # one rounding to the storage dtype
        return t.to(self.act_dtype)

    # ------------------------------------------------------------------ packing
    def pack_x(self, X, Xp):
        B, C, T = X.shape
        Xp.zero_()
        Xp[:, :, :C] = self._r(X.permute(0, 2, 1))

    def pack_weight(self, w, out, mode):
        G, O, I, Kt = w.shape
        Np, Kp = out.shape[1], out.shape[2] // Kt
        out.zero_()
        o4 = out.view(G, Np, Kt, Kp)
        if mode == 0:
            o4[:, :O, :, :I] = self._r(w.permute(0, 1, 3, 2))
        else:
            o4[:, :I, :, :O] = self._r(w.flip(3).permute(0, 2, 3, 1))

    def pack_weights(self, items):
        for w, out, mode in items:
            self.pack_weight(w, out, mode)

    # ------------------------------------------------------------------ contractions
    def tapgemm(self, A, Ca, W, N, taps, dil, bias, Res, group, Out, epi="store", Out2=None, stats=None):
        B, T, _ = A.shape
        G, Np, Ktot = W.shape
        Kp = Ktot // taps
        a = A[:, :, :Ca].float()
        w = W.view(G, Np, taps, Kp)[:, :N, :, :Ca].float()
        half = taps // 2
        acc = torch.zeros(B, T, N)
        for j in range(taps):
            off = (j - half) * dil
            sh = torch.zeros_like(a)
            lo, hi = max(0, -off), min(T, T - off)
            if hi > lo:
                sh[:, lo:hi] = a[:, lo + off:hi + off]
            wj = w[:, :, j, :]                                   # [G,N,Ca]
            wb = wj[group.long()] if group is not None else wj[0].expand(B, N, Ca)
            acc += torch.einsum("btc,bnc->btn", sh, wb)
        if bias is not None:
            bb = bias.view(G, N)
            acc += (bb[group.long()] if group is not None else bb[0].expand(B, N))[:, None, :]
        if Res is not None:
            acc += Res[:, :, :N].float()
        Out.zero_()
        if epi == "store":
            Out[:, :, :N] = self._r(acc)
        elif epi == "store_f32":
            Out[:, :, :N] = acc
            if stats is not None:
                self.bn_stats(Out, N, stats.view(2, N))
        elif epi == "gelu":
            Out2.zero_()
            Out2[:, :, :N] = self._r(acc)
            Out[:, :, :N] = self._r(F.gelu(acc))
        elif epi == "glu":
            Ch = N // 2
            Out2.zero_()
            Out2[:, :, :N] = self._r(acc)
            Out[:, :, :Ch] = self._r(acc[:, :, :Ch] * torch.sigmoid(acc[:, :, Ch:]))
        else:
            raise ValueError(epi)

    def group_order(self, group, G, order):
        order.copy_(torch.sort(group.long(), stable=True).indices.to(torch.int32))

    def wgrad(self, dU, M, Xin, C, taps, dil, group, order, dW):
        B, T, _ = dU.shape
        G = dW.shape[0] if dW.dim() == 4 else 1
        du, x = dU[:, :, :M].float(), Xin[:, :, :C].float()
        half = taps // 2
        out = torch.zeros(G, M, C, taps)
        for j in range(taps):
            off = (j - half) * dil
            sh = torch.zeros_like(x)
            lo, hi = max(0, -off), min(T, T - off)
            if hi > lo:
                sh[:, lo:hi] = x[:, lo + off:hi + off]
            per = torch.einsum("btm,btc->bmc", du, sh)
            if group is None:
                out[0, :, :, j] = per.sum(0)
            else:
                out[:, :, :, j].index_add_(0, group.long(), per)
        dW += out.view(dW.shape)

    def colsum(self, U, C, group, out):
        s = U[:, :, :C].float().sum(1)                           # [B,C]
        if group is None:
            out += s.sum(0).view(out.shape)
        else:
            out.view(-1, C).index_add_(0, group.long(), s)

    def matmul(self, A, Bm, out, transA=False, transB=False, alpha=1.0, beta=0.0):
        a = A.transpose(-1, -2) if transA else A
        b = Bm.transpose(-1, -2) if transB else Bm
        r = alpha * torch.matmul(a, b)
        out.copy_(r + beta * out if beta != 0.0 else r)

    # ------------------------------------------------------------------ spatial attention
    def sa_softmax(self, a, mask, W, Wm):
        w = F.softmax(a, dim=-1)
        W.copy_(w)
        Wm.copy_(w * mask[None, :] if mask is not None else w)

    def sa_softmax_bwd(self, W, dWm, mask, da):
        g = dWm * mask[None, :] if mask is not None else dWm
        da.copy_(W * (g - (W * g).sum(-1, keepdim=True)))

    # ------------------------------------------------------------------ BN / GELU / GLU
    def bn_stats(self, U, C, sums):
        u = U[:, :, :C].float()
        sums[0] += u.sum((0, 1))
        sums[1] += (u * u).sum((0, 1))

    def bn_finalize(self, sums, gamma, beta, rm, rv, count, eps, momentum, training, scale, shift, mean, rstd):
        if training:
            m = sums[0] / count
            var = (sums[1] / count - m * m).clamp_min(0)
            rm.mul_(1 - momentum).add_(momentum * m)
            rv.mul_(1 - momentum).add_(momentum * var * count / max(count - 1, 1))
        else:
            m, var = rm.clone(), rv.clone()
        rs = torch.rsqrt(var + eps)
        mean.copy_(m)
        rstd.copy_(rs)
        scale.copy_(gamma * rs)
        shift.copy_(beta - m * gamma * rs)

    def bn_gelu(self, U, scale, shift, H, C):
        H.zero_()
        H[:, :, :C] = self._r(F.gelu(U[:, :, :C].float() * scale + shift))

    def bn_gelu_train(self, U, sums, gamma, beta, rm, rv, count, eps, momentum, scale, shift, mean, rstd, H, C):
        self.bn_finalize(sums, gamma, beta, rm, rv, count, eps, momentum, True, scale, shift, mean, rstd)
        self.bn_gelu(U, scale, shift, H, C)

    def bn_gelu_bwd_apply(self, dH, U, scale, shift, gamma, mean, rstd, sums2, count, dU, C, dbias=None):
        u = U[:, :, :C].float()
        dy = dH[:, :, :C].float() * _gelu_grad(u * scale + shift)
        xh = (u - mean) * rstd
        du = gamma * rstd * (dy - sums2[0] / count - xh * sums2[1] / count)
        dU.zero_()
        dU[:, :, :C] = self._r(du)
        if dbias is not None:
            dbias += dU[:, :, :C].float().sum((0, 1))

    def bn_gelu_bwd_reduce(self, dH, U, scale, shift, mean, rstd, dY, sums2, C):
        u = U[:, :, :C].float()
        gp = _gelu_grad(u * scale + shift)
        dy = self._r(dH[:, :, :C].float() * gp).float() if dY is not None else dH[:, :, :C].float() * gp
        xh = (u - mean) * rstd
        if dY is not None:
            dY.zero_()
            dY[:, :, :C] = self._r(dy)
        sums2[0] += dy.sum((0, 1))
        sums2[1] += (dy * xh).sum((0, 1))

    def bn_bwd_apply(self, dY, U, gamma, mean, rstd, sums2, count, dU, C, dbias=None):
        xh = (U[:, :, :C].float() - mean) * rstd
        du = gamma * rstd * (dY[:, :, :C].float() - sums2[0] / count - xh * sums2[1] / count)
        dU.zero_()
        dU[:, :, :C] = self._r(du)
        if dbias is not None:
            dbias += dU[:, :, :C].float().sum((0, 1))

    def glu(self, V, Out, C):
        Out.zero_()
        Out[:, :, :C] = self._r(V[:, :, :C].float() * torch.sigmoid(V[:, :, C:2 * C].float()))

    def glu_bwd(self, dOut, V, dV, C, dbias=None):
        d, a, s = dOut[:, :, :C].float(), V[:, :, :C].float(), torch.sigmoid(V[:, :, C:2 * C].float())
        dV.zero_()
        dV[:, :, :C] = self._r(d * s)
        dV[:, :, C:2 * C] = self._r(d * a * s * (1 - s))
        if dbias is not None:
            dbias += dV[:, :, :2 * C].float().sum((0, 1))

    def gelu(self, P, Q):
        Q.copy_(self._r(F.gelu(P.float())))

    def gelu_bwd(self, dQ, P, dP, C, dbias=None):
        dP.zero_()
        dP[:, :, :C] = self._r(dQ[:, :, :C].float() * _gelu_grad(P[:, :, :C].float()))
        if dbias is not None:
            dbias += dP[:, :, :C].float().sum((0, 1))

    def time_mean(self, Z, Zm, Fo):
        Zm.copy_(self._r(Z.float().mean(1, keepdim=True)))

    def time_mean_bwd(self, dZm, dZ, Fo):
        dZ.copy_(self._r((dZm.float() / dZ.shape[1]).expand_as(dZ)))

    # ------------------------------------------------------------------ CLIP
    def avgpool(self, X, Out, C, k=3, stride=2):
        Out.copy_(self._r(F.avg_pool1d(X.float().permute(0, 2, 1), k, stride=stride).permute(0, 2, 1)))

    def avgpool_bwd(self, dOut, dX, C, k=3, stride=2):
        with torch.enable_grad():
            x = torch.zeros(dX.shape, dtype=torch.float32).permute(0, 2, 1).contiguous().requires_grad_(True)
            y = F.avg_pool1d(x, k, stride=stride)
            g, = torch.autograd.grad(y, x, dOut.float().permute(0, 2, 1))
        dX.copy_(self._r(g.permute(0, 2, 1)))

    def clip_prep_y(self, Y, Yp, ysq):
        B = Y.shape[0]
        y = (Y if Y.dim() == 3 else Y[:, :, None]).float()
        Yp.copy_(self._r(y.permute(0, 2, 1).reshape(B, -1)).view(Yp.shape))
        ysq += (Y.float().reshape(B, -1) ** 2).sum(-1)

    def rowsumsq(self, Z, sq):
        sq += (Z.float().reshape(Z.shape[0], -1) ** 2).sum(-1)

    def clip_logits(self, Yp, Zp, raw):
        raw += Yp.float().reshape(raw.shape[0], -1) @ Zp.float().reshape(raw.shape[1], -1).T

    def clip_ce_partial(self, raw, ysq, zsq, temp, Lg, row_max, row_sumexp, col_lse):
        Lv = raw * torch.exp(temp) / torch.sqrt(ysq[:, None] * zsq[None, :])
        Lg.copy_(Lv)
        m = Lv.max(dim=1).values
        row_max.copy_(m)
        row_sumexp.copy_(torch.exp(Lv - m[:, None]).sum(1))
        col_lse.copy_(torch.logsumexp(Lv, dim=0))

    def clip_lse_merge(self, parts, row_lse):
        row_lse.copy_(torch.logsumexp(parts[:, 0, :] + torch.log(parts[:, 1, :]), dim=0))

    def clip_ce_finish(self, Lg, row_lse, col_lse, ysq, zsq, temp, off, wscale, out2, coef, cvec):
        By, Bz = Lg.shape
        E = torch.zeros(By, Bz)
        j = torch.arange(Bz)
        E[off + j, j] = 1.0
        dL = (torch.exp(Lg - row_lse[:, None]) + torch.exp(Lg - col_lse[None, :]) - 2 * E) * wscale
        pos = Lg[off + j, j]
        out2[0] += ((row_lse[off + j] - pos) + (col_lse - pos)).sum() * wscale
        out2[1] += (dL * Lg).sum()
        c = dL * torch.exp(temp) / torch.sqrt(ysq[:, None] * zsq[None, :])
        coef.zero_()
        coef[:, :By] = self._r(c.T)
        cvec += (dL * Lg).sum(0) / zsq

    def clip_dz(self, coef, Yp, Z, cvec, gscale, dZ):
        By = Yp.shape[0]
        r = coef[:, :By].float() @ Yp.float().reshape(By, -1) - Z.float().reshape(Z.shape[0], -1) * cvec[:, None]
        dZ.copy_(self._r(gscale * r).view(dZ.shape))

    def retrieval_rank(self, raw, zsq, rank, off=0):
        Bz = raw.shape[1]
        sim = raw[off:off + Bz] / torch.sqrt(zsq)[None, :]
        own = sim.diagonal()
        rank.copy_(((sim > own[:, None]).sum(1)).to(torch.int32))

    def cast(self, src, dst):
        dst.copy_(src.to(dst.dtype))
