"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain torch, fp32 or fp64) of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product path (``meg-decoding_b200/``) never does.

Pinned against the reference itself: ``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (with import shims), loads the same weights, and asserts this restatement equals them
(fixtures under ``tests/golden/`` hold the *reference's* outputs).  The reference repo has no golden vectors
of its own, so those fixtures are the pin.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------- spatial attention
def fourier_basis(loc: torch.Tensor, K: int):
    """cos/sin tables (K*K, C); m = k*K + l, phi = 2*pi*(k*x + l*y).  models.py:174-195."""
    loc = torch.as_tensor(loc)
    kk = torch.arange(K).repeat_interleave(K).to(loc.dtype)
    ll = torch.arange(K).repeat(K).to(loc.dtype)
    phi = 2 * math.pi * (kk[:, None] * loc[None, :, 0] + ll[:, None] * loc[None, :, 1])
    return torch.cos(phi), torch.sin(phi)


def dropout_mask(loc: torch.Tensor, center: int, d_drop: float) -> torch.Tensor:
    """mask[c] = 0 if ||loc_c - loc_center|| < d_drop else 1 (strict <, no rescale).  models.py:235-238."""
    loc = torch.as_tensor(loc)
    d = (loc - loc[center]).norm(dim=-1)
    return torch.where(d < d_drop, 0.0, 1.0).to(loc.dtype)


def spatial_attention_weights(z_re, z_im, cos, sin):
    """softmax over sensors of a = Re(z)@cos + Im(z)@sin.  models.py:204-213."""
    return F.softmax(z_re @ cos + z_im @ sin, dim=-1)


# --------------------------------------------------------------------------------------- encoder forward
def encoder_forward(P: Dict[str, torch.Tensor], X: torch.Tensor, subject_idxs, loc, *, K: int, d_drop: float,
                    seq2seq: bool, training: bool, drop_center: Optional[int] = None,
                    bn_buffers: Optional[Dict[str, torch.Tensor]] = None, dilations=None,
                    return_intermediates: bool = False, ks=None, pooled: bool = False):
    """BrainEncoder.forward restated functionally.  models.py:200-220, 232-241, 267-273, 308-322, 367-383.

    P maps reference state_dict keys to tensors; ``subject_block.spatial_attention.z`` may be complex or
    given split as ``...z_re`` / ``...z_im``.  BN uses batch statistics when ``training`` (biased variance for
    normalisation, running stats updated in ``bn_buffers`` with momentum 0.1 and the unbiased variance),
    running statistics otherwise.  ``dilations`` (optional, 5x3) enables the paper variant
    (tests/test_modules/models.py:275-297); the reference itself uses 1 everywhere (models.py:289,297,305).
    ``ks`` (optional, 5 odd kernel sizes) and ``pooled`` restate ``BrainEncoderSeq2Static`` (models.py:465-512):
    block k uses kernel size ks[k] (``args.ConvBlocks.ks``), ``AvgPool1d(3, stride=2)`` follows blocks 0..3 and
    ``AdaptiveAvgPool1d(1)`` block 4 (models.py:483-487), so the 1x1 head runs on one time step.
    """
    dt = X.dtype
    inter = {}
    if "subject_block.spatial_attention.z" in P:
        z = P["subject_block.spatial_attention.z"]
        z_re, z_im = z.real.to(dt), z.imag.to(dt)
    else:
        z_re, z_im = P["subject_block.spatial_attention.z_re"], P["subject_block.spatial_attention.z_im"]
    loc = torch.as_tensor(loc).to(dt)
    cos, sin = fourier_basis(loc, K)
    W = spatial_attention_weights(z_re, z_im, cos, sin)
    if training:
        assert drop_center is not None, "training mode needs the spatial-dropout centre index"
        Xd = X * dropout_mask(loc, drop_center, d_drop)[None, :, None]
    else:
        Xd = X
    assert X.shape[1] == loc.shape[0]                                   # models.py:233
    h = torch.einsum("oi,bit->bot", W, Xd)
    inter["sa"] = h
    h = F.conv1d(h, P["subject_block.conv.weight"], P["subject_block.conv.bias"])
    Ws = torch.stack([P[f"subject_block.subject_layer.{int(s)}.weight"][:, :, 0] for s in subject_idxs])
    h = torch.bmm(Ws, h)                                                # == per-sample conv + cat, models.py:270-272
    inter["subject"] = h

    def bn(x, pre):
        g, b = P[pre + ".weight"], P[pre + ".bias"]
        if training:
            mean = x.mean(dim=(0, 2))
            var = x.var(dim=(0, 2), unbiased=False)
            if bn_buffers is not None:
                n = x.shape[0] * x.shape[2]
                with torch.no_grad():
                    bn_buffers[pre + ".running_mean"].mul_(0.9).add_(0.1 * mean.detach())
                    bn_buffers[pre + ".running_var"].mul_(0.9).add_(0.1 * var.detach() * n / (n - 1))
                    bn_buffers[pre + ".num_batches_tracked"] += 1
        else:
            src = bn_buffers if bn_buffers is not None else P
            mean, var = src[pre + ".running_mean"].to(dt), src[pre + ".running_var"].to(dt)
        return (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5) * g[None, :, None] + b[None, :, None]

    for k in range(5):
        pre = f"conv_blocks.conv{k}."
        dil = dilations[k] if dilations is not None else (1, 1, 1)
        kk = ks[k] if ks is not None else 3

        def conv(x, name, d):
            # padding="same" for an odd kernel: d * (k - 1) / 2 on both sides
            return F.conv1d(x, P[pre + name + ".weight"], P[pre + name + ".bias"], padding=d * (kk - 1) // 2, dilation=d)

        u0 = conv(h, "conv0", dil[0])
        if k > 0:
            u0 = u0 + h
        h0 = F.gelu(bn(u0, pre + "batchnorm0"))
        u1 = conv(h0, "conv1", dil[1]) + h0
        h1 = F.gelu(bn(u1, pre + "batchnorm1"))
        v = conv(h1, "conv2", dil[2])
        h = F.glu(v, dim=-2)
        if pooled:                                                       # models.py:483-487
            h = F.avg_pool1d(h, 3, stride=2) if k < 4 else h.mean(dim=2, keepdim=True)
        inter[f"block{k}"] = h
    h = F.gelu(F.conv1d(h, P["conv_final1.weight"], P["conv_final1.bias"]))
    h = F.gelu(F.conv1d(h, P["conv_final2.weight"], P["conv_final2.bias"]))
    if not seq2seq:
        h = h.mean(dim=2)
    if return_intermediates:
        return h, inter
    return h


# --------------------------------------------------------------------------------------- CLIP loss
def clip_loss(x: torch.Tensor, y: torch.Tensor, temp: torch.Tensor, reduction: str = "mean",
              return_logits: bool = False):
    """CLIPLoss.forward (fast path).  loss.py:66-112.  Called as loss_func(Y, Z): rows = x (speech)."""
    B = x.size(0)
    assert B > 1, "Batch size must be greater than 1."                  # loss.py:68
    xf, yf = x.reshape(B, -1), y.reshape(B, -1)
    xf = xf / xf.norm(dim=-1, keepdim=True)
    yf = yf / yf.norm(dim=-1, keepdim=True)
    logits = (xf @ yf.T) * torch.exp(temp)
    t = torch.arange(B)
    loss = (F.cross_entropy(logits, t, reduction=reduction) + F.cross_entropy(logits.t(), t, reduction=reduction)) / 2
    return (logits, loss) if return_logits else loss


def clip_loss_one_sided(x_all: torch.Tensor, y_shards, temp: torch.Tensor):
    """Data-parallel restatement: every rank holds all speech rows ``x_all`` (B_glob) and
    its own brain shard; returns the per-rank loss contributions whose sum is the global ``clip_loss``."""
    Bg = x_all.size(0)
    xf = x_all.reshape(Bg, -1)
    xf = xf / xf.norm(dim=-1, keepdim=True)
    s = torch.exp(temp)
    G, offs, o = [], [], 0
    for ys in y_shards:
        yf = ys.reshape(ys.size(0), -1)
        yf = yf / yf.norm(dim=-1, keepdim=True)
        G.append(s * (yf @ xf.T))                                       # (B_loc, B_glob)
        offs.append(o)
        o += ys.size(0)
    lse_speech = torch.logsumexp(torch.cat(G, 0), dim=0)                # over all brains, per speech row
    out = []
    for g, o in zip(G, offs):
        n = g.size(0)
        idx = torch.arange(n)
        pos = g[idx, o + idx]
        lse_brain = torch.logsumexp(g, dim=1)
        out.append(0.5 * ((lse_speech[o:o + n] - pos).sum() + (lse_brain - pos).sum()) / Bg)
    return out


# --------------------------------------------------------------------------------------- retrieval metric
def topk_accuracy(Z: torch.Tensor, Y: torch.Tensor, k: int = 10):
    """Classifier.forward restated without the Python double loop.  models.py:403-460: cosine similarity of
    every speech segment against every brain embedding, top-1 / top-k hit rate (rows = speech)."""
    B = Z.size(0)
    z = Z.reshape(B, -1)
    y = Y.reshape(B, -1)
    sim = (y / y.norm(dim=-1, keepdim=True)) @ (z / z.norm(dim=-1, keepdim=True)).T
    t = torch.arange(B)
    top1 = (sim.argmax(dim=1) == t).float().mean().item()
    kk = min(k, B)
    topk = (sim.topk(kk, dim=1).indices == t[:, None]).any(dim=1).float().mean().item()
    return top1, topk


# --------------------------------------------------------------------------------------- helpers
def state_to_torch(state: Dict[str, np.ndarray], dtype=torch.float32, requires_grad: bool = False):
    """numpy state (megb200.synth.make_state) -> torch tensors; the complex z is split into real leaves."""
    P = {}
    for k, v in state.items():
        t = torch.from_numpy(np.asarray(v))
        if k.endswith(".z"):
            P[k + "_re"] = t.real.to(dtype).clone().requires_grad_(requires_grad)
            P[k + "_im"] = t.imag.to(dtype).clone().requires_grad_(requires_grad)
            continue
        if t.is_floating_point():
            t = t.to(dtype)
            if requires_grad and "running_" not in k:
                t = t.clone().requires_grad_(True)
        P[k] = t
    return P


def paper_dilations():
    """The dilated ConvBlock of the paper as the reference's alternative formulation has it
    (tests/test_modules/models.py:275-297, blocks numbered k = 1..5): conv1 ``2 ** (2 * k % 5)``, conv2
    ``2 ** (2 * k + 1 % 5)`` -- operator precedence makes that ``2 ** (2 * k + 1)``, up to 2048 -- and conv3 2."""
    return [(2 ** (2 * k % 5), 2 ** (2 * k + 1 % 5), 2) for k in range(1, 6)]


def train_step(P, temp, X, Y, subject_idxs, loc, *, K, d_drop, seq2seq, drop_center, bn_buffers=None,
               reduction="mean", dilations=None, ks=None, pooled=False):
    """One training step's math: Z = encoder(X), loss = CLIP(Y, Z), grads by autograd (train.py:201-215)."""
    Z = encoder_forward(P, X, subject_idxs, loc, K=K, d_drop=d_drop, seq2seq=seq2seq, training=True,
                        drop_center=drop_center, bn_buffers=bn_buffers, dilations=dilations, ks=ks, pooled=pooled)
    loss = clip_loss(Y, Z, temp, reduction)
    leaves = {k: v for k, v in P.items() if v.requires_grad}
    grads = torch.autograd.grad(loss, list(leaves.values()) + [temp], allow_unused=True)
    g = dict(zip(list(leaves.keys()) + ["temp"], grads))
    return Z.detach(), loss.detach(), g
