"""Drop-in ``meg_decoding.utils.loss.CLIPLoss`` (reference ``meg_decoding/utils/loss.py:55-112``).

    loss_func = CLIPLoss(args).to(device)      # train.py:155
    loss = loss_func(Y, Z)                     # train.py:203 — x = speech targets, y = brain embeddings

Same constructor (reads ``args.reduction`` / ``args.init_temperature``), same forward signature
``forward(x, y, fast=True, return_logits=False)``, same ``state_dict`` (single key ``temp``).  The arithmetic is the
megb200 kernel sequence (normalise -> B x B logits contraction -> symmetric cross-entropy) in one autograd node; gradients
flow to ``y`` and ``temp`` only, like the reference (``x`` is the precomputed target and carries none).

Data parallel: after ``megb200.parallel.init(...)`` the loss sees global negatives — the normalised targets are
all-gathered over NCCL and the row log-sum-exp is completed by an all-reduce.
"""
from __future__ import annotations

import torch
import torch.nn as nn

import megb200
from megb200 import parallel
from megb200.engine import ClipEngine

__all__ = ["CLIPLoss"]


class _ClipFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, Y, Zi, temp, want_logits):
        eng = module._engine(Zi.dtype)
        prepared = module._take_prefetched(Y, eng)
        loss, logits, saved = eng.forward(Y, Zi, temp.detach(), module.reduction, parallel.context(), want_logits,
                                          prepared=prepared)
        ctx.eng, ctx.saved = eng, saved
        if want_logits:
            ctx.mark_non_differentiable(logits)
            return loss, logits
        return loss

    @staticmethod
    def backward(ctx, gloss, *unused):
        g = gloss.reshape(1).to(torch.float32).contiguous()
        dZ, dtemp = ctx.eng.backward(ctx.saved, g)
        ctx.saved = None
        return None, None, dZ, dtemp.reshape(1), None


class CLIPLoss(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.reduction = args.reduction
        if self.reduction not in ("mean", "sum"):
            raise ValueError(f"megb200 CLIPLoss supports reduction 'mean' or 'sum', got {self.reduction!r}")
        self.temp = nn.Parameter(torch.tensor([float(args.init_temperature)]))
        self._engs = {}
        self._prefetched = None
        self._side_stream = None

    def _engine(self, dtype) -> ClipEngine:
        ops = megb200.get_ops("bf16" if dtype == torch.bfloat16 and megb200.get_precision() != "bf16_simt"
                              else ("bf16_simt" if dtype == torch.bfloat16 else "fp32"))
        eng = self._engs.get(id(ops))
        if eng is None or eng.ops is not ops:
            eng = self._engs[id(ops)] = ClipEngine(ops)
        return eng

    # ------------------------------------------------------------------ optional overlap of the target preparation
    def prefetch_targets(self, x, dtype=None):
        """Optional (megb200 extension, not in the reference): start packing + all-gathering the targets ``x`` on a side
        stream.  The targets do not depend on the encoder, so calling this BEFORE ``brain_encoder(X, ...)`` overlaps the
        transpose/cast pass and, in data parallel, the NCCL all-gather of the targets with the encoder forward.  The next
        ``forward(x, y)`` with the same tensor picks the result up; without this call nothing changes."""
        dev = self.temp.device
        if dev.type != "cuda":
            return
        if dtype is None:
            dtype = torch.float32 if megb200.get_precision() == "fp32" else torch.bfloat16
        eng = self._engine(dtype)
        Y = self._as_targets(x, dev)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        self._side_stream.wait_stream(cur)
        with torch.cuda.stream(self._side_stream):
            prepared = eng.prepare_targets(Y, parallel.context())
            done = torch.cuda.Event()
            done.record(self._side_stream)
        for t in ((Y.rows, Y.sumsq) if hasattr(Y, "rows") else (Y,)):
            t.record_stream(self._side_stream)
        self._prefetched = (self._key(Y), tuple(Y.shape), eng, prepared, done)

    @staticmethod
    def _key(Y):
        return Y.rows.data_ptr() if hasattr(Y, "rows") else Y.data_ptr()

    @staticmethod
    def _as_targets(x, dev):
        """The targets as the kernels take them: dense f32 or bf16 (B, F, T) | (B, F) on the device (bf16 is kept -- the
        loss contracts bf16 operands anyway, and a pipeline that stores the precomputed features in bf16 halves the
        host->device copy), or ``megb200.data.PackedTargets`` (already position-major + norms: no per-step pass at all)."""
        if hasattr(x, "rows"):
            return x if x.device == dev else x.to(dev)
        x = x.detach()
        if x.dtype != torch.bfloat16:
            x = x.to(torch.float32)
        return x.to(device=dev).contiguous()

    def _take_prefetched(self, Y, eng):
        pf, self._prefetched = self._prefetched, None
        if pf is None or pf[0] != self._key(Y) or pf[1] != tuple(Y.shape) or pf[2] is not eng:
            return None
        torch.cuda.current_stream(Y.device).wait_event(pf[4])
        for t in pf[3]:
            t.record_stream(torch.cuda.current_stream(Y.device))
        return pf[3]

    def forward(self, x, y, fast=True, return_logits=False):
        batch_size = x.size(0)
        assert batch_size > 1, "Batch size must be greater than 1."                      # loss.py:68
        if not fast:
            raise NotImplementedError("megb200: only the fast path of CLIPLoss (loss.py:86-99) is implemented; the slow "
                                      "CosineSimilarity path is never called by the reference scripts")
        dev = self.temp.device
        tag = getattr(y, "_megb200_runner", None)
        if tag is not None and torch.is_grad_enabled():
            # graph mode (megb200.graph.ModuleGraphs): y is the Z the encoder's recording just produced; loss forward +
            # the whole backward are one graph replay
            if return_logits:
                raise NotImplementedError("megb200 graph mode: return_logits needs the eager path "
                                          "(megb200.graph.accelerate(encoder, False))")
            return tag[0].loss_forward(self, x, tag[1])
        if y.dtype not in (torch.bfloat16, torch.float32):
            y = y.float()
        # brain embeddings -> position-major PC layout [B, T, F] (a free view for BrainEncoder's output)
        if y.dim() == 3:
            Zi = y.permute(0, 2, 1)
            if not Zi.is_contiguous():
                Zi = Zi.contiguous()
        elif y.dim() == 2:
            Zi = y.contiguous().unsqueeze(1)
        else:
            raise ValueError("CLIPLoss expects (B, F, T) or (B, F) embeddings")
        Y = self._as_targets(x, dev)
        out = _ClipFn.apply(self, Y, Zi.to(dev), self.temp, bool(return_logits))
        if return_logits:
            loss, logits = out
            return logits, loss
        return out
