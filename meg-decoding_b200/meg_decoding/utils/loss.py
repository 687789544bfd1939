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
        loss, logits, saved = eng.forward(Y, Zi, temp.detach(), module.reduction, parallel.context(), want_logits)
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

    def _engine(self, dtype) -> ClipEngine:
        ops = megb200.get_ops("bf16" if dtype == torch.bfloat16 and megb200.get_precision() != "bf16_simt"
                              else ("bf16_simt" if dtype == torch.bfloat16 else "fp32"))
        eng = self._engs.get(id(ops))
        if eng is None or eng.ops is not ops:
            eng = self._engs[id(ops)] = ClipEngine(ops)
        return eng

    def forward(self, x, y, fast=True, return_logits=False):
        batch_size = x.size(0)
        assert batch_size > 1, "Batch size must be greater than 1."                      # loss.py:68
        if not fast:
            raise NotImplementedError("megb200: only the fast path of CLIPLoss (loss.py:86-99) is implemented; the slow "
                                      "CosineSimilarity path is never called by the reference scripts")
        dev = self.temp.device
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
        Y = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = _ClipFn.apply(self, Y, Zi.to(dev), self.temp, bool(return_logits))
        if return_logits:
            loss, logits = out
            return logits, loss
        return out
