"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the
CPU tests).  The hot path has exactly one exchange step, at the loss:

  * all-gather of the packed, constant speech targets and their squared norms (no gradient flows through it);
  * MAX + SUM all-reduce of the per-speech-row softmax partials (B_glob floats each);
  * SUM all-reduce of the parameter gradients (the loss already carries the 1/B_glob factor, so SUM — not mean).

BatchNorm statistics stay rank-local, as north_star lists no BN collective.  The spatial-dropout centre must be the
same on every rank for global-batch parity: ``share_drop_center`` broadcasts rank 0's draw.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class DistContext:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """[n, ...] per rank -> [world*n, ...] in rank order."""
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


_CTX: Optional[DistContext] = None


def init(group=None) -> Optional[DistContext]:
    """Enable global-negative CLIP over the (already initialised) default or given process group."""
    global _CTX
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        _CTX = None
    else:
        _CTX = DistContext(group)
    return _CTX


def shutdown():
    global _CTX
    _CTX = None


def context() -> Optional[DistContext]:
    return _CTX


def share_drop_center(center: int, device) -> int:
    """Rank 0's spatial-dropout centre for everyone (models.py:236 draws it from a process-local RNG)."""
    if _CTX is None:
        return center
    t = torch.tensor([center], dtype=torch.int64, device=device)
    dist.broadcast(t, src=0, group=_CTX.group)
    return int(t.item())


def allreduce_grads(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20):
    """SUM all-reduce of the gradients in flat f32 buckets (9.3 M parameters = 37 MB: one or two NCCL launches).

    Parameters whose grad is None on this rank (subjects absent from the local shard) contribute zeros so that
    every rank issues identical collectives; they end up with the summed gradient of the ranks that had them.
    """
    if _CTX is None:
        return
    plist = [p for p in params if p.requires_grad]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        views = []
        for p in bucket:
            if p.grad is None:          # absent on this rank: contribute zeros, receive the other ranks' sum
                p.grad = torch.zeros_like(p)
            g = p.grad
            if not g.is_contiguous():
                g = p.grad = g.contiguous()
            views.append(torch.view_as_real(g).reshape(-1) if g.is_complex() else g.reshape(-1))
        flat = torch.cat(views)                                    # one gather kernel
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_CTX.group)
        torch._foreach_copy_(views, list(flat.split([v.numel() for v in views])))   # one multi-tensor scatter
        bucket, size = [], 0

    for p in plist:
        bucket.append(p)
        size += p.numel() * (8 if p.is_complex() else 4)
        if size >= bucket_bytes:
            flush()
    flush()
