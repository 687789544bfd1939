"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the
CPU tests).  The hot path has exactly one exchange step, at the loss:

  * all-gather of the packed, constant speech targets and their squared norms (no gradient flows through it).  On CUDA
    this is ``PeerGather``: the slabs live in peer-mapped memory (CUDA IPC) and every rank PULLS the other ranks' slots
    with copy-engine copies over NVLink on a side stream -- no SM is taken from the persistent one-CTA-per-SM
    contraction kernels that run meanwhile (an NCCL all-gather of 1.3 GB on SMs cost +25 % per tap-GEMM launch at 8
    GPUs); two 4-byte NCCL all-reduces fence it.  ``MEGB200_DP_GATHER=nccl`` (or any failure to map peer memory)
    selects the plain NCCL all-gather;
  * all-gather of the per-speech-row softmax partials ((max, sum exp) pairs, 2 x B_glob floats per rank), merged by
    ``megb200_clip_lse_merge`` -- one collective instead of a MAX and a SUM all-reduce;
  * SUM all-reduce of the parameter gradients (the loss already carries the 1/B_glob factor, so SUM — not mean), issued
    straight on the flat f32 pool the backward pass writes its gradients into (no gather / scatter copies); the loss
    scalar and the temperature gradient ride in the same buffer.

BatchNorm statistics stay rank-local, as north_star lists no BN collective.  The spatial-dropout centre must be the
same on every rank for global-batch parity: ``share_drop_center`` broadcasts rank 0's draw.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class PeerGather:
    """All-gather of equal-sized per-rank slabs through peer-mapped device memory, moved by the copy engines.

    ``buf`` is [world, slab_elems] of ``dtype`` on every rank; rank r's producer kernel writes slot r of its OWN buffer
    (``slot()``), ``gather()`` then pulls slot p from rank p's buffer into the local slot p for every peer p (ring-staggered
    so that each source is read by one reader at a time).  Cross-rank ordering: a 4-byte NCCL all-reduce after the
    producers (every slot complete before anyone reads it) and one after the copies (nobody overwrites a slot while a
    peer still reads it).  All calls enqueue on the current stream and are CUDA-graph capturable (memcpy nodes)."""

    def __init__(self, ctx, slab_elems: int, dtype, device):
        import ctypes as C

        from . import lib
        self.ctx, self.device = ctx, device
        self.world, self.rank = ctx.world, ctx.rank
        self.slab_elems = int(slab_elems)
        self.buf = torch.empty(self.world, self.slab_elems, dtype=dtype, device=device)
        self.slab_bytes = self.slab_elems * self.buf.element_size()
        self._fence = torch.zeros(1, dtype=torch.float32, device=device)
        self._bases, self.peer_ptr = {}, {}
        # Every rank runs the SAME sequence of collectives whatever fails locally (a rank that raised half-way would
        # leave its peers waiting in a collective it never joins): export -> all-gather of the handles -> import ->
        # agreement; only then does the constructor raise, on every rank or on none.
        err, mine = None, None
        try:
            handle = (C.c_ubyte * 64)()
            off = C.c_longlong(0)
            if device.type == "cuda":
                torch.cuda.synchronize(device)
            lib.call("megb200_ipc_export", self.buf.data_ptr(), C.addressof(handle), C.addressof(off))
            mine = (bytes(handle), int(off.value), self.slab_bytes)
        except Exception as e:              # e.g. memory from a VMM-backed (expandable segments) pool has no IPC handle
            err = e
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=ctx.group)
        if err is None:
            try:
                failed = [p for p, x in enumerate(everyone) if x is None]
                if failed:
                    raise RuntimeError("IPC export failed on rank(s) %s" % failed)
                self._import_peers(everyone)
            except Exception as e:
                err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=ctx.group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("PeerGather: " + ("%s: %s" % (type(err).__name__, err) if err is not None
                                                 else "peer mapping failed on another rank"))
        self.copies_per_gather = self.world - 1

    def _import_peers(self, everyone):
        import contextlib
        import ctypes as C

        from . import lib
        with (torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()):
            for p, (h, o, nb) in enumerate(everyone):
                if p == self.rank:
                    continue
                if nb != self.slab_bytes:
                    raise RuntimeError("ranks disagree on the slab size")
                hb = (C.c_ubyte * 64).from_buffer_copy(h)
                base = C.c_void_p(0)
                lib.call("megb200_ipc_import", C.addressof(hb), C.addressof(base))
                self._bases[p] = int(base.value)
                self.peer_ptr[p] = int(base.value) + o

    def slot(self, r=None) -> torch.Tensor:
        return self.buf[self.rank if r is None else r]

    def fence(self):
        """Cross-rank ordering point on the current stream (a 4-byte NCCL all-reduce)."""
        dist.all_reduce(self._fence, group=self.ctx.group)

    def pull(self):
        """Copy every peer's slot into the local buffer (copy engines; between two fences)."""
        from . import lib
        st = torch._C._cuda_getCurrentRawStream(self.device.index)
        for i in range(1, self.world):
            p = (self.rank + i) % self.world
            lib.call("megb200_copy_async", self.buf.data_ptr() + p * self.slab_bytes,
                     self.peer_ptr[p] + p * self.slab_bytes, self.slab_bytes, st)

    def gather(self) -> torch.Tensor:
        self.fence()           # every rank's slot is complete
        self.pull()
        self.fence()           # every rank is done reading: slots may be rewritten
        return self.buf

    def close(self):
        from . import lib
        for base in self._bases.values():
            try:
                lib.call("megb200_ipc_release", base)
            except RuntimeError:
                pass
        self._bases, self.peer_ptr = {}, {}


class DistContext:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._gathers = {}
        self.gather_mode = None          # "peer_copy_engine" | "nccl" once the first target gather has run
        self.gather_note = None

    def target_gather(self, B: int, D: int, dtype, device):
        """The PeerGather pair (packed targets [world, B*D], squared norms [world, B]) for this shape, or None when peer
        memory is not in play (CPU / gloo, MEGB200_DP_GATHER=nccl, or mapping failed -- noted in ``gather_note``)."""
        import os
        key = (B, D, dtype, str(device))
        if key in self._gathers:
            return self._gathers[key]
        pair = None
        want = os.environ.get("MEGB200_DP_GATHER", "peer")
        if device.type == "cuda" and want != "nccl" and dist.get_backend(self.group) == "nccl":
            # PeerGather's constructor raises on every rank or on none, so all ranks take the same branch here
            first = None
            try:
                first = PeerGather(self, B * D, dtype, device)
                pair = (first, PeerGather(self, B, torch.float32, device))
            except Exception as e:          # no peer access / IPC refused: the NCCL all-gather, on EVERY rank
                self.gather_note = f"peer mapping failed: {type(e).__name__}: {e}"[:200]
                if first is not None:
                    first.close()
                pair = None
        self.gather_mode = "peer_copy_engine" if pair is not None else "nccl"
        self._gathers[key] = pair
        return pair

    def close(self):
        for pair in self._gathers.values():
            if pair is not None:
                for g in pair:
                    g.close()
        self._gathers = {}

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
    if _CTX is not None:
        _CTX.close()
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


def allreduce_flat(flat: torch.Tensor):
    """SUM all-reduce of one flat buffer in place (the backward pass's gradient pool)."""
    if _CTX is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_CTX.group)
    return flat


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
