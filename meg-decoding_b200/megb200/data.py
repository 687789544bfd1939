"""Host -> device input pipeline for the hot path: double-buffered, pinned-memory prefetch on a copy stream so that the
H2D transfer of batch i+1 (454 MB of f32 X and Y per 256-sample Gwilliams-shaped batch, ~8 ms over PCIe 5) overlaps the
compute of batch i.  The reference moves every batch synchronously (``X.to(device), Y.to(device)``, train.py:199)."""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence

import torch


class DevicePrefetcher:
    """Wraps an iterable of host batches (tuples of tensors; non-tensors and CPU index tensors listed in ``keep_on_host``
    pass through) and yields the same tuples with the tensors resident on ``device``.

    Copies are issued on a dedicated stream from pinned memory (tensors that are not pinned are pinned once per batch),
    ``depth`` batches ahead, into a fixed ring of ``depth + 1`` device buffer sets allocated on first use — the caching
    allocator is never involved in steady state, so a host that runs ahead of the device (no per-step sync) cannot
    trigger a cudaMalloc.  The consumer's stream waits on the copy's event; a slot is overwritten only after the
    consumer stream has passed the point where the NEXT batch was requested, so a yielded batch is valid until the
    following ``next()`` call plus the work already enqueued on the consumer stream (clone it to keep it longer).
    """

    def __init__(self, batches: Iterable[Sequence], device, keep_on_host: Sequence[int] = (), depth: int = 2):
        self.device = torch.device(device)
        self.keep = set(keep_on_host)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self.ring = [dict() for _ in range(self.depth + 1)]       # slot -> {(position, shape, dtype): device buffer}
        self.released = [None] * (self.depth + 1)                 # slot -> event: consumer is done with the slot
        self.n_issued = 0
        self.queue = []
        self.current = None
        self.reset(batches)

    def reset(self, batches: Iterable[Sequence]) -> "DevicePrefetcher":
        """Start over on a new iterable (next epoch), keeping the device ring."""
        self._release_current()
        self.queue.clear()
        self.it: Iterator = iter(batches)
        for _ in range(self.depth):
            self._issue()
        return self

    def _release_current(self):
        if self.current is not None:               # the batch handed out last time is done once the stream gets here
            rel = torch.cuda.Event()
            rel.record(torch.cuda.current_stream(self.device))
            self.released[self.current] = rel
            self.current = None

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return
        slot = self.n_issued % len(self.ring)
        self.n_issued += 1
        out = []
        with torch.cuda.stream(self.stream):
            if self.released[slot] is not None:
                self.stream.wait_event(self.released[slot])
                self.released[slot] = None
            for i, t in enumerate(batch):
                if not torch.is_tensor(t) or i in self.keep:
                    out.append(t)
                    continue
                if not t.is_pinned():
                    t = t.pin_memory()
                key = (i, tuple(t.shape), t.dtype)
                buf = self.ring[slot].get(key)
                if buf is None:
                    buf = self.ring[slot][key] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                buf.copy_(t, non_blocking=True)
                out.append(buf)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((out, ev, slot))

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        self._release_current()
        if not self.queue:
            raise StopIteration
        out, ev, slot = self.queue.pop(0)
        cur.wait_event(ev)
        self.current = slot
        self._issue()
        return tuple(out)


def robust_scale_(X: torch.Tensor, baseline_len: int, clamp_lim: float = 20.0, clamp: bool = True) -> torch.Tensor:
    """In-place device form of the reference collator's brain-data transform ("next" row N2): baseline correction,
    per-(sample, channel) RobustScaler over time, clamp.  X: contiguous f32 CUDA tensor [..., T]."""
    from . import lib
    if not (X.is_cuda and X.dtype == torch.float32 and X.is_contiguous()):
        raise RuntimeError("robust_scale_: contiguous f32 CUDA tensor expected")
    T = X.shape[-1]
    with torch.cuda.device(X.device):
        lib.call("megb200_robust_scale", X.data_ptr(), X.data_ptr(), X.numel() // T, T, int(baseline_len), float(clamp_lim),
                 int(bool(clamp)), torch._C._cuda_getCurrentRawStream(X.device.index))
    return X


class DeviceCollator:
    """``Gwilliams2022Collator`` (gwilliams2022.py:639-662) with the arithmetic on the GPU: same constructor argument
    (``args.preprocs``: brain_resample_rate, baseline_len_sec, clamp, clamp_lim) and the same ``batch -> (X, Y,
    subject_idx)`` contract, but X comes back resident on ``device`` (f32 [B, C, T]) — use it as the DataLoader's
    ``collate_fn`` with ``num_workers=0`` or call it on the collated CPU batch in the training loop.  The reference
    spends one sklearn ``RobustScaler`` fit per sample on CPU workers here."""

    def __init__(self, args, device="cuda"):
        pre = args.preprocs
        self.baseline_len_samp = int(pre["brain_resample_rate"] * pre["baseline_len_sec"])
        self.clamp, self.clamp_lim = bool(pre["clamp"]), float(pre["clamp_lim"])
        self.device = torch.device(device)

    def __call__(self, batch):
        X = torch.stack([item[0] for item in batch]).to(torch.float32)
        Y = torch.stack([item[1] for item in batch])
        subject_idx = torch.IntTensor([item[2] for item in batch])
        X = X.pin_memory().to(self.device, non_blocking=True) if not X.is_cuda else X.contiguous()
        return robust_scale_(X, self.baseline_len_samp, self.clamp_lim, self.clamp), Y, subject_idx

    forward = __call__


class AsyncScalarReader:
    """Pipelined device -> host read of one scalar per step (the loss the reference logs with ``loss.item()``,
    train.py:208).  ``push(t)`` enqueues a copy of ``t`` into a pinned slot behind the step's kernels and returns the value
    pushed ``lag`` steps earlier (None for the first ``lag`` calls), so the host never drains the stream between steps and
    step i+1 is enqueued while step i still runs; ``drain()`` returns the values still in flight.  Every step's value
    reaches the host — ``lag`` steps late.
    """

    def __init__(self, device, lag: int = 1):
        self.device = torch.device(device)
        self.lag = max(0, int(lag))
        self.cuda = self.device.type == "cuda"
        self.slots = [torch.empty(1, dtype=torch.float32, pin_memory=self.cuda) for _ in range(self.lag + 1)]
        self.pending = []          # (slot index, event)
        self.n = 0

    def push(self, t: torch.Tensor):
        i = self.n % len(self.slots)
        self.n += 1
        self.slots[i].copy_(t.detach().reshape(1).to(torch.float32), non_blocking=True)
        ev = None
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self.pending.append((i, ev))
        if len(self.pending) > self.lag:
            return self._pop()
        return None

    def _pop(self) -> float:
        i, ev = self.pending.pop(0)
        if ev is not None:
            ev.synchronize()
        return float(self.slots[i][0])

    def drain(self):
        out = []
        while self.pending:
            out.append(self._pop())
        return out
