"""Host -> device input pipeline for the hot path: double-buffered, pinned-memory prefetch on a copy stream so that the
H2D transfer of batch i+1 (454 MB of f32 X and Y per 256-sample Gwilliams-shaped batch, ~8 ms over PCIe 5) overlaps the
compute of batch i.  The reference moves every batch synchronously (``X.to(device), Y.to(device)``, train.py:199)."""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence

import torch


class PackedTargets:
    """Precomputed targets in the form the loss kernels contract: ``rows`` [B, T*F] position-major (time-major, features
    contiguous -- ``Y.permute(0, 2, 1).reshape(B, -1)`` of the reference's (B, F, T) tensor) in the activation dtype, and
    ``sumsq`` [B] f32 = |Y_b|^2 taken from the UNROUNDED values.  The speech targets are precomputed wav2vec2 features
    (north_star), so a data pipeline can store them like this once: the per-step transpose + cast pass over 566 MB
    disappears and the host->device copy of a 256-sample Gwilliams batch drops from 377 MB (f32) to 189 MB (bf16).
    ``CLIPLoss.forward(x, y)`` accepts one wherever it accepts the dense ``x``; results are identical to passing the
    dense f32 tensor (same rounding, same norms)."""

    def __init__(self, rows: torch.Tensor, sumsq: torch.Tensor, feature_shape=None):
        assert rows.dim() == 2 and sumsq.dim() == 1 and rows.shape[0] == sumsq.shape[0]
        self.rows, self.sumsq = rows, sumsq
        self.feature_shape = tuple(feature_shape) if feature_shape is not None else (rows.shape[1],)

    @classmethod
    def from_dense(cls, Y: torch.Tensor, dtype=torch.bfloat16) -> "PackedTargets":
        """Dataset-preparation helper (plain torch, any device; not on the per-step path): dense (B, F, T) | (B, F)."""
        B = Y.shape[0]
        Yf = Y.to(torch.float32)
        y3 = Yf if Yf.dim() == 3 else Yf[:, :, None]
        rows = y3.permute(0, 2, 1).reshape(B, -1).to(dtype).contiguous()
        return cls(rows, (Yf.reshape(B, -1) ** 2).sum(-1), tuple(Y.shape[1:]))

    def size(self, dim=None):
        shape = (self.rows.shape[0],) + self.feature_shape
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return self.size()

    def numel(self) -> int:
        return self.rows.numel()

    @property
    def device(self):
        return self.rows.device

    def to(self, device, non_blocking: bool = False) -> "PackedTargets":
        return PackedTargets(self.rows.to(device, non_blocking=non_blocking),
                             self.sumsq.to(device, non_blocking=non_blocking), self.feature_shape)

    def pin_memory(self) -> "PackedTargets":
        return PackedTargets(self.rows.pin_memory(), self.sumsq.pin_memory(), self.feature_shape)

    def __getitem__(self, idx) -> "PackedTargets":
        """Row gather (e.g. the batch's segment ids out of a store that keeps every segment's targets resident in HBM)."""
        return PackedTargets(self.rows[idx], self.sumsq[idx], self.feature_shape)

    def detach(self):
        return self


class DeviceTargetStore:
    """The dataset's precomputed targets resident in HBM, in the form the loss contracts (``PackedTargets``: position-major
    rows in the activation dtype + f32 norms): the companion of ``DeviceSegmentStore`` for the label side.  The reference
    keeps ``self.Y`` (one (F, T) wav2vec2 block per speech segment, gwilliams2022.py:415-422) on the host and collates it
    per batch; here a batch is ``store.gather(segment_ids)`` -- one row gather on the device -- so that a training step's
    host->device traffic is the brain data (or, with ``DeviceSegmentStore``, indices only).  A Gwilliams-sized label set
    (a few thousand segments x 737 KB in bf16) is a few GB of the 180 GB."""

    def __init__(self, Y_all, device="cuda", dtype=torch.bfloat16, chunk: int = 256):
        self.device = torch.device(device)
        if hasattr(Y_all, "rows"):
            self.packed = Y_all.to(self.device)
        else:                                    # dense (N, F, T) | (N, F): packed chunk by chunk to bound peak memory
            parts = [PackedTargets.from_dense(torch.as_tensor(Y_all[i:i + chunk]), dtype).to(self.device)
                     for i in range(0, len(Y_all), chunk)]
            self.packed = PackedTargets(torch.cat([p.rows for p in parts]), torch.cat([p.sumsq for p in parts]),
                                        parts[0].feature_shape)
        self.n = int(self.packed.rows.shape[0])

    def __len__(self):
        return self.n

    def gather(self, seg_ids) -> "PackedTargets":
        """Targets of the batch's segments (ids on the host or the device; IndexError outside [0, n) like the reference's
        list indexing).  Returns device-resident PackedTargets accepted by ``CLIPLoss.forward`` in place of ``Y``."""
        ids = torch.as_tensor(seg_ids)
        if ids.numel():
            lo, hi = int(ids.min()), int(ids.max())
            if lo < -self.n or hi >= self.n:
                raise IndexError("segment id {} is out of range".format(hi if hi >= self.n else lo))
        ids = ids.to(self.device, dtype=torch.int64, non_blocking=True)
        ids = torch.where(ids < 0, ids + self.n, ids)
        return PackedTargets(self.packed.rows.index_select(0, ids), self.packed.sumsq.index_select(0, ids),
                             self.packed.feature_shape)


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


class DeviceSegmentStore:
    """The Gwilliams datasets' ``__getitem__`` (gwilliams2022.py:130-143) for whole batches with the recordings resident
    in HBM ("next" row N3): ``recordings[key][f"task{t}"]`` = f32 (C, length), ``meg_onsets[key][f"task{t}"]`` = integer
    onsets, ``num_segments_foreach_task`` and ``valid_subjects`` exactly as the reference dataset object holds them.

    ``gather(seg_ids, rng=np.random)`` draws one recording key per sample like the reference
    (``np.random.choice(list(self.X.keys()))``) — or takes explicit ``rec`` indices — and returns
    ``(X [B, C, T] on the device, subject_idx int32 [B] on the host, index int64 [B, 3] = task / i_in_task / onset)``.
    """

    def __init__(self, recordings, meg_onsets, num_segments_foreach_task, valid_subjects, seq_len_samp, device="cuda"):
        import numpy as np
        self.device = torch.device(device)
        self.keys = list(recordings.keys())
        self.n_tasks = len(num_segments_foreach_task)
        self.T = int(seq_len_samp)
        vs = np.asarray(valid_subjects)
        self.subject_of_rec = np.array([int(np.where(vs == k.split("_")[0])[0][0]) for k in self.keys], dtype=np.int32)
        self._rec = []                                   # keeps the device recordings alive
        base, length, off, ons = [], [], [], []
        self.C = None
        for k in self.keys:
            for t in range(self.n_tasks):
                a = torch.as_tensor(np.ascontiguousarray(recordings[k][f"task{t}"]), dtype=torch.float32).to(self.device)
                self.C = a.shape[0] if self.C is None else self.C
                assert a.dim() == 2 and a.shape[0] == self.C, "all recordings must have the same channel count"
                self._rec.append(a)
                base.append(a.data_ptr())
                length.append(a.shape[1])
                o = np.asarray(meg_onsets[k][f"task{t}"], dtype=np.int64)
                assert len(o) == int(num_segments_foreach_task[t]), "one onset per segment of the task"
                off.append(sum(len(x) for x in ons))
                ons.append(o)
        i64 = dict(dtype=torch.int64, device=self.device)
        self.rec_base, self.rec_len = torch.tensor(base, **i64), torch.tensor(length, **i64)
        self.onset_off, self.onsets = torch.tensor(off, **i64), torch.as_tensor(np.concatenate(ons)).to(self.device)
        self.nseg_accum = torch.as_tensor(np.cumsum(np.asarray(num_segments_foreach_task, dtype=np.int64))).to(self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)

    def gather(self, seg_ids, rec=None, rng=None, check: bool = True):
        import numpy as np
        from . import lib
        seg = torch.as_tensor(seg_ids, dtype=torch.int64)
        B = seg.numel()
        if rec is None:
            rng = rng if rng is not None else np.random
            rec = np.array([self.keys.index(str(rng.choice(self.keys))) for _ in range(B)], dtype=np.int32)
        rec = np.asarray(rec, dtype=np.int32)
        out = torch.empty(B, self.C, self.T, dtype=torch.float32, device=self.device)
        index = torch.empty(B, 3, dtype=torch.int64, device=self.device)
        seg_d, rec_d = seg.to(self.device), torch.from_numpy(rec).to(self.device)
        self.error.zero_()
        with torch.cuda.device(self.device):
            lib.call("megb200_gather_segments", self.rec_base.data_ptr(), self.rec_len.data_ptr(), self.onset_off.data_ptr(),
                     self.onsets.data_ptr(), self.nseg_accum.data_ptr(), len(self.keys), self.n_tasks, seg_d.data_ptr(),
                     rec_d.data_ptr(), B, self.C, self.T, out.data_ptr(), index.data_ptr(), self.error.data_ptr(),
                     torch._C._cuda_getCurrentRawStream(self.device.index))
        if check:
            e = int(self.error.item())
            if e & 1:
                raise IndexError("segment id or recording index out of range")          # the reference raises IndexError too
            if e & 2:
                raise ValueError("a segment window runs past the end of its recording")  # torch.stack would fail downstream
        valid = (rec >= 0) & (rec < len(self.keys))
        subj = np.where(valid, self.subject_of_rec[np.clip(rec, 0, len(self.keys) - 1)], -1).astype(np.int32)
        return out, torch.from_numpy(subj), index


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
