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
    ``depth`` batches ahead; the consumer's stream waits on the copy's event, and the buffers are tied to the consumer
    stream with ``record_stream`` so the caching allocator cannot recycle them early.
    """

    def __init__(self, batches: Iterable[Sequence], device, keep_on_host: Sequence[int] = (), depth: int = 2):
        self.it: Iterator = iter(batches)
        self.device = torch.device(device)
        self.keep = set(keep_on_host)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue = []
        for _ in range(self.depth):
            self._issue()

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return
        out = []
        with torch.cuda.stream(self.stream):
            for i, t in enumerate(batch):
                if not torch.is_tensor(t) or i in self.keep:
                    out.append(t)
                    continue
                if not t.is_pinned():
                    t = t.pin_memory()
                out.append(t.to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((out, ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        out, ev = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in out:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(cur)
        self._issue()
        return tuple(out)
