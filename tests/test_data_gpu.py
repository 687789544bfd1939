"""GPU tier: the host -> device input pipeline used by the end-to-end bench leg (megb200/data.py)."""
import pytest
import torch

from megb200.data import AsyncScalarReader, DevicePrefetcher

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_prefetcher_yields_every_batch_in_order_from_a_fixed_ring(depth):
    n = 9
    host = [(torch.full((4, 1 << 18), float(i)).pin_memory(), torch.arange(8) + i, torch.tensor([i], dtype=torch.int32))
            for i in range(n)]
    reader = AsyncScalarReader(DEV, lag=1)
    ptrs, sums = set(), []
    pf = DevicePrefetcher(iter(host), DEV, keep_on_host=(2,), depth=depth)
    for i, (x, idx, keep) in enumerate(pf):
        assert x.is_cuda and idx.is_cuda and not keep.is_cuda and int(keep) == i
        ptrs.add(x.data_ptr())
        # slow consumer work on the batch so that the copy stream would overrun it if slot reuse were not ordered
        acc = x
        for _ in range(20):
            acc = acc * 1.0 + 0.0
        v = reader.push(acc.mean() + idx.float().mean() * 0)
        if v is not None:
            sums.append(v)
    sums += reader.drain()
    assert sums == [float(i) for i in range(n)]
    assert len(ptrs) == depth + 1, "device buffers must come from the fixed ring, not the allocator"
