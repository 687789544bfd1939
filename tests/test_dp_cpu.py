"""CPU tier: the N>1 path (one-sided global-negative CLIP + gradient all-reduce) with torch.distributed gloo,
world size 2, on top of the op restatement.  Checks that the data-parallel loss / gradients equal the single-process
loss over the concatenated batch and that the bucketed gradient all-reduce sums correctly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "meg-decoding_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import megb200
    from megb200 import parallel, synth
    from meg_decoding.utils.loss import CLIPLoss
    from oracle.ops_restatement import TorchOps
    megb200._set_ops_override(TorchOps(torch.float32))
    try:
        parallel.init()
        assert parallel.context().world == world
        g = torch.Generator().manual_seed(0)
        Bg, Fo, T = 8, 16, 5
        Y = torch.randn(Bg, Fo, T, generator=g)
        Z = torch.randn(Bg, Fo, T, generator=g)
        n = Bg // world
        lossf = CLIPLoss(synth.Args(reduction="mean", init_temperature=1.7))
        Zl = Z[rank * n:(rank + 1) * n].clone().requires_grad_(True)
        loss = lossf(Y[rank * n:(rank + 1) * n], Zl)
        loss.backward()
        parallel.allreduce_grads(lossf.parameters())
        # a parameter with no local gradient still takes part in the collective
        extra = torch.nn.Parameter(torch.ones(3))
        if rank == 0:
            extra.grad = torch.full((3,), 2.0)
        zc = torch.nn.Parameter(torch.zeros(2, dtype=torch.cfloat))
        zc.grad = torch.full((2,), complex(rank + 1, 1.0), dtype=torch.cfloat)
        parallel.allreduce_grads([extra, zc])
        # plain numpy payloads: torch tensors travel through the queue as shared-memory handles that die with this process
        q.put((rank, float(loss.detach()), Zl.grad.numpy().copy(), lossf.temp.grad.numpy().copy(),
               extra.grad.numpy().copy(), zc.grad.numpy().copy()))
        center = parallel.share_drop_center(5 + rank, "cpu")
        assert center == 5
    finally:
        megb200._set_ops_override(None)
        parallel.shutdown()
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_clip_matches_single_process():
    from oracle import restatement as R
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=150)
        res[r[0]] = (r[1],) + tuple(torch.from_numpy(a) for a in r[2:])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    Bg, Fo, T = 8, 16, 5
    Y = torch.randn(Bg, Fo, T, generator=g)
    Z = torch.randn(Bg, Fo, T, generator=g).requires_grad_(True)
    temp = torch.tensor([1.7], requires_grad=True)
    full = R.clip_loss(Y, Z, temp)
    full.backward()
    n = Bg // world
    for r in range(world):
        loss, dZ, dtemp, extra, zc = res[r]
        assert abs(loss - float(full)) < 1e-5, "every rank reports the global loss"
        assert torch.allclose(dZ, Z.grad[r * n:(r + 1) * n], atol=1e-6), "local dZ equals the global-loss gradient"
        assert torch.allclose(dtemp, temp.grad, atol=1e-5), "temp gradient is summed over ranks"
        assert torch.equal(extra, torch.full((3,), 2.0))
        assert torch.equal(zc, torch.full((2,), complex(3.0, 2.0), dtype=torch.cfloat))
