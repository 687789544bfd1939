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


def _peer_worker(rank, world, port, q, fail_rank, fail_stage):
    """PeerGather's constructor with the three IPC entry points mocked (no GPU here): one rank's export or import
    fails; every rank must leave the constructor with the same verdict, through the same collectives."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "meg-decoding_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C

    from megb200 import lib, parallel
    calls = []

    def fake_call(name, *args):
        calls.append(name)
        if rank == fail_rank and name == "megb200_ipc_" + fail_stage:
            raise RuntimeError(name + ": invalid resource handle (mock)")
        if name == "megb200_ipc_export":
            C.memmove(args[1], bytes([rank + 1]) * 64, 64)
            C.c_longlong.from_address(args[2]).value = 128
        elif name == "megb200_ipc_import":
            C.c_void_p.from_address(args[1]).value = 0x1000 * C.c_ubyte.from_address(args[0]).value
        return 0

    lib.call = fake_call
    try:
        ctx = parallel.DistContext()
        verdict = "ok"
        try:
            pg = parallel.PeerGather(ctx, 16, torch.float32, torch.device("cpu"))
            assert pg.peer_ptr == {1 - rank: 0x1000 * (2 - rank) + 128}
            pg.close()
            assert calls[-1] == "megb200_ipc_release"
        except RuntimeError as e:
            verdict = str(e)
        # the ranks are still in step: one more collective completes
        t = torch.tensor([rank + 1.0])
        dist.all_reduce(t)
        q.put((rank, verdict, float(t), list(calls)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("fail_rank,fail_stage", [(-1, "none"), (1, "export"), (0, "import")])
def test_peer_gather_setup_fails_on_every_rank_or_none(fail_rank, fail_stage):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q, fail_rank, fail_stage)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=100)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r in range(world):
        verdict, s, calls = res[r]
        assert s == 3.0
        if fail_rank < 0:
            assert verdict == "ok"
        else:
            assert verdict.startswith("PeerGather: "), verdict
            if r == fail_rank:
                assert "(mock)" in verdict
            elif fail_stage == "export":
                assert "export failed on rank(s) [1]" in verdict       # the peers learn it from the handle exchange
                assert "megb200_ipc_import" not in calls
            else:
                assert "another rank" in verdict
                assert calls[-1] == "megb200_ipc_release"              # the mapping that did succeed is released
