"""N-GPU tier (needs >= 2 CUDA devices; skipped otherwise): the data-parallel path on the REAL collectives -- NCCL over
NVLink plus the copy-engine peer gather of the packed targets -- against the oracle's emulation of what data parallel
means for this model: the reference run per shard with rank-local BatchNorm statistics, ONE global
CLIP loss over the concatenated embeddings, gradients of that loss (= the SUM of the ranks' gradients).

Checked on every rank, for the drop-in modules (eager launches + parallel.allreduce_grads) and for the CUDA-graph step
(megb200.graph.GraphedTrainStep, gradients all-reduced as one flat buffer inside the graph):
  * the reported loss is the GLOBAL loss (1e-3 in the fp32 check mode, 2e-2 in bf16);
  * the gathered targets are bit-identical to an NCCL all-gather of the same slabs (the peer gather moves bytes);
  * parameter gradients after the all-reduce equal the oracle's global-loss gradients.
Run it with:  gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu -q -s
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, precision, gather):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "meg-decoding_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MEGB200_DP_GATHER"] = gather
    import numpy as np
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import megb200
    from megb200 import parallel, synth
    from megb200.graph import GraphedTrainStep
    from tests import helpers as H
    out = {"rank": rank}
    try:
        megb200.set_precision(precision)
        ctx = parallel.init()
        cfg = synth.get_config("narrow", B=8)
        enc, lossf = H.make_modules(cfg, dev)
        enc.train()
        center = parallel.share_drop_center(synth.drop_center(cfg) if rank == 0 else 0, dev)
        enc.drop_center = center
        X, Y, subj = H.make_inputs(cfg, dev, seed=rank)
        params = list(enc.parameters()) + list(lossf.parameters())
        names = [k for k, _ in enc.named_parameters()] + ["temp"]

        # ---- drop-in modules, eager launches
        for p in params:
            p.grad = None
        lossf.prefetch_targets(Y)
        Z = enc(X, subj)
        loss = lossf(Y, Z)
        loss.backward()
        parallel.allreduce_grads(params)
        torch.cuda.synchronize()
        out["gather_mode"], out["gather_note"] = ctx.gather_mode, ctx.gather_note
        out["eager_loss"] = float(loss)
        out["Z"] = Z.detach().float().cpu().numpy()
        out["eager_grads"] = {n: (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).float().cpu().numpy()
                              for n, p in zip(names, params) if p.grad is not None}

        # ---- the gathered targets: peer copy-engine gather vs an NCCL all-gather of the same slab, bit for bit
        eng = lossf._engine(torch.float32 if precision == "fp32" else torch.bfloat16)
        Yall, ysq_all = eng.prepare_targets(Y, ctx)
        B, D = Y.shape[0], Y.numel() // Y.shape[0]
        mine = Yall.view(world, B, D)[rank].contiguous()
        ref_all = ctx.all_gather_rows(mine)
        torch.cuda.synchronize()
        out["gather_bit_exact"] = bool(torch.equal(ref_all.view(torch.int16 if Yall.dtype == torch.bfloat16 else torch.int32),
                                                   Yall.view(torch.int16 if Yall.dtype == torch.bfloat16 else torch.int32)))
        out["ysq_bit_exact"] = bool(torch.equal(ctx.all_gather_rows(ysq_all.view(world, B)[rank].contiguous()), ysq_all))

        # ---- CUDA-graph step (collectives inside the recording); two replays on the same batch
        saved = {k: v.clone() for k, v in enc.named_buffers()}
        gstep = GraphedTrainStep(enc, lossf, X, Y, subj, drop_center=center)
        with torch.no_grad():
            for k, v in enc.named_buffers():
                v.copy_(saved[k])
        for _ in range(2):
            gl = gstep()
        torch.cuda.synchronize()
        out["graph_loss"] = float(gl)
        out["graph_grads"] = {n: (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).float().cpu().numpy()
                              for n, p in zip(names, params) if p.grad is not None}
        out["launches_per_step"] = gstep.launches_per_step
        gstep = None
        out["ok"] = True
    except Exception as e:          # report instead of hanging the other rank in a collective
        import traceback
        out["ok"], out["error"] = False, f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"
    q.put(out)
    try:
        torch.cuda.synchronize()
        parallel.shutdown()
        dist.destroy_process_group()
    except Exception:
        pass


def _oracle(world):
    import numpy as np
    from megb200 import synth
    from oracle import restatement as R
    cfg = synth.get_config("narrow", B=8)
    P = R.state_to_torch(synth.make_state(cfg), requires_grad=True)
    temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
    loc, center = synth.sensor_locations(cfg.C), synth.drop_center(cfg)
    Zs, Ys = [], []
    for r in range(world):
        X, Y, subj = synth.make_batch(cfg, seed=r)
        Zs.append(R.encoder_forward(P, torch.from_numpy(X), torch.from_numpy(subj), loc, K=cfg.K, d_drop=cfg.d_drop,
                                    seq2seq=cfg.seq2seq, training=True, drop_center=center))     # rank-local BatchNorm
        Ys.append(torch.from_numpy(Y))
    loss = R.clip_loss(torch.cat(Ys), torch.cat(Zs), temp)                                        # ONE global loss
    leaves = {k: v for k, v in P.items() if v.requires_grad}
    grads = torch.autograd.grad(loss, list(leaves.values()) + [temp], allow_unused=True)
    g = dict(zip(list(leaves.keys()) + ["temp"], grads))
    zk = "subject_block.spatial_attention.z"
    g[zk] = torch.stack([g.pop(zk + "_re"), g.pop(zk + "_im")], dim=-1)
    return [z.detach() for z in Zs], float(loss.detach()), g


@pytest.mark.timeout(600)
@pytest.mark.parametrize("precision,gather", [("fp32", "peer"), ("bf16", "peer"), ("bf16", "nccl")])
def test_two_rank_nccl_step_matches_per_shard_bn_oracle(precision, gather):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from tests import helpers as H
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, precision, gather)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=500)
        res[r["rank"]] = r
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert res[r]["ok"], res[r].get("error")
    Zs, loss_ref, gref = _oracle(world)
    tz, tl, tg = (1e-3, 1e-3, 5e-3) if precision == "fp32" else (2e-2, 2e-2, 0.35)
    for r in range(world):
        o = res[r]
        if gather == "peer":
            assert o["gather_mode"] == "peer_copy_engine", f"peer gather not in use: {o['gather_note']}"
        else:
            assert o["gather_mode"] == "nccl"
        assert o["gather_bit_exact"] and o["ysq_bit_exact"], "gathered targets differ from an NCCL all-gather"
        ez = H.rel_fro(o["Z"], Zs[r])
        el = abs(o["eager_loss"] - loss_ref) / abs(loss_ref)
        eg = abs(o["graph_loss"] - loss_ref) / abs(loss_ref)
        worst = {"eager": (0.0, ""), "graph": (0.0, "")}
        for kind in ("eager", "graph"):
            for k, g in gref.items():
                if g is None:
                    continue
                if ".conv0.bias" in k or ".conv1.bias" in k:
                    if k.startswith("conv_blocks"):
                        continue                      # bias in front of a train-mode BatchNorm: true gradient is zero
                got = o[kind + "_grads"][k]
                e = H.rel_fro(got.reshape(-1), g.reshape(-1))
                if e > worst[kind][0]:
                    worst[kind] = (e, k)
        print(f"[parity] DP world=2 {precision} gather={o['gather_mode']} rank {r}: Z={ez:.2e} loss eager={el:.2e} "
              f"graph={eg:.2e} worst grad eager={worst['eager'][0]:.2e} ({worst['eager'][1]}) "
              f"graph={worst['graph'][0]:.2e} ({worst['graph'][1]}) launches/step={o['launches_per_step']}")
        assert ez <= tz and el <= tl and eg <= tl
        assert worst["eager"][0] <= tg, worst["eager"]
        assert worst["graph"][0] <= tg, worst["graph"]
