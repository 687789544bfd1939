"""GPU tier: a training step replayed from a CUDA graph (megb200.graph) must reproduce the eager step — it is the same
kernel sequence — for every new batch, subject assignment and spatial-dropout centre.  Not bit for bit: the BatchNorm
sums are f32 atomics whose order differs run to run, and a flipped last bit moves a few bf16 roundings downstream."""
import numpy as np
import pytest
import torch

import megb200
from megb200 import synth
from megb200.graph import GraphedTrainStep
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _eager(enc, lossf, X, Y, subj, center):
    for p in list(enc.parameters()) + list(lossf.parameters()):
        p.grad = None
    enc.drop_center = center
    loss = lossf(Y, enc(X, subj))
    loss.backward()
    grads = {k: v.grad.clone() for k, v in list(enc.named_parameters()) + list(lossf.named_parameters()) if v.grad is not None}
    return loss.detach().clone(), grads


# precision -> (loss rel, gradient rel-Frobenius, buffer rel-Frobenius).  The f32 check mode is the tight comparison (only
# f32 atomics reorder); in bf16 a flipped rounding moves whole activations on these 4-6 sample batches, so the bounds
# there only have to catch structural errors (a stale mask or batch gives O(1) differences).
TOL = {"fp32": (1e-4, 2e-3, 1e-4), "bf16": (1e-2, 0.15, 2e-2)}


@pytest.fixture(autouse=True)
def _restore_precision():
    old = megb200.get_precision()
    yield
    megb200.set_precision(old)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["narrow", "tiny_static"])
def test_graph_replay_equals_eager_step(name, precision):
    megb200.set_precision(precision)
    tl, tg, tb = TOL[precision]
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg, DEV)
    enc.train()
    state0 = {k: v.clone() for k, v in enc.state_dict().items()}
    X, Y, subj = H.make_inputs(cfg, DEV)
    step = GraphedTrainStep(enc, lossf, X, Y, subj, drop_center=synth.drop_center(cfg))
    assert step.launches_per_step > 20
    for k, v in enc.state_dict().items():
        assert torch.equal(v, state0[k]), f"building the graph changed {k}"
    gen = torch.Generator().manual_seed(3)
    for it in range(3):
        Xn = torch.randn(X.shape, generator=gen)
        Yn = torch.randn(Y.shape, generator=gen)
        sn = torch.randint(0, cfg.num_subjects, (cfg.B,), generator=gen, dtype=torch.int32)
        center = int(torch.randint(0, cfg.C, (1,), generator=gen))
        bufs = {k: v.clone() for k, v in enc.named_buffers()}
        loss_g = step(Xn.pin_memory(), Yn.pin_memory(), sn, drop_center=center).clone()
        grads_g = {k: v.grad.clone() for k, v in list(enc.named_parameters()) + list(lossf.named_parameters())}
        bufs_g = {k: v.clone() for k, v in enc.named_buffers()}
        with torch.no_grad():                       # rewind the BatchNorm statistics and redo the step eagerly
            for k, v in enc.named_buffers():
                v.copy_(bufs[k])
        loss_e, grads_e = _eager(enc, lossf, Xn.to(DEV), Yn.to(DEV), sn.to(DEV), center)
        enc.drop_center = None
        # run-to-run spread of the eager bf16 step itself is ~5e-4 on these tiny batches (atomic order -> bf16 roundings)
        assert abs(float(loss_g) - float(loss_e)) <= tl * abs(float(loss_e)), (it, float(loss_g), float(loss_e))
        for k, g in grads_e.items():
            if not k.endswith("bias") or "final" in k:       # conv biases in front of a BatchNorm have ~0 gradient
                a, b = (torch.view_as_real(t) if t.is_complex() else t for t in (grads_g[k], g))
                if b.numel() == 1:        # the temperature gradient is a small difference of large sums
                    assert abs(float(a) - float(b)) <= 2 * tg * abs(float(b)) + (5e-3 if precision == "bf16" else 1e-5), \
                        (it, k, float(a), float(b))
                else:
                    assert H.rel_fro(a, b) <= tg, (it, k, H.rel_fro(a, b))
        for k, v in enc.named_buffers():
            assert H.rel_fro(bufs_g[k].float(), v.float()) <= tb, (it, k)      # late-layer batch means inherit the bf16 spread


def _reference_loop(enc, lossf, classifier, optimizer, batches, centers):
    """The statements of the reference's hot loop (train.py:185-215), verbatim in order."""
    losses, accs = [], []
    enc.train()
    for (X, Y, subject_idxs), c in zip(batches, centers):
        enc.drop_center = c
        X, Y = X.to(DEV), Y.to(DEV)
        Z = enc(X, subject_idxs)
        loss = lossf(Y, Z)
        with torch.no_grad():
            top1, top10 = classifier(Z, Y)
        losses.append(loss.item())
        accs.append((top1, top10))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    return losses, accs


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["narrow", "tiny_static"])
def test_unchanged_training_loop_replayed_from_module_graphs(precision, name):
    """megb200.graph.ModuleGraphs: the reference loop, unchanged (enc(X, s) / lossf(Y, Z) / classifier / loss.item() /
    zero_grad / backward / Adam step), with the launches replayed from two CUDA graphs, against the same loop with eager
    launches from the same initial state: same loss trajectory over 7 optimizer steps (3 eager warm-up + 4 replayed)."""
    from meg_decoding.models import Classifier
    from megb200.graph import accelerate
    megb200.set_precision(precision)
    cfg = synth.get_config(name)
    batches = [H.make_inputs(cfg, "cpu", seed=i) for i in range(7)]
    centers = [(3 * i + 1) % cfg.C for i in range(7)]
    out = {}
    for mode in ("eager", "graph"):
        enc, lossf = H.make_modules(cfg, DEV)
        runner = accelerate(enc, mode == "graph")
        opt = torch.optim.Adam(list(enc.parameters()) + list(lossf.parameters()), lr=1e-3)
        losses, accs = _reference_loop(enc, lossf, Classifier(cfg | {"normalize_image_features": False}), opt, batches, centers)
        out[mode] = (losses, accs, {k: v.detach().float().clone() for k, v in enc.state_dict().items() if v.is_floating_point()})
        if mode == "graph":
            out["graph_modules"] = (enc, lossf)
            assert runner.stats["eager_steps"] == 3 and runner.stats["graph_steps"] == 4 and runner.stats["captures"] == 2
    le, lg = out["eager"][0], out["graph"][0]
    # fp32 is the tight comparison; in bf16 a flipped rounding moves whole activations on these 4-6 sample batches and
    # seven Adam steps amplify it (the eager loop differs from ITSELF by this much run to run)
    tol = 2e-3 if precision == "fp32" else 0.15
    print(f"[parity] module graphs {name} {precision}: eager {['%.4f' % v for v in le]} graph {['%.4f' % v for v in lg]}")
    for a, b in zip(le, lg):
        assert abs(a - b) / abs(a) <= tol
    assert all(np.isfinite(v) for v in le + lg)
    w = "conv_final2.weight"
    assert H.rel_fro(out["graph"][2][w], out["eager"][2][w]) <= (1e-3 if precision == "fp32" else 0.1)
    # guard rails (the modules of the graph run are in replay mode now): Z is differentiable only through the loss; a stale
    # loss cannot be back-propagated; eval / no_grad take the plain eager path
    enc, lossf = out["graph_modules"]
    enc.train()
    X, Y, subj = (t for t in batches[0])
    Z = enc(X.to(DEV), subj)
    assert hasattr(Z, "_megb200_runner")
    with pytest.raises(RuntimeError):
        Z.float().sum().backward()
    Z = enc(X.to(DEV), subj)
    old = lossf(Y.to(DEV), Z)
    Z2 = enc(X.to(DEV), subj)
    with pytest.raises(RuntimeError):
        old.backward()
    with pytest.raises(RuntimeError):
        lossf(Y.to(DEV), Z)                                   # the Z of an earlier step
    enc.eval()
    with torch.no_grad():
        Ze = enc(X.to(DEV), subj)
    assert not hasattr(Ze, "_megb200_runner")
    # absent subjects keep grad None (reference semantics) also when the step is replayed
    enc.train()
    for p in enc.parameters():
        p.grad = None
    s0 = torch.zeros_like(subj)
    lossf(Y.to(DEV), enc(X.to(DEV), s0)).backward()
    assert enc.subject_block.subject_layer[0].weight.grad is not None
    assert enc.subject_block.subject_layer[1].weight.grad is None
