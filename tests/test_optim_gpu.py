"""GPU tier, "next" row N4: the one-launch Adam (megb200.optim.Adam -> megb200_adam_multi) against torch.optim.Adam, the
optimizer the reference constructs (train.py:160-162), on the real parameter set incl. the complex spatial-attention z,
with absent-subject gradients (None) and weight decay."""
import pytest
import torch

from megb200 import synth
from megb200.optim import Adam
from tests import helpers as H

DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_matches_torch_adam(wd):
    cfg = synth.get_config("narrow")
    enc, lossf = H.make_modules(cfg, DEV)
    params = list(enc.parameters()) + list(lossf.parameters())
    twins = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ours = Adam(params, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    ref = torch.optim.Adam(twins, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for it in range(4):
        for i, (p, q) in enumerate(zip(params, twins)):
            if i % 7 == 3 and it % 2 == 1:             # some parameters without a gradient on some steps
                p.grad = q.grad = None
                continue
            g = torch.randn(p.shape, device=DEV, generator=gen, dtype=torch.float32)
            if p.is_complex():
                g = torch.complex(g, torch.randn(p.shape, device=DEV, generator=gen))
            p.grad, q.grad = g.clone(), g.clone()
        ours.step()
        ref.step()
    worst = 0.0
    for p, q in zip(params, twins):
        a, b = (torch.view_as_real(t) if t.is_complex() else t for t in (p.detach(), q.detach()))
        err = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)
        worst = max(worst, err)
    print(f"[parity] adam wd={wd}: worst rel err {worst:.2e} over {len(params)} tensors")
    assert worst <= 2e-6


def _rand_grads(params, twins, gen):
    for p, q in zip(params, twins):
        g = torch.randn(p.shape, device=DEV, generator=gen, dtype=torch.float32)
        if p.is_complex():
            g = torch.complex(g, torch.randn(p.shape, device=DEV, generator=gen))
        p.grad, q.grad = g.clone(), g.clone()


def _worst(params, twins):
    worst = 0.0
    for p, q in zip(params, twins):
        a, b = (torch.view_as_real(t) if t.is_complex() else t for t in (p.detach(), q.detach()))
        worst = max(worst, float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12))
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["megb200", "torch"])
def test_adam_resume_from_checkpoint(source):
    """save -> load_state_dict -> step, against torch.optim.Adam doing the same (the reference's optimizer resumes):
    the loaded moments and step counts must be USED (not zeroed) and the kernel must write the NEW state tensors.
    `source` = whose checkpoint is loaded: this class's own, or one written by torch.optim.Adam."""
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg, DEV)
    params = list(enc.parameters()) + list(lossf.parameters())
    twins = [torch.nn.Parameter(p.detach().clone()) for p in params]
    kw = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    ours, ref = Adam(params, **kw), torch.optim.Adam(twins, **kw)
    gen = torch.Generator(device=DEV).manual_seed(2)
    for _ in range(3):
        _rand_grads(params, twins, gen)
        ours.step()
        ref.step()
    # deep copies, as torch.save / torch.load would give: Optimizer.load_state_dict ALIASES the tensors it is handed, so
    # a live state_dict() would keep changing under the optimizer that loaded it
    import copy
    ckpt = copy.deepcopy((ours if source == "megb200" else ref).state_dict())
    ckpt_ref = copy.deepcopy(ref.state_dict())
    # resume into FRESH optimizers (load before the first step) ...
    ours2, ref2 = Adam(params, **kw), torch.optim.Adam(twins, **kw)
    ours2.load_state_dict(copy.deepcopy(ckpt))
    ref2.load_state_dict(copy.deepcopy(ckpt_ref))
    for _ in range(2):
        _rand_grads(params, twins, gen)
        ours2.step()
        ref2.step()
    assert _worst(params, twins) <= 2e-6
    assert int(ours2.state[params[0]]["step"]) == 5
    # ... and into a LIVE optimizer (load after steps: the cached pointer table must be dropped)
    old_m = ours2.state[params[1]]["exp_avg"]
    ours2.load_state_dict(copy.deepcopy(ckpt))
    ref2.load_state_dict(copy.deepcopy(ckpt_ref))
    keep = old_m.clone()
    for _ in range(2):
        _rand_grads(params, twins, gen)
        ours2.step()
        ref2.step()
    assert _worst(params, twins) <= 2e-6
    assert torch.equal(old_m, keep), "the kernel kept writing the moment buffer of the replaced state"
    assert int(ours2.state[params[0]]["step"]) == 5


def test_adam_rejects_cpu_parameters():          # CPU tier: the product path has no CPU fallback
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        Adam([p]).step()
