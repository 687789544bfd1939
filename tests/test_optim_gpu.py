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


def test_adam_rejects_cpu_parameters():          # CPU tier: the product path has no CPU fallback
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        Adam([p]).step()
