""""Next" row N2 — the batch collator.  CPU tier: the oracle restatement against the golden vectors produced by the
UNMODIFIED reference functions (oracle/make_golden_collator.py).  GPU tier: megb200_robust_scale (through
megb200.data) against the same vectors and against the oracle at the full Gwilliams batch shape."""
import os

import numpy as np
import pytest
import torch

from oracle import collator_restatement as CR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "collator.npz")


def _cases():
    g = np.load(GOLD)
    i = 0
    while f"X{i}" in g:
        L, lim, clamp = g[f"meta{i}"]
        yield i, g[f"X{i}"], g[f"out{i}"], int(L), float(lim), bool(clamp)
        i += 1


def test_oracle_collator_matches_reference_golden():
    n = 0
    for i, X, want, L, lim, clamp in _cases():
        got = CR.collate_brain(X, L, lim, clamp)
        assert np.abs(got - want).max() <= 2e-6 * (np.abs(want).max() + 1e-30), i
        n += 1
    assert n >= 4


@pytest.mark.gpu
def test_device_collator_matches_reference_golden():
    from megb200.data import robust_scale_
    for i, X, want, L, lim, clamp in _cases():
        got = robust_scale_(torch.from_numpy(X.copy()).cuda(), L, lim, clamp).cpu().numpy()
        err = np.abs(got - want).max() / (np.abs(want).max() + 1e-30)
        bad = np.unravel_index(np.abs(got - want).argmax(), got.shape)
        print(f"[parity] collator case {i}: rel err {err:.2e} (worst at {bad}: got {got[bad]:.6g} want {want[bad]:.6g})")
        assert err <= 1e-5, (i, err)


@pytest.mark.gpu
def test_device_collator_full_batch_against_oracle_and_module():
    from megb200 import synth
    from megb200.data import DeviceCollator
    g = np.random.Generator(np.random.PCG64(7))
    B, C, T = 64, 208, 360
    X = (g.standard_normal((B, C, T)) * g.uniform(0.5, 40.0, size=(B, C, 1))).astype(np.float32)
    args = synth.Args(preprocs={"brain_resample_rate": 120, "baseline_len_sec": 0.5, "clamp": True, "clamp_lim": 20})
    coll = DeviceCollator(args, "cuda:0")
    batch = [(torch.from_numpy(X[b]), torch.zeros(4, 3), b % 5) for b in range(B)]
    Xd, Y, subj = coll(batch)
    assert Xd.is_cuda and Xd.shape == (B, C, T) and Y.shape == (B, 4, 3) and subj.dtype == torch.int32
    want = CR.collate_brain(X, 60, 20.0, True)
    err = np.abs(Xd.cpu().numpy() - want).max() / 20.0
    print(f"[parity] collator B={B}: max abs err / clamp = {err:.2e}")
    assert err <= 1e-5
    # size-independent properties: clamp range; zero median and unit IQR (numpy's interpolated quantiles) on every row
    assert float(Xd.abs().max()) <= 20.0
    q = torch.quantile(Xd.double(), torch.tensor([0.25, 0.5, 0.75], dtype=torch.float64, device=Xd.device), dim=2)
    assert float(q[1].abs().max()) <= 1e-5
    assert float((q[2] - q[0] - 1.0).abs().max()) <= 1e-5


@pytest.mark.gpu
def test_device_collator_argument_errors():
    from megb200.data import robust_scale_
    with pytest.raises(RuntimeError):
        robust_scale_(torch.zeros(2, 3, 8), 2)                     # CPU tensor
    with pytest.raises(RuntimeError):
        robust_scale_(torch.zeros(2, 3, 8, device="cuda"), 0)      # baseline_len < 1
