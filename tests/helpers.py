"""Shared test helpers: build the drop-in modules from a synthetic config, load golden fixtures, error metrics."""
import os

import numpy as np
import torch

from megb200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_fro(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def sample_index(n: int, limit: int = 4096) -> np.ndarray:
    """Same index set as oracle/make_golden.py::sample_index."""
    if n <= limit:
        return np.arange(n)
    return (np.arange(limit, dtype=np.int64) * (n // limit)) + (np.arange(limit) % 7)


def subsample(t) -> np.ndarray:
    flat = torch.as_tensor(t).detach().float().cpu().reshape(-1).numpy()
    return flat[sample_index(flat.size)]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def make_modules(cfg, device="cpu", seed=0):
    """(BrainEncoder, CLIPLoss) with the deterministic synthetic weights loaded, on `device`."""
    from meg_decoding.models import get_model
    from meg_decoding.utils.loss import CLIPLoss
    cfg = synth.Args(cfg)
    cfg["sensor_locations"] = synth.sensor_locations_raw(cfg.C)
    enc = get_model(cfg)             # BrainEncoder, or BrainEncoderSeq2Static for model == "brain_endcoder_seq2static"
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state(cfg, seed).items()}
    enc.load_state_dict(sd, strict=True)
    lossf = CLIPLoss(cfg)
    return enc.to(device), lossf.to(device)


def make_inputs(cfg, device="cpu", seed=0, B=None):
    X, Y, subj = synth.make_batch(cfg, seed, B=B)
    return torch.from_numpy(X).to(device), torch.from_numpy(Y).to(device), torch.from_numpy(subj)


def train_step(enc, lossf, X, Y, subj, center):
    """forward + CLIP + backward through the public module API; returns (Z, loss, {name: grad})."""
    enc.train()
    enc.drop_center = center
    for p in list(enc.parameters()) + list(lossf.parameters()):
        p.grad = None
    Z = enc(X, subj)
    loss = lossf(Y, Z)
    loss.backward()
    grads = {k: p.grad for k, p in enc.named_parameters()}
    grads["temp"] = lossf.temp.grad
    return Z.detach(), loss.detach(), grads


def compare_with_golden(name, Z, loss, grads, tol_z, tol_loss, tol_grad, bn=None, tol_bn=None):
    """Checks a train step against the reference's stored outputs; returns the worst errors per kind."""
    g = load_golden(name)
    errs = {"Z": rel_fro(subsample(Z), g["Z_sample"]),
            "loss": abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))}
    assert tuple(Z.shape) == tuple(g["Z_shape"]), (Z.shape, g["Z_shape"])
    assert errs["Z"] <= tol_z, f"{name}: Z rel-Frobenius error {errs['Z']:.3e} > {tol_z}"
    assert errs["loss"] <= tol_loss, f"{name}: loss rel error {errs['loss']:.3e} > {tol_loss}"
    worst = 0.0
    for k, gr in grads.items():
        if "gnone:" + k in g:
            assert gr is None or float(gr.abs().max()) == 0.0, f"{k}: reference grad is None"
            continue
        ref_s, ref_n = g["gsample:" + k], float(g["gnorm:" + k])
        assert gr is not None, f"missing grad for {k}"
        cur = torch.view_as_real(gr) if gr.is_complex() else gr
        if ".conv0.bias" in k or ".conv1.bias" in k:
            if k.startswith("conv_blocks"):
                # bias in front of a train-mode BatchNorm: true gradient is exactly zero (rounding noise only)
                wn = float(g["gnorm:" + k.replace(".bias", ".weight")])
                assert float(cur.double().norm()) <= max(tol_grad, 1e-4) * wn, k
                continue
        e = float(np.linalg.norm(subsample(cur).astype(np.float64) - ref_s) / (np.linalg.norm(ref_s) + 1e-30))
        worst = max(worst, e)
        assert e <= tol_grad, f"{name}: grad {k} rel error {e:.3e} > {tol_grad} (|g|={ref_n:.3e})"
    errs["grad"] = worst
    if bn is not None:
        for k, v in bn.items():
            ref = g["bn:" + k]
            if np.issubdtype(ref.dtype, np.integer):
                assert int(v) == int(ref), k
            else:
                assert rel_fro(v.detach().cpu(), ref) <= tol_bn, k
    return errs
