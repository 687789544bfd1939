"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Recipe: stub the five import-time
modules the hot path never uses (turtle, termcolor, mne, mne_bids, hydra), import ``meg_decoding.models`` and
``meg_decoding.utils.loss`` from /root/reference, replace ``ch_locations_2d`` by the synthetic layout, load the
deterministic weights of ``megb200.synth.make_state`` with ``load_state_dict``, run forward + CLIPLoss +
backward in fp32 on CPU, and store the reference's outputs (sub-sampled where large).  It also asserts that
``oracle/restatement.py`` reproduces the reference (<= 2e-5 relative) — that is what pins the oracle.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import os
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "meg-decoding_b200"))
sys.path.insert(0, ROOT)

from megb200 import synth  # noqa: E402
from oracle import restatement as R  # noqa: E402

REF = "/root/reference"
GOLDEN_CONFIGS = ["tiny", "tiny_static", "narrow", "narrow_static"]


def import_reference():
    for name, attrs in (("turtle", ["forward"]), ("termcolor", ["cprint"]), ("mne", []), ("mne_bids", []),
                        ("hydra", ["initialize", "compose"])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, lambda *a_, **k_: None)
            sys.modules[name] = m
    # the reference package must win over this repo's drop-in mirror of the same name
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "meg_decoding" or k.startswith("meg_decoding.") or k == "constants"]:
        del sys.modules[k]
    import meg_decoding.models as M
    import meg_decoding.utils.loss as L
    assert M.__file__.startswith(REF), M.__file__
    return M, L


def subsample(a: np.ndarray, limit: int = 4096) -> np.ndarray:
    """Deterministic strided sample of a flat view (index set reproduced by tests via ``sample_index``)."""
    flat = a.reshape(-1)
    return flat[sample_index(flat.size, limit)]


def sample_index(n: int, limit: int = 4096) -> np.ndarray:
    if n <= limit:
        return np.arange(n)
    return (np.arange(limit, dtype=np.int64) * (n // limit)) + (np.arange(limit) % 7)


def run_reference(M, L, cfg, seed=0):
    loc = synth.sensor_locations(cfg.C)
    M.ch_locations_2d = lambda args: torch.from_numpy(loc)
    enc = M.BrainEncoder(cfg)
    lossf = L.CLIPLoss(cfg)
    state = synth.make_state(cfg, seed)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}
    missing = enc.load_state_dict(sd, strict=True)
    X, Y, subj = synth.make_batch(cfg, seed)
    X, Y = torch.from_numpy(X), torch.from_numpy(Y)
    subj_t = torch.from_numpy(subj)
    center = synth.drop_center(cfg)

    # the reference draws the centre from the global numpy RNG (models.py:236): force our index
    real_randint = np.random.randint
    np.random.randint = lambda *a, **k: center
    try:
        enc.train()
        Z = enc(X, subj_t)
    finally:
        np.random.randint = real_randint
    loss = lossf(Y, Z)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in enc.named_parameters()}
    grads["temp"] = lossf.temp.grad.detach().clone()
    bn = {k: v.detach().clone() for k, v in enc.state_dict().items() if "running_" in k or "num_batches" in k}
    enc.eval()
    with torch.no_grad():
        Z_eval = enc(X, subj_t)
        logits, loss_eval = lossf(Y, Z_eval, return_logits=True)
    return dict(Z=Z.detach(), loss=loss.detach(), grads=grads, bn=bn, Z_eval=Z_eval, loss_eval=loss_eval,
                logits_eval=logits, state=state, X=X, Y=Y, subj=subj_t, loc=loc, center=center)


def check_restatement(cfg, ref, tol=2e-5):
    P = R.state_to_torch(ref["state"], requires_grad=True)
    temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
    bufs = {k: v.clone() for k, v in P.items() if "running_" in k or "num_batches" in k}
    Z, loss, g = R.train_step(P, temp, ref["X"], ref["Y"], ref["subj"], ref["loc"], K=cfg.K, d_drop=cfg.d_drop,
                              seq2seq=cfg.seq2seq, drop_center=ref["center"], bn_buffers=bufs)

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))

    errs = {"Z": rel(Z, ref["Z"]), "loss": abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"]))}
    for k, gr in ref["grads"].items():
        if k.endswith(".z"):
            errs["g:" + k] = max(rel(g[k + "_re"], gr.real), rel(g[k + "_im"], gr.imag))
        elif gr is None:
            assert g[k] is None or float(g[k].abs().max()) == 0.0, k
        elif k == "temp":
            errs["g:temp"] = rel(g["temp"], gr)
        else:
            # conv biases directly in front of a train-mode BatchNorm have an exactly-zero true gradient;
            # what the reference reports there is fp32 rounding noise: compare against the weight-grad scale
            if re.fullmatch(r"conv_blocks\.conv\d\.conv[01]\.bias", k):
                scale = float(ref["grads"][k.replace(".bias", ".weight")].norm())
                assert float(gr.norm()) < 1e-4 * scale and float(g[k].norm()) < 1e-4 * scale, k
            else:
                errs["g:" + k] = rel(g[k], gr)
    for k, v in ref["bn"].items():
        if v.is_floating_point():
            errs["bn:" + k] = rel(bufs[k], v)
        else:
            assert int(bufs[k]) == int(v)
    Ze = R.encoder_forward(R.state_to_torch(ref["state"]) | bufs, ref["X"], ref["subj"], ref["loc"], K=cfg.K,
                           d_drop=cfg.d_drop, seq2seq=cfg.seq2seq, training=False)
    errs["Z_eval"] = rel(Ze, ref["Z_eval"])
    worst = max(errs.values())
    bad = {k: v for k, v in errs.items() if v > tol}
    # gradients pass through ~40 fp32 layers; allow them a looser (still tiny) bound
    bad = {k: v for k, v in bad.items() if not (k.startswith("g:") and v < 5e-4)}
    assert not bad, f"restatement != reference for {cfg.name}: {bad}"
    return worst


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(8)
    M, L = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in GOLDEN_CONFIGS:
        cfg = synth.get_config(name)
        ref = run_reference(M, L, cfg)
        worst = check_restatement(cfg, ref)
        out = {
            "loss": np.float64(ref["loss"]), "loss_eval": np.float64(ref["loss_eval"]),
            "Z_shape": np.array(ref["Z"].shape), "Z_norm": np.float64(ref["Z"].double().norm()),
            "Z_sample": subsample(ref["Z"].numpy()), "Z_eval_sample": subsample(ref["Z_eval"].numpy()),
            "Z_eval_norm": np.float64(ref["Z_eval"].double().norm()),
            "logits_eval_sample": subsample(ref["logits_eval"].numpy()),
            "center": np.int64(ref["center"]),
        }
        for k, g in ref["grads"].items():
            if g is None:
                out["gnone:" + k] = np.int64(1)
                continue
            if g.is_complex():
                out["gnorm:" + k] = np.float64(torch.view_as_real(g).double().norm())
                out["gsample:" + k] = subsample(torch.view_as_real(g).numpy())
            else:
                out["gnorm:" + k] = np.float64(g.double().norm())
                out["gsample:" + k] = subsample(g.numpy())
        for k, v in ref["bn"].items():
            out["bn:" + k] = v.numpy()
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={float(ref['loss']):.6f} restatement worst rel err={worst:.2e} -> {path} "
              f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
