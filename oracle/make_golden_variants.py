"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden fixtures for the "next" rows N1 / N4 from the UNMODIFIED reference:

  * ``BrainEncoderSeq2Static`` (meg_decoding/models.py:465-512): per-block kernel sizes ``args.ConvBlocks.ks``,
    ``AvgPool1d(3, 2)`` between the blocks, ``AdaptiveAvgPool1d(1)`` after the last -> tests/golden/{tiny,narrow}_pooled.npz
  * the dilated paper block ``ConvBlockTest`` (tests/test_modules/models.py:267-312) stacked into the reference
    ``BrainEncoder`` in place of its ``ConvBlock``s -> tests/golden/{tiny,narrow}_dilated.npz
  * ``Classifier`` (meg_decoding/models.py:386-460), the top-1 / top-10 / top-k retrieval accuracies
    -> tests/golden/classifier.npz (and the assertion that oracle.restatement.topk_accuracy reproduces it)

Same recipe as oracle/make_golden.py (import shims, synthetic layout, deterministic weights loaded with load_state_dict);
runs only in the build container.   python oracle/make_golden_variants.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "meg-decoding_b200"))
sys.path.insert(0, ROOT)

from megb200 import synth  # noqa: E402
from oracle import make_golden as MG  # noqa: E402
from oracle import restatement as R  # noqa: E402


class _NS(dict):
    __getattr__ = dict.__getitem__


def build_reference_encoder(M, cfg):
    """The reference module of a variant config, weights from synth.make_state."""
    loc = synth.sensor_locations(cfg.C)
    M.ch_locations_2d = lambda args: torch.from_numpy(loc)
    if cfg.get("model") == "brain_endcoder_seq2static":
        a = synth.Args(cfg)
        a["ConvBlocks"] = _NS(ks=list(cfg["ConvBlocks"]["ks"]))
        enc = M.get_model(a)
        assert type(enc).__name__ == "BrainEncoderSeq2Static"
    else:
        enc = M.BrainEncoder(cfg)
        if cfg.get("conv_dilations") == "paper":
            # the reference keeps the dilated block as an alternative formulation under tests/test_modules (its imports
            # name the package by its old name): stack it into the encoder in place of ConvBlock, parameter for parameter
            sys.modules.setdefault("speech_decoding", sys.modules["meg_decoding"])
            sys.modules.setdefault("speech_decoding.models", M)
            sys.modules.setdefault("speech_decoding.utils", sys.modules["meg_decoding.utils"])
            sys.modules.setdefault("speech_decoding.utils.layout", sys.modules["meg_decoding.utils.layout"])
            sys.path.insert(0, os.path.join(MG.REF, "tests"))
            import test_modules.models as TM
            blocks = torch.nn.Sequential()
            for k in range(1, 6):
                blk = TM.ConvBlockTest(k, cfg.D1, cfg.D2)
                dil = [blk.conv1.dilation[0], blk.conv2.dilation[0], blk.conv3.dilation[0]]
                assert dil == list(R.paper_dilations()[k - 1]), dil
                blocks.add_module(f"conv{k - 1}", blk)
            enc.conv_blocks = blocks
    state = synth.make_state(cfg)
    sd = {}
    for k, v in state.items():
        if cfg.get("conv_dilations") == "paper" and k.startswith("conv_blocks."):
            # ConvBlockTest names its layers conv1..3 / batchnorm1..2 (ConvBlock: conv0..2 / batchnorm0..1)
            head, name, rest = k.split(".", 2)[0:3]
            part, tail = rest.split(".", 1)
            num = int(part[-1]) + 1
            k2 = f"{head}.{name}.{part[:-1]}{num}.{tail}"
            sd[k2] = torch.from_numpy(np.asarray(v))
        else:
            sd[k] = torch.from_numpy(np.asarray(v))
    enc.load_state_dict(sd, strict=True)
    return enc, state, loc


def run_variant(M, L, cfg):
    enc, state, loc = build_reference_encoder(M, cfg)
    lossf = L.CLIPLoss(cfg)
    X, Y, subj = synth.make_batch(cfg, 0)
    X, Y, subj_t = torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(subj)
    center = synth.drop_center(cfg)
    real = np.random.randint
    np.random.randint = lambda *a, **k: center
    try:
        enc.train()
        Z = enc(X, subj_t)
    finally:
        np.random.randint = real
    loss = lossf(Y, Z)
    loss.backward()
    grads = {}
    for k, p in enc.named_parameters():
        if cfg.get("conv_dilations") == "paper" and k.startswith("conv_blocks."):
            head, name, rest = k.split(".", 2)
            part, tail = rest.split(".", 1)
            k = f"{head}.{name}.{part[:-1]}{int(part[-1]) - 1}.{tail}"
        grads[k] = p.grad.detach().clone() if p.grad is not None else None
    grads["temp"] = lossf.temp.grad.detach().clone()
    return dict(Z=Z.detach(), loss=loss.detach(), grads=grads, state=state, X=X, Y=Y, subj=subj_t, loc=loc, center=center)


def check_variant(cfg, ref, tol=2e-5):
    P = R.state_to_torch(ref["state"], requires_grad=True)
    temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
    pooled = cfg.get("model") == "brain_endcoder_seq2static"
    Z, loss, g = R.train_step(P, temp, ref["X"], ref["Y"], ref["subj"], ref["loc"], K=cfg.K, d_drop=cfg.d_drop,
                              seq2seq=cfg.seq2seq, drop_center=ref["center"], dilations=synth.conv_dilations(cfg),
                              ks=synth.conv_kernel_sizes(cfg), pooled=pooled)

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))

    errs = {"Z": rel(Z, ref["Z"]), "loss": abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"]))}
    import re
    for k, gr in ref["grads"].items():
        if gr is None or re.fullmatch(r"conv_blocks\.conv\d\.conv[01]\.bias", k):
            continue
        if k.endswith(".z"):
            errs["g:" + k] = max(rel(g[k + "_re"], gr.real), rel(g[k + "_im"], gr.imag))
        else:
            errs["g:" + k] = rel(g[k], gr)
    bad = {k: v for k, v in errs.items() if v > (5e-4 if k.startswith("g:") else tol)}
    assert not bad, f"restatement != reference for {cfg.name}: {bad}"
    return max(errs.values())


def classifier_golden(M):
    """Reference Classifier on seeded (Z, Y) pairs with a weak coupling so that the accuracies sit strictly inside (0, 1)."""
    import constants
    clf = M.Classifier(synth.Args(normalize_image_features=False))
    out = {"cases": []}
    rows = []
    for seed, (B, Fo, T) in enumerate([(24, 8, 6), (24, 8, 6), (40, 16, 1), (12, 4, 3), (64, 8, 4), (11, 8, 2)]):
        g = torch.Generator().manual_seed(100 + seed)
        shape = (B, Fo, T) if T > 1 else (B, Fo)
        Y = torch.randn(*shape, generator=g)
        Z = 0.25 * Y + torch.randn(*shape, generator=g)
        top1, top10, top5 = clf(Z.to(constants.device), Y.to(constants.device), top_k=5)
        r1, r10 = R.topk_accuracy(Z, Y, k=10)
        _, r5 = R.topk_accuracy(Z, Y, k=5)
        assert abs(r1 - top1) < 1e-6 and abs(r10 - top10) < 1e-6 and abs(r5 - top5) < 1e-6, (seed, top1, r1, top10, r10, top5, r5)
        rows.append([seed, B, Fo, T, top1, top10, top5])
        print(f"classifier seed {seed} {shape}: top1 {top1:.4f} top10 {top10:.4f} top5 {top5:.4f} (restatement equal)")
    return np.array(rows, dtype=np.float64)


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(8)
    M, L = MG.import_reference()
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in ["tiny_pooled", "narrow_pooled", "tiny_dilated", "narrow_dilated"]:
        cfg = synth.get_config(name)
        ref = run_variant(M, L, cfg)
        worst = check_variant(cfg, ref)
        out = {"loss": np.float64(ref["loss"]), "Z_shape": np.array(ref["Z"].shape),
               "Z_norm": np.float64(ref["Z"].double().norm()), "Z_sample": MG.subsample(ref["Z"].numpy()),
               "center": np.int64(ref["center"])}
        for k, g in ref["grads"].items():
            if g is None:
                out["gnone:" + k] = np.int64(1)
            elif g.is_complex():
                out["gnorm:" + k] = np.float64(torch.view_as_real(g).double().norm())
                out["gsample:" + k] = MG.subsample(torch.view_as_real(g).numpy())
            else:
                out["gnorm:" + k] = np.float64(g.double().norm())
                out["gsample:" + k] = MG.subsample(g.numpy())
        path = os.path.join(gdir, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: Z {tuple(ref['Z'].shape)} loss={float(ref['loss']):.6f} restatement worst rel err={worst:.2e} -> {path} "
              f"({os.path.getsize(path) / 1024:.0f} KiB)")
    np.savez_compressed(os.path.join(gdir, "classifier.npz"), cases=classifier_golden(M))


if __name__ == "__main__":
    main()
