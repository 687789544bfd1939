"""GPU tier, path level: the drop-in modules (public API -> C-ABI -> sm_100a kernels) against

  * the committed golden fixtures = outputs of the UNMODIFIED reference on the same seeded inputs and weights
    (tests/golden/*.npz, made by oracle/make_golden.py);
  * the oracle restatement run on the host CPU at mid size;
  * size-independent properties at the full BASELINE.json shape (batch-permutation equivariance, bit-exact subject
    routing, agreement of the tensor-core path with the fp32 check mode).

Tolerances are north_star's: loss and embeddings within 1e-3 relative in the fp32 check mode, 2e-2 in bf16; subject
and segment indexing bit-exact.  Metrics: ||Z - Z_ref||_F / ||Z_ref||_F and |l - l_ref| / |l_ref|.
"""
import numpy as np
import pytest
import torch

import megb200
from megb200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# precision mode -> (Z tol, loss tol, grad tol)
MODES = {
    "fp32": (1e-3, 1e-3, 5e-3),
    "bf16": (2e-2, 2e-2, 0.35),          # tcgen05 engine (the product path)
    "bf16_simt": (2e-2, 2e-2, 0.35),     # same storage on the fp32 CUDA-core engine (cross-check)
}


@pytest.fixture(params=list(MODES))
def mode(request):
    old = megb200.get_precision()
    megb200.set_precision(request.param)
    yield request.param
    megb200.set_precision(old)


def test_native_library_is_loaded():
    from megb200 import lib
    assert lib.load().megb200_version() == 100
    maps = open("/proc/self/maps").read()
    assert "libmegb200.so" in maps


@pytest.mark.parametrize("name", ["tiny", "tiny_static", "narrow", "narrow_static"])
def test_train_step_matches_reference_golden(mode, name):
    tz, tl, tg = MODES[mode]
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
    torch.cuda.synchronize()
    assert Z.is_cuda and loss.is_cuda
    bn = {k: v for k, v in enc.state_dict().items() if "running_" in k or "num_batches" in k}
    errs = H.compare_with_golden(name, Z, loss, grads, tz, tl, tg, bn=bn, tol_bn=5e-3 if mode == "fp32" else 3e-2)
    print(f"[parity] {name} {mode}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))


@pytest.mark.parametrize("name", ["tiny_static", "narrow_static"])
def test_fused_clip_train_step_matches_reference_golden(name):
    """The never-materialised CLIP form (the GOD sweep's large-batch path) forced on the static golden configs: the
    whole train step against the reference's stored outputs, bf16 tolerances."""
    from megb200 import lib
    tz, tl, tg = MODES["bf16"]
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    with lib.options(clip_fused=1):
        Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
        torch.cuda.synchronize()
    errs = H.compare_with_golden(name, Z, loss, grads, tz, tl, tg)
    print(f"[parity] {name} bf16 fused CLIP: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))


@pytest.mark.parametrize("name", ["tiny_pooled", "narrow_pooled", "tiny_dilated", "narrow_dilated"])
def test_encoder_variants_match_reference_golden(mode, name):
    """"Next" row N4: BrainEncoderSeq2Static (kernel sizes 11/5/3, AvgPool1d(3,2) between blocks, global average; models.py:
    465-512) and the dilated paper block (dilations up to 2048 > T; tests/test_modules/models.py:267-312) against the
    UNMODIFIED reference's stored outputs (oracle/make_golden_variants.py)."""
    tz, tl, tg = MODES[mode]
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
    torch.cuda.synchronize()
    errs = H.compare_with_golden(name, Z, loss, grads, tz, tl, tg)
    print(f"[parity] {name} {mode}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))


def test_classifier_matches_reference_golden(mode):
    """"Next" row N1: the drop-in Classifier (tcgen05 logits + rank kernel) against the reference Classifier's stored
    (top1, top10, top5) on the seeded cases of tests/golden/classifier.npz.  bf16 may flip one near-tie per case."""
    from meg_decoding.models import Classifier
    clf = Classifier(synth.Args(normalize_image_features=False))
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    for seed, B, Fo, T, top1, top10, top5 in H.load_golden("classifier")["cases"]:
        B = int(B)
        g = torch.Generator().manual_seed(100 + int(seed))
        shape = (B, int(Fo), int(T)) if T > 1 else (B, int(Fo))
        Y = torch.randn(*shape, generator=g)
        Z = 0.25 * Y + torch.randn(*shape, generator=g)
        if shape[1] % 8 != 0:
            continue                                     # the kernels need F % 8 == 0 (as BrainEncoder does)
        r1, r10, r5 = clf(Z.to(DEV).to(dt), Y.to(DEV), top_k=5)
        slack = 1e-6 if mode == "fp32" else 1.0 / B + 1e-6
        print(f"[parity] classifier {mode} seed {int(seed)}: top1 {r1:.4f}/{top1:.4f} top10 {r10:.4f}/{top10:.4f} top5 {r5:.4f}/{top5:.4f}")
        assert abs(r1 - top1) <= slack and abs(r10 - top10) <= slack and abs(r5 - top5) <= slack


def test_eval_mode_matches_reference_golden(mode):
    tz, tl, _ = MODES[mode]
    cfg = synth.get_config("narrow")
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
        logits, loss = lossf(Y, Z, return_logits=True)
    g = H.load_golden("narrow")
    assert H.rel_fro(H.subsample(Z), g["Z_eval_sample"]) <= tz
    assert abs(float(loss) - float(g["loss_eval"])) / abs(float(g["loss_eval"])) <= tl
    assert H.rel_fro(H.subsample(logits), g["logits_eval_sample"]) <= 5 * tz


def test_mid_size_against_cpu_oracle(mode):
    """Brennan2018 shape (60 channels, 33 subjects, T=360) at B=16: oracle on the host CPU, same inputs."""
    from oracle import restatement as R
    tz, tl, _ = MODES[mode]
    cfg = synth.get_config("brennan", B=16)
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    center = synth.drop_center(cfg)
    Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, center)
    P = R.state_to_torch(synth.make_state(cfg), requires_grad=True)
    temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
    Zr, lr, gr = R.train_step(P, temp, X.cpu(), Y.cpu(), subj, synth.sensor_locations(cfg.C), K=cfg.K,
                              d_drop=cfg.d_drop, seq2seq=True, drop_center=center)
    ez, el = H.rel_fro(Z.float().cpu(), Zr), abs(float(loss) - float(lr)) / abs(float(lr))
    gt = abs(float(grads["temp"]) - float(gr["temp"])) / abs(float(gr["temp"]))
    gw = H.rel_fro(grads["conv_final2.weight"].cpu(), gr["conv_final2.weight"])
    print(f"[parity] brennan B=16 {mode}: Z={ez:.2e} loss={el:.2e} dtemp={gt:.2e} dWf2={gw:.2e}")
    assert ez <= tz and el <= tl
    assert gt <= (5e-3 if mode == "fp32" else 5e-2)
    assert gw <= (5e-3 if mode == "fp32" else 5e-2)


def test_subject_routing_bit_exact_and_permutation_equivariant(mode):
    cfg = synth.get_config("narrow")
    enc, _ = H.make_modules(cfg, DEV)
    X, _, _ = H.make_inputs(cfg, DEV, B=12)
    subj = torch.tensor([0, 1, 2, 3, 4, 4, 3, 2, 1, 0, 2, 2], dtype=torch.int32)
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
        perm = torch.randperm(12, generator=torch.Generator().manual_seed(5))
        Zp = enc(X[perm.to(DEV)], subj[perm])
        Z64 = enc(X, subj.long())
        Zdev = enc(X, subj.to(DEV))
    assert torch.equal(Zp, Z[perm.to(DEV)]), "batch permutation must permute the outputs bit-exactly"
    assert torch.equal(Z64, Z) and torch.equal(Zdev, Z), "int64 / device-resident indices must route identically"
    # one-hot routing: give two samples identical inputs but different subjects -> different outputs; same subject -> equal
    X2 = X.clone()
    X2[1] = X2[0]
    with torch.no_grad():
        Za = enc(X2, torch.tensor([3, 3] + [0] * 10, dtype=torch.int32))
        Zb = enc(X2, torch.tensor([3, 4] + [0] * 10, dtype=torch.int32))
    assert torch.equal(Za[0], Za[1])
    assert not torch.equal(Zb[0], Zb[1])
    # device-resident indices are range-checked too (the grouped kernels index weights with them): ModuleList's IndexError
    S = cfg.num_subjects
    for bad in (S, -S - 1):
        with pytest.raises(IndexError):
            enc(X, torch.full((12,), bad, dtype=torch.int32, device=DEV))
    with torch.no_grad():
        Zneg = enc(X, (subj - S).to(DEV))                # negative indices wrap like ModuleList[-k]
    assert torch.equal(Zneg, Z)


def test_full_size_tensor_core_path_agrees_with_check_mode():
    """BASELINE.json configs[1] shape at B=32: bf16/tcgen05 path vs the fp32 check mode on the same device,
    plus a checksum-style invariant: every logits row/column softmax sums to one (sum of dL over a row is 0)."""
    cfg = synth.get_config("gwilliams", B=32)
    res = {}
    for mode in ("fp32", "bf16"):
        megb200.set_precision(mode)
        enc, lossf = H.make_modules(cfg, DEV)
        X, Y, subj = H.make_inputs(cfg, DEV)
        Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
        res[mode] = (Z.float(), float(loss), grads)
    megb200.set_precision("bf16")
    ez = H.rel_fro(res["bf16"][0], res["fp32"][0])
    el = abs(res["bf16"][1] - res["fp32"][1]) / abs(res["fp32"][1])
    print(f"[parity] gwilliams B=32 bf16 vs fp32 check mode: Z={ez:.2e} loss={el:.2e}")
    assert ez <= 2e-2 and el <= 2e-2
    for k in ("conv_final2.weight", "conv_blocks.conv2.conv1.weight", "subject_block.conv.weight", "temp"):
        e = H.rel_fro(res["bf16"][2][k].float(), res["fp32"][2][k].float())
        print(f"[parity]   grad {k}: {e:.2e}")
        assert e <= 0.2, k


# ------------------------------------------------------------------------------------------------ config 2, full size
_ORACLE_CACHE = {}


def _oracle_full(name, B, seed):
    """CPU oracle (fp32, all host threads) of one training step at the full BASELINE.json shape, with the per-layer
    activations; cached so that the bf16 and the fp32 runs of one seed share it."""
    from oracle import restatement as R
    key = (name, B, seed)
    if key not in _ORACLE_CACHE:
        cfg = synth.get_config(name, B=B)
        P = R.state_to_torch(synth.make_state(cfg, seed), requires_grad=True)
        temp = torch.tensor([float(cfg.init_temperature)], requires_grad=True)
        X, Y, subj = synth.make_batch(cfg, seed)
        X, Y, subj = torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(subj)
        Z, inter = R.encoder_forward(P, X, subj, synth.sensor_locations(cfg.C), K=cfg.K, d_drop=cfg.d_drop,
                                     seq2seq=cfg.seq2seq, training=True, drop_center=synth.drop_center(cfg),
                                     return_intermediates=True)
        loss = R.clip_loss(Y, Z, temp)
        names = ["conv_final2.weight", "conv_blocks.conv2.conv1.weight", "conv_blocks.conv0.conv0.weight",
                 "subject_block.conv.weight"]
        grads = torch.autograd.grad(loss, [P[n] for n in names] + [temp])
        _ORACLE_CACHE.clear()            # one seed resident at a time (~1.5 GB of host memory each)
        _ORACLE_CACHE[key] = dict(
            Z=Z.detach(), loss=float(loss.detach()), inter={k: v.detach() for k, v in inter.items() if k != "sa"},
            grads={n: g for n, g in zip(names + ["temp"], grads)})
    return _ORACLE_CACHE[key]


def _layer_errors(enc, X, subj, center, ref):
    """Per-layer rel-Frobenius errors of one more forward through the engine (position-major activations -> NCW)."""
    sa = enc.subject_block.spatial_attention
    eng = enc._engine()
    with torch.no_grad():
        params = enc._param_list()
        _, ctx = eng.forward(enc._param_dict(params), X, subj.to(X.device).to(torch.int32), sa.drop_masks[center], True)
    out = {"subject": H.rel_fro(ctx["H2"][:, :, :enc.D1].float().permute(0, 2, 1).cpu(), ref["subject"])}
    for k in range(5):
        t = ctx["blocks"][k + 1]["x"] if k < 4 else ctx["x5"]
        out[f"block{k}"] = H.rel_fro(t[:, :, :enc.D2].float().permute(0, 2, 1).cpu(), ref[f"block{k}"])
    return out


@pytest.mark.parametrize("prec,seed", [("bf16", 0), ("fp32", 0), ("bf16", 1), ("bf16", 2)])
def test_config2_full_batch_against_cpu_oracle(prec, seed):
    """BASELINE.json configs[1] at its real size (B=256, C=208, T=360, 27 subjects): the kernels that run here are
    the product kernels (persistent CTA-pair tap-GEMMs, one-wave weight-gradient splits, B=256 CLIP forms).  Loss,
    embeddings, per-layer activations and a few gradients against the CPU oracle on the same seeded inputs; three
    seeds in bf16 because the 2e-2 bar leaves little margin (the margin is printed)."""
    tz, tl, _ = MODES[prec]
    ref = _oracle_full("gwilliams", 256, seed)
    old = megb200.get_precision()
    megb200.set_precision(prec)
    try:
        cfg = synth.get_config("gwilliams", B=256)
        enc, lossf = H.make_modules(cfg, DEV, seed=seed)
        X, Y, subj = H.make_inputs(cfg, DEV, seed=seed)
        center = synth.drop_center(cfg)
        Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, center)
        torch.cuda.synchronize()
        ez = H.rel_fro(Z.float().cpu(), ref["Z"])
        el = abs(float(loss) - ref["loss"]) / abs(ref["loss"])
        layers = _layer_errors(enc, X, subj, center, ref["inter"])
        gerr = {k: H.rel_fro((torch.view_as_real(grads[k]) if grads[k].is_complex() else grads[k]).float().cpu(), g)
                for k, g in ref["grads"].items()}
    finally:
        megb200.set_precision(old)
    print(f"[parity] gwilliams B=256 {prec} seed {seed}: Z={ez:.3e} (bar {tz:.0e}, margin {100 * (1 - ez / tz):.0f} %) "
          f"loss={el:.2e} | " + " ".join(f"{k}={v:.2e}" for k, v in layers.items()) + " | grads " +
          " ".join(f"{k.split('.')[-2] if '.' in k else k}={v:.2e}" for k, v in gerr.items()))
    assert el <= tl, f"loss rel error {el:.3e} > {tl}"
    for k, v in layers.items():          # a failure bisects to the first layer that leaves the budget
        assert v <= tz, f"layer {k}: rel-Fro {v:.3e} > {tz}"
    assert ez <= tz, f"Z rel-Fro {ez:.3e} > {tz}"
    gt = 5e-3 if prec == "fp32" else 0.2
    for k, v in gerr.items():
        assert v <= (gt if k != "temp" else (5e-3 if prec == "fp32" else 5e-2)), f"grad {k}: {v:.3e}"
    if prec == "bf16":
        from megb200 import lib
        assert lib.get_option("last_wgrad_splits") >= 1 and lib.get_option("last_clip_dz_kernel") == 0


@pytest.mark.parametrize("name", ["narrow", "narrow_static"])
def test_packed_and_bf16_targets_match_dense_f32(name):
    """Data formats on the target side of the path: dense f32 (reference collator), dense bf16, PackedTargets.  On the
    SAME embeddings the packed form must give the dense-f32 loss and gradients (same rounding point; the norms differ only
    in f32 summation order); bf16 targets differ by their own rounding of the norm."""
    from megb200.data import PackedTargets
    from megb200.graph import GraphedTrainStep
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg, DEV)
    X, Y, subj = H.make_inputs(cfg, DEV)
    center = synth.drop_center(cfg)
    enc.eval()
    with torch.no_grad():
        Z0 = enc(X, subj)
    res = {}
    for kind, tgt in (("dense", Y), ("packed", PackedTargets.from_dense(Y.cpu()).to(DEV)), ("bf16", Y.to(torch.bfloat16))):
        Z = Z0.clone().requires_grad_(True)
        lossf.temp.grad = None
        loss = lossf(tgt, Z)
        loss.backward()
        res[kind] = (float(loss), Z.grad.float().clone(), float(lossf.temp.grad))
    l0, g0, t0 = res["dense"]
    for kind, tl, tg in (("packed", 1e-5, 1e-3), ("bf16", 5e-3, 5e-2)):
        l, g, t = res[kind]
        print(f"[parity] targets {kind} vs dense f32 ({name}): loss {abs(l - l0) / abs(l0):.2e} dZ {H.rel_fro(g, g0):.2e} "
              f"dtemp {abs(t - t0) / abs(t0):.2e}")
        assert abs(l - l0) / abs(l0) <= tl and H.rel_fro(g, g0) <= tg and abs(t - t0) / abs(t0) <= max(tg, 1e-3)
    # the whole step through the public modules and through the CUDA-graph step on packed targets
    _, lp, _ = H.train_step(enc, lossf, X, PackedTargets.from_dense(Y.cpu()).to(DEV), subj, center)
    _, ld, _ = H.train_step(enc, lossf, X, Y, subj, center)
    assert abs(float(lp) - float(ld)) / abs(float(ld)) <= 1e-2           # run-to-run bf16 / atomic-order noise only
    gstep = GraphedTrainStep(enc, lossf, X, PackedTargets.from_dense(Y.cpu()).to(DEV), subj, drop_center=center)
    gl = float(gstep())
    assert abs(gl - float(ld)) / abs(float(ld)) <= 1e-2


def test_bf16_is_the_default_precision_and_output_is_bf16():
    assert megb200.get_precision() == "bf16"
    cfg = synth.get_config("tiny")
    enc, _ = H.make_modules(cfg, DEV)
    X, _, subj = H.make_inputs(cfg, DEV)
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
    assert Z.dtype == torch.bfloat16 and Z.shape == (cfg.B, cfg.F, cfg.T)


def test_output_dtype_float32_for_numpy_callers():
    """evaluate.py:340 / examples/inference.py:105 do ``Z.detach().cpu().numpy()``: ``args.output_dtype = "float32"`` makes
    the no-grad forward return f32 (one cast pass), same values as the bf16 output."""
    cfg = synth.get_config("tiny", output_dtype="float32")
    enc, _ = H.make_modules(cfg, DEV)
    X, _, subj = H.make_inputs(cfg, DEV)
    enc.eval()
    with torch.no_grad():
        Z32 = enc(X, subj)
        enc.output_dtype = None
        Z16 = enc(X, subj)
    assert Z32.dtype == torch.float32 and Z16.dtype == torch.bfloat16
    assert torch.equal(Z32, Z16.float())
    assert Z32.detach().cpu().numpy().shape == (cfg.B, cfg.F, cfg.T)


def test_classifier_topk_matches_oracle(mode):
    """N1: the retrieval metric (models.py:403-460) from the tcgen05 logits + the rank kernel; integer ranks."""
    from meg_decoding.models import Classifier
    from oracle import restatement as R
    g = torch.Generator().manual_seed(9)
    B, Fo, T = 64, 64, 40
    Y = torch.randn(B, Fo, T, generator=g)
    Z = 0.03 * Y + torch.randn(B, Fo, T, generator=g)       # weak coupling: accuracies strictly inside (0, 1)
    clf = Classifier(synth.Args(normalize_image_features=False))
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    Zd = Z.to(DEV).to(dt)
    top1, top10 = clf(Zd, Y.to(DEV))
    r1, r10 = R.topk_accuracy(Zd.float().cpu(), Y, k=10)
    print(f"[parity] classifier {mode}: top1 {top1:.3f} (oracle {r1:.3f}) top10 {top10:.3f} (oracle {r10:.3f})")
    assert abs(top1 - r1) <= 1.0 / B + 1e-6 and abs(top10 - r10) <= 1.0 / B + 1e-6   # near-ties may flip one rank in bf16
