"""CPU tier: the step orchestration (megb200.engine, the hand-derived backward chain, the drop-in nn.Module
surface) run on top of the oracle's plain-torch restatement of the ops, checked against the reference's golden
outputs.  No kernels run here; this pins the math and the module protocol the CUDA path is built on."""
import numpy as np
import pytest
import torch

import megb200
from megb200 import synth
from oracle.ops_restatement import TorchOps
from tests import helpers as H


@pytest.fixture()
def torch_ops():
    megb200._set_ops_override(TorchOps(torch.float32))
    yield
    megb200._set_ops_override(None)


@pytest.mark.parametrize("name", ["tiny", "tiny_static", "narrow_static"])
def test_orchestration_matches_reference_golden(torch_ops, name):
    cfg = synth.get_config(name)
    enc, lossf = H.make_modules(cfg)
    X, Y, subj = H.make_inputs(cfg)
    Z, loss, grads = H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
    bn = {k: v for k, v in enc.state_dict().items() if "running_" in k or "num_batches" in k}
    errs = H.compare_with_golden(name, Z, loss, grads, tol_z=2e-5, tol_loss=2e-5, tol_grad=1e-3, bn=bn, tol_bn=2e-5)
    print(name, errs)


def test_eval_mode_matches_reference_golden(torch_ops):
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg)
    X, Y, subj = H.make_inputs(cfg)
    # one train step first so that the running statistics are the ones the golden eval pass used
    H.train_step(enc, lossf, X, Y, subj, synth.drop_center(cfg))
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
        logits, loss = lossf(Y, Z, return_logits=True)
    g = H.load_golden("tiny")
    assert H.rel_fro(H.subsample(Z), g["Z_eval_sample"]) < 2e-5
    assert abs(float(loss) - float(g["loss_eval"])) / abs(float(g["loss_eval"])) < 2e-5
    assert H.rel_fro(H.subsample(logits), g["logits_eval_sample"]) < 2e-5


def test_state_dict_layout_matches_reference(torch_ops):
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg)
    keys = set(enc.state_dict().keys())
    assert keys == set(synth.make_state(cfg).keys())
    assert len(keys) == 87 + cfg.num_subjects                           # SURVEY §5 checkpoint row
    assert enc.state_dict()["subject_block.spatial_attention.z"].dtype == torch.complex64
    assert list(lossf.state_dict().keys()) == ["temp"]


def test_checkpoint_written_by_the_reference_on_cuda_loads(torch_ops):
    """On CUDA the reference's z is not a registered parameter (models.py:186-188), so its checkpoints lack the key:
    they must load (strict) with a warning and leave z untouched."""
    cfg = synth.get_config("tiny")
    enc, _ = H.make_modules(cfg, seed=1)
    sd = H.make_modules(cfg, seed=0)[0].state_dict()
    del sd["subject_block.spatial_attention.z"]
    z_before = enc.subject_block.spatial_attention.z.detach().clone()
    with pytest.warns(UserWarning, match="no 'subject_block.spatial_attention.z'"):
        res = enc.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(enc.subject_block.spatial_attention.z.detach(), z_before)
    assert torch.equal(enc.conv_final2.weight, sd["conv_final2.weight"])
    import warnings
    with warnings.catch_warnings():                  # a partial non-strict load is not a "reference checkpoint": silent
        warnings.simplefilter("error")
        enc.load_state_dict({"conv_final2.weight": sd["conv_final2.weight"]}, strict=False)


def test_subject_index_validation():
    from megb200.engine import validate_subject_idxs
    s = torch.tensor([0, 3, -1, -4], dtype=torch.int32)
    assert validate_subject_idxs(s, 4).tolist() == [0, 3, 3, 0]           # ModuleList wraps negatives
    for bad in ([0, 4], [-5, 1]):
        with pytest.raises(IndexError):
            validate_subject_idxs(torch.tensor(bad), 4)
    with pytest.raises(TypeError):
        validate_subject_idxs(torch.tensor([0.0, 1.0]), 4)


def test_reference_error_conventions(torch_ops):
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg)
    X, Y, subj = H.make_inputs(cfg)
    with pytest.raises(AssertionError):                                  # models.py:233
        enc(X[:, :-1], subj)
    with pytest.raises(IndexError):                                      # ModuleList index out of range
        enc(X, torch.full_like(subj, cfg.num_subjects))
    with pytest.raises(AssertionError):                                  # loss.py:68
        lossf(Y[:1], torch.zeros(1, Y.shape[1], Y.shape[2]))


def test_subject_routing_is_permutation_equivariant(torch_ops):
    cfg = synth.get_config("tiny")
    enc, _ = H.make_modules(cfg)
    X, _, subj = H.make_inputs(cfg)
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
        perm = torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(3))
        Zp = enc(X[perm], subj[perm])
    assert torch.equal(Zp, Z[perm])
    # int64 indices (Brennan's default collate) behave like int32
    with torch.no_grad():
        assert torch.equal(enc(X, subj.long()), Z)


def test_absent_subjects_get_no_gradient(torch_ops):
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg)
    X, Y, subj = H.make_inputs(cfg)
    subj = torch.zeros_like(subj)
    subj[1] = 2
    _, _, grads = H.train_step(enc, lossf, X, Y, subj, 0)
    assert grads["subject_block.subject_layer.0.weight"] is not None
    assert grads["subject_block.subject_layer.1.weight"] is None
    assert grads["subject_block.subject_layer.3.weight"] is None


def test_one_sided_dp_loss_equals_global(torch_ops):
    """SURVEY appendix A6 with emulated ranks: per-rank contributions sum to the single-process loss."""
    from oracle import restatement as R
    torch.manual_seed(0)
    Bg, D, world = 12, 40, 3
    Y, Z = torch.randn(Bg, D), torch.randn(Bg, D)
    temp = torch.tensor([1.3])
    full = R.clip_loss(Y, Z, temp)
    parts = R.clip_loss_one_sided(Y, list(Z.chunk(world)), temp)
    assert abs(float(sum(parts)) - float(full)) < 1e-6


def test_classifier_matches_oracle_topk(torch_ops):
    """N1: Classifier (top-1 / top-10 retrieval accuracy, models.py:403-460) on the logits + rank path."""
    from meg_decoding.models import Classifier
    from oracle import restatement as R
    g = torch.Generator().manual_seed(5)
    B, Fo, T = 24, 16, 6
    Y = torch.randn(B, Fo, T, generator=g)
    Z = 0.6 * Y + torch.randn(B, Fo, T, generator=g)          # correlated: accuracies strictly between 0 and 1
    clf = Classifier(synth.Args(normalize_image_features=False))
    top1, top10 = clf(Z, Y)
    r1, r10 = R.topk_accuracy(Z, Y, k=10)
    assert abs(top1 - r1) < 1e-6 and abs(top10 - r10) < 1e-6
    assert 0.0 < top1 <= top10 <= 1.0
    t1, t10, t3 = clf(Z, Y, top_k=3)
    assert abs(t3 - R.topk_accuracy(Z, Y, k=3)[1]) < 1e-6


def test_async_scalar_reader_returns_every_value_in_order():
    from megb200.data import AsyncScalarReader
    for lag in (0, 1, 3):
        r = AsyncScalarReader("cpu", lag=lag)
        got = []
        for i in range(7):
            v = r.push(torch.tensor(float(i) + 0.5))
            assert (v is None) == (i < lag)
            if v is not None:
                got.append(v)
        got += r.drain()
        assert got == [i + 0.5 for i in range(7)]


def test_packed_and_bf16_targets_equal_dense_f32(torch_ops):
    """CLIPLoss accepts the targets dense in f32 (the reference's collator), dense in bf16, or pre-packed
    (megb200.data.PackedTargets: position-major rows + norms).  Packed must reproduce dense f32 exactly (same rounding
    point, same norms); on this f32 op restatement that means bit for bit."""
    from megb200.data import PackedTargets
    cfg = synth.get_config("tiny")
    enc, lossf = H.make_modules(cfg)
    X, Y, subj = H.make_inputs(cfg)
    enc.eval()
    with torch.no_grad():
        Z = enc(X, subj)
        dense = lossf(Y, Z)
        packed = lossf(PackedTargets.from_dense(Y, dtype=torch.float32), Z)
        half = lossf(Y.to(torch.bfloat16), Z)
    assert torch.equal(dense, packed)
    assert abs(float(half) - float(dense)) / abs(float(dense)) < 5e-3
    pt = PackedTargets.from_dense(Y, dtype=torch.float32)
    assert pt.size(0) == cfg.B and pt.shape == tuple(Y.shape) and pt[[1, 0]].rows.shape == (2, pt.rows.shape[1])
    # static (B, F) targets
    cfg2 = synth.get_config("tiny_static")
    enc2, lossf2 = H.make_modules(cfg2)
    X2, Y2, subj2 = H.make_inputs(cfg2)
    enc2.eval()
    with torch.no_grad():
        Z2 = enc2(X2, subj2)
        assert torch.equal(lossf2(Y2, Z2), lossf2(PackedTargets.from_dense(Y2, dtype=torch.float32), Z2))


def test_device_target_store_gathers_packed_rows():
    """megb200.data.DeviceTargetStore (the label-side companion of DeviceSegmentStore): a batch is a row gather out of the
    resident packed targets; same loss as the dense collated batch."""
    from megb200.data import DeviceTargetStore, PackedTargets
    g = torch.Generator().manual_seed(3)
    Yall = torch.randn(10, 16, 5, generator=g)
    store = DeviceTargetStore(Yall, device="cpu", dtype=torch.float32, chunk=4)
    assert len(store) == 10
    ids = torch.tensor([7, 0, 3, -1])
    got = store.gather(ids)
    ref = PackedTargets.from_dense(Yall[[7, 0, 3, 9]], torch.float32)
    assert torch.equal(got.rows, ref.rows) and torch.allclose(got.sumsq, ref.sumsq) and got.shape == (4, 16, 5)
    with pytest.raises(IndexError):
        store.gather([10])
