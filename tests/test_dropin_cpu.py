"""CPU tier, boundary: the reference's training loop on the drop-in modules.

  * the hot loop of the reference's ``train.py`` (lines 185-221: forward, CLIPLoss, Classifier, ``loss.item()``,
    ``zero_grad / backward / step``) is cut out of the reference checkout AT TEST TIME and executed, unmodified, on the
    drop-in ``BrainEncoder`` / ``CLIPLoss`` / ``Classifier`` with ``torch.optim.Adam`` and a synthetic loader (needs
    /root/reference: skipped on the GPU box; arithmetic = the op restatement, no kernels);
  * with ``MEG_DECODING_REF`` set, the modules the drop-in does not provide resolve from the reference checkout while
    ``models`` / ``utils.loss`` / ``utils.layout`` keep resolving to the drop-in."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

import megb200
from megb200 import synth
from oracle.ops_restatement import TorchOps
from tests import helpers as H

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference checkout not present")


@needs_ref
@pytest.mark.parametrize("dataset", ["Gwilliams2022", "Brennan2018"])
def test_reference_train_loop_runs_unchanged_on_the_dropin(dataset):
    src = open(os.path.join(REF, "train.py")).read().splitlines()
    body = textwrap.dedent("\n".join(src[184:221]))              # train.py:185-221: brain_encoder.train() ... optimizer.step()
    assert body.lstrip().startswith("brain_encoder.train()") and "loss_func(Y, Z)" in body and "optimizer.step()" in body
    from meg_decoding.models import Classifier
    megb200._set_ops_override(TorchOps(torch.float32))
    try:
        cfg = synth.get_config("tiny", dataset=dataset, normalize_image_features=False)
        brain_encoder, loss_func = H.make_modules(cfg)
        classifier = Classifier(cfg)
        optimizer = torch.optim.Adam(list(brain_encoder.parameters()) + list(loss_func.parameters()), lr=3e-3)   # train.py:160
        batches = []
        for i in range(3):
            X, Y, subj = H.make_inputs(cfg, seed=i)
            batches.append((X, Y, subj) if dataset == "Gwilliams2022" else (X, Y, subj, torch.arange(cfg.B)))
        w0 = brain_encoder.conv_final2.weight.detach().clone()
        ns = dict(brain_encoder=brain_encoder, loss_func=loss_func, classifier=classifier, optimizer=optimizer,
                  train_loader=batches, tqdm=lambda x: x, device=torch.device("cpu"), torch=torch, args=cfg,
                  train_losses=[], trainTop1accs=[], trainTop10accs=[])
        exec(compile(body, "reference train.py:185-221", "exec"), ns)
        assert len(ns["train_losses"]) == 3 and all(torch.isfinite(torch.tensor(ns["train_losses"])))
        assert all(0.0 <= a <= 1.0 for a in ns["trainTop1accs"] + ns["trainTop10accs"])
        assert not torch.equal(brain_encoder.conv_final2.weight.detach(), w0), "the optimizer step changed nothing"
        if dataset == "Gwilliams2022":               # three optimizer steps on the same distribution: the loss moves down
            assert ns["train_losses"][-1] < ns["train_losses"][0] + 0.5
    finally:
        megb200._set_ops_override(None)


@needs_ref
def test_reference_submodules_resolve_behind_the_dropin():
    code = ("import meg_decoding, meg_decoding.utils, meg_decoding.models as M, meg_decoding.utils.loss as L;"
            "import meg_decoding.utils.reproducibility as R, speech_decoding.utils.reproducibility as R2;"
            "import speech_decoding.models as SM;"
            "print(M.__file__); print(L.__file__); print(R.__file__); print(R2.__file__); print(SM.BrainEncoder is M.BrainEncoder)")
    env = dict(os.environ, MEG_DECODING_REF=REF,
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "meg-decoding_b200"), REF]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    m, l, r, r2, same = out.stdout.strip().splitlines()[-5:]
    assert m.startswith(os.path.join(ROOT, "meg-decoding_b200")) and l.startswith(os.path.join(ROOT, "meg-decoding_b200"))
    assert r.startswith(REF) and r2.startswith(REF) and same == "True"
