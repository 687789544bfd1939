"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for the "next" row N2 from the UNMODIFIED reference functions.

Imports ``baseline_correction_single`` and ``scaleAndClamp`` from /root/reference (stubbing termcolor / omegaconf, which
the two functions never touch), runs them on seeded inputs that include the edge cases (constant channel = zero IQR,
ties, odd and even T, values far outside the clamp), asserts that oracle/collator_restatement.py reproduces them, and
writes tests/golden/collator.npz.

    python oracle/make_golden_collator.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import collator_restatement as CR  # noqa: E402

REF = "/root/reference"


def import_reference():
    for name, attrs in (("termcolor", ["cprint"]), ("omegaconf", ["open_dict"])):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                for a in attrs:
                    setattr(m, a, lambda *a_, **k_: None)
                sys.modules[name] = m
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_preproc_utils", os.path.join(REF, "meg_decoding/utils/preproc_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_case(seed, B, C, T):
    g = np.random.Generator(np.random.PCG64(seed))
    X = (g.standard_normal((B, C, T)) * g.uniform(0.1, 30.0, size=(B, C, 1)) + g.uniform(-5, 5, size=(B, C, 1))).astype(np.float32)
    X[0, 0, :] = 3.25                                   # constant channel: IQR = 0 -> scale 1
    X[0, 1, :] = np.round(X[0, 1, :])                   # many ties
    X[-1, -1, ::7] *= 500.0                             # outliers far beyond the clamp
    if C > 2:
        X[0, 2, : T // 2] = -1.0                        # more than half of the row equal: IQR may be 0 or tiny
    return X


def main():
    ref = import_reference()
    out = {}
    worst = 0.0
    for i, (B, C, T, L, lim, clamp) in enumerate([(3, 5, 360, 60, 20.0, True), (2, 4, 37, 5, 20.0, True),
                                                  (2, 3, 200, 24, 5.0, True), (2, 3, 64, 1, 20.0, False)]):
        X = make_case(100 + i, B, C, T)
        Xt = torch.from_numpy(X.copy())
        want = ref.scaleAndClamp(ref.baseline_correction_single(Xt, L), lim, clamp).numpy()
        got = CR.collate_brain(X, L, lim, clamp)
        err = float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))
        worst = max(worst, err)
        assert want.shape == X.shape and err <= 2e-6, (i, err)
        out[f"X{i}"], out[f"out{i}"] = X, want.astype(np.float32)
        out[f"meta{i}"] = np.array([L, lim, float(clamp)], dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collator.npz"), **out)
    print(f"collator golden written; restatement vs reference worst rel err {worst:.2e}")


if __name__ == "__main__":
    main()
