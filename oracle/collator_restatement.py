"""ORACLE — TEST INFRASTRUCTURE ONLY ("next" row N2: the batch pre-processing collator).

CPU restatement (numpy) of what ``Gwilliams2022Collator.forward`` does to the brain data
(``meg_decoding/dataclass/gwilliams2022.py:653-662``):

  1. ``baseline_correction_single`` (``meg_decoding/utils/preproc_utils.py:128-142``): per (sample, channel) subtract
     the mean of the first ``baseline_len_samp`` time steps;
  2. ``scaleAndClamp`` (``preproc_utils.py:69-90``): per sample ``sklearn.preprocessing.RobustScaler().fit_transform``
     over time — per channel ``(x - median) / (q75 - q25)`` with numpy's linear-interpolated quantiles and scikit-learn's
     zero-scale rule (a scale below 10 eps becomes 1, ``sklearn/preprocessing/_data.py::_handle_zeros_in_scale``) — then
     ``clamp_(-clamp_lim, clamp_lim)``.

scikit-learn (1.9.0 in this image; the reference pins none) is the third-party dependency that holds step 2; the
restatement below is pinned against the reference's own functions by ``oracle/make_golden_collator.py``.
Only tests/ and bench legs may import this module.
"""
from __future__ import annotations

import numpy as np


def baseline_correct(X: np.ndarray, baseline_len: int) -> np.ndarray:
    """X: f32 [B, C, T] -> X - mean(X[..., :baseline_len]) per (sample, channel)   (preproc_utils.py:137-140)."""
    X = np.asarray(X, dtype=np.float32)
    base = X[:, :, :baseline_len].mean(axis=2, dtype=np.float32)
    return (X - base[:, :, None]).astype(np.float32)


def robust_scale_clamp(X: np.ndarray, clamp_lim: float, clamp: bool) -> np.ndarray:
    """Per (sample, channel) row over time: (x - median) / IQR, IQR < 10 eps -> 1, optional clamp (preproc_utils.py:80-88)."""
    X = np.asarray(X, dtype=np.float32)
    med = np.median(X, axis=2).astype(np.float32)                 # center_: f32 (median of an f32 column)
    q25, q75 = np.percentile(X, (25.0, 75.0), axis=2)             # f64: numpy interpolates f32 input in double
    scale = q75 - q25                                             # scale_: f64
    scale = np.where(scale < 10 * np.finfo(np.float64).eps, 1.0, scale)
    out = ((X - med[:, :, None]).astype(np.float32) / scale[:, :, None]).astype(np.float32)
    if clamp:
        out = np.clip(out, -np.float32(clamp_lim), np.float32(clamp_lim))
    return out


def collate_brain(X: np.ndarray, baseline_len: int, clamp_lim: float = 20.0, clamp: bool = True) -> np.ndarray:
    """gwilliams2022.py:658-659."""
    return robust_scale_clamp(baseline_correct(X, baseline_len), clamp_lim, clamp)
