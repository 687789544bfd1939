"""Deterministic synthetic configs, weights and batches for the BrainEncoder + CLIPLoss hot path.

Shapes follow SURVEY.md §8(d): Gwilliams2022-shaped MEG (208 sensors, 27 subjects, 360 samples),
Brennan2018-shaped EEG (60 channels, 33 subjects), GOD-shaped MEG (160 ROI channels, 3 subjects, T=200,
seq2static, 512-d image targets).  Everything is generated with ``numpy.random.Generator(PCG64)`` so the
same arrays are reproduced on any box (the reference's own init RNG is torch's and is not stable across
versions; weights are therefore *loaded* into both the reference and the new modules).

Reference shape provenance: ``meg_decoding/dataclass/gwilliams2022.py:53,250-268,415-422``,
``brennan2018.py:249-258``, ``configs/config.yaml:16-49``.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


class Args(dict):
    """Attribute-style dict: what the reference's constructors read off ``args`` (OmegaConf stand-in)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_BASE = dict(
    D1=270, D2=320, F=1024, K=32, d_drop=0.1, seq2seq=True, dataset="Gwilliams2022", root_dir=".",
    preprocs={"last4layers": True}, reduction="mean", init_temperature=5.1, model="brain_encoder",
)

CONFIGS: Dict[str, dict] = {
    # small shapes for golden fixtures / fast parity tests (all dims deliberately awkward)
    "tiny": dict(_BASE, B=6, C=20, T=40, num_subjects=4, D1=30, D2=32, F=64, K=8,
                 preprocs={"last4layers": False}),
    "tiny_static": dict(_BASE, B=5, C=12, T=24, num_subjects=3, D1=30, D2=32, F=48, K=8, seq2seq=False,
                        preprocs={"last4layers": False}),
    # real channel widths, few samples: exercises D1=270 / D2=320 / F=1024 padding and tiling
    "narrow": dict(_BASE, B=4, C=60, T=360, num_subjects=5, dataset="Brennan2018"),
    "narrow_static": dict(_BASE, B=6, C=160, T=200, num_subjects=3, F=512, seq2seq=False, dataset="GOD",
                          preprocs={"last4layers": False}),
    # encoder variants ("next" row N4): BrainEncoderSeq2Static (models.py:465-512: per-block kernel sizes, AvgPool1d(3,2)
    # between blocks, global average after block 4) and the dilated paper block (tests/test_modules/models.py:267-312)
    "tiny_pooled": dict(_BASE, B=5, C=12, T=100, num_subjects=3, D1=30, D2=32, F=48, K=8, seq2seq=False,
                        preprocs={"last4layers": False}, model="brain_endcoder_seq2static",
                        ConvBlocks={"ks": [11, 5, 5, 3, 3]}),
    "narrow_pooled": dict(_BASE, B=6, C=160, T=200, num_subjects=3, F=512, seq2seq=False, dataset="GOD",
                          preprocs={"last4layers": False}, model="brain_endcoder_seq2static",
                          ConvBlocks={"ks": [11, 11, 5, 3, 3]}),
    "tiny_dilated": dict(_BASE, B=6, C=20, T=40, num_subjects=4, D1=30, D2=32, F=64, K=8,
                         preprocs={"last4layers": False}, conv_dilations="paper"),
    "narrow_dilated": dict(_BASE, B=4, C=60, T=360, num_subjects=5, dataset="Brennan2018", conv_dilations="paper"),
    # BASELINE.json configs
    "brennan": dict(_BASE, B=64, C=60, T=360, num_subjects=33, dataset="Brennan2018"),          # configs[0]
    "gwilliams": dict(_BASE, B=256, C=208, T=360, num_subjects=27),                             # configs[1]
    "god": dict(_BASE, B=2048, C=160, T=200, num_subjects=3, F=512, seq2seq=False, dataset="GOD",
                preprocs={"last4layers": False}),                                               # configs[3]
    "brennan49": dict(_BASE, B=256, C=60, T=360, num_subjects=49, dataset="Brennan2018"),       # configs[4]
}


def get_config(name: str, **overrides) -> Args:
    cfg = Args(CONFIGS[name])
    cfg.update(overrides)
    cfg["name"] = name
    return cfg


def out_features(cfg) -> int:
    """models.py:348 — F is forced to 1024 when wav2vec2 last-4-layer targets are used."""
    return 1024 if cfg.preprocs["last4layers"] else cfg.F


def conv_kernel_sizes(cfg):
    """Per-block Conv1d kernel sizes: 3 (models.py:277), or ``args.ConvBlocks.ks`` for BrainEncoderSeq2Static (models.py:481)."""
    cb = cfg.get("ConvBlocks") if hasattr(cfg, "get") else getattr(cfg, "ConvBlocks", None)
    if cb is None or cfg.get("model") != "brain_endcoder_seq2static":
        return [3] * 5
    return [int(k) for k in (cb["ks"] if isinstance(cb, dict) else cb.ks)]


def paper_dilations():
    """Dilations of the paper's block as the reference's alternative formulation writes them
    (tests/test_modules/models.py:275-297, k = 1..5; ``2 * k + 1 % 5`` parses as ``2 * k + 1``)."""
    return [(2 ** (2 * k % 5), 2 ** (2 * k + 1 % 5), 2) for k in range(1, 6)]


def conv_dilations(cfg):
    d = cfg.get("conv_dilations") if hasattr(cfg, "get") else getattr(cfg, "conv_dilations", None)
    if d is None:
        return [(1, 1, 1)] * 5                                   # the reference code as it runs (models.py:289,297,305)
    if isinstance(d, str):
        if d != "paper":
            raise ValueError(f"unknown conv_dilations {d!r}")
        return paper_dilations()
    return [tuple(int(x) for x in row) for row in d]


def sensor_locations_raw(C: int, seed: int = 0) -> np.ndarray:
    """Raw synthetic 2-D sensor coordinates (what mne's layout would return), uniform in the unit square."""
    return np.random.Generator(np.random.PCG64(seed)).random((C, 2))


def sensor_locations(C: int, seed: int = 0) -> np.ndarray:
    """Synthetic 2-D layout: uniform in the unit square, min-max normalised then *0.8+0.1 (layout.py:42-45)."""
    loc = sensor_locations_raw(C, seed)
    loc = (loc - loc.min(axis=0)) / (loc.max(axis=0) - loc.min(axis=0))
    return (loc * 0.8 + 0.1).astype(np.float32)


def _uniform(rng, shape, bound):
    return ((rng.random(shape, dtype=np.float64) * 2.0 - 1.0) * bound).astype(np.float32)


def make_state(cfg, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random-init weights keyed exactly like the reference's ``BrainEncoder.state_dict()``.

    Conv weights/biases ~ U(+-1/sqrt(fan_in)) like torch's default init; BN affine is made non-trivial;
    ``z`` ~ U[0,1) for both real and imaginary parts (models.py:186).
    """
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    D1, D2, Fo, K, S = cfg.D1, cfg.D2, out_features(cfg), cfg.K, cfg.num_subjects
    st: Dict[str, np.ndarray] = {}
    z = rng.random((D1, K * K)) + 1j * rng.random((D1, K * K))
    st["subject_block.spatial_attention.z"] = z.astype(np.complex64)
    b = 1.0 / math.sqrt(D1)
    st["subject_block.conv.weight"] = _uniform(rng, (D1, D1, 1), b)
    st["subject_block.conv.bias"] = _uniform(rng, (D1,), b)
    for s in range(S):
        st[f"subject_block.subject_layer.{s}.weight"] = _uniform(rng, (D1, D1, 1), b)
    ks = conv_kernel_sizes(cfg)
    for k in range(5):
        cin = D1 if k == 0 else D2
        pre = f"conv_blocks.conv{k}."
        for name, ci, co in (("conv0", cin, D2), ("conv1", D2, D2), ("conv2", D2, 2 * D2)):
            bb = 1.0 / math.sqrt(ci * ks[k])
            st[pre + name + ".weight"] = _uniform(rng, (co, ci, ks[k]), bb)
            st[pre + name + ".bias"] = _uniform(rng, (co,), bb)
        for bn in ("batchnorm0", "batchnorm1"):
            st[pre + bn + ".weight"] = (0.5 + rng.random(D2)).astype(np.float32)
            st[pre + bn + ".bias"] = _uniform(rng, (D2,), 0.2)
            st[pre + bn + ".running_mean"] = np.zeros(D2, np.float32)
            st[pre + bn + ".running_var"] = np.ones(D2, np.float32)
            st[pre + bn + ".num_batches_tracked"] = np.zeros((), np.int64)
    bb = 1.0 / math.sqrt(D2)
    st["conv_final1.weight"] = _uniform(rng, (2 * D2, D2, 1), bb)
    st["conv_final1.bias"] = _uniform(rng, (2 * D2,), bb)
    bb = 1.0 / math.sqrt(2 * D2)
    st["conv_final2.weight"] = _uniform(rng, (Fo, 2 * D2, 1), bb)
    st["conv_final2.bias"] = _uniform(rng, (Fo,), bb)
    return st


def make_batch(cfg, seed: int = 0, B: int | None = None, index_dtype=np.int32):
    """(X f32 [B,C,T] clamped to +-20, Y f32 [B,F,T] | [B,F], subject_idxs int32 [B]) — gwilliams2022.py:654-662."""
    B = cfg.B if B is None else B
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    X = np.clip(rng.standard_normal((B, cfg.C, cfg.T), dtype=np.float32), -20.0, 20.0)
    Fo = out_features(cfg)
    yshape = (B, Fo, cfg.T) if cfg.seq2seq else (B, Fo)
    Y = rng.standard_normal(yshape, dtype=np.float32)
    subj = rng.integers(0, cfg.num_subjects, size=(B,)).astype(index_dtype)
    return X, Y, subj


def drop_center(cfg, seed: int = 1) -> int:
    """Spatial-dropout centre sensor; the reference draws it from the global NumPy RNG (models.py:236)."""
    return int(np.random.Generator(np.random.PCG64(3000 + seed)).integers(0, cfg.C))


def encoder_flops_per_sample(cfg) -> float:
    """Algorithmic fwd+bwd FLOPs per sample, 3x-forward convention."""
    D1, D2, Fo, C, T = cfg.D1, cfg.D2, out_features(cfg), cfg.C, cfg.T
    macs = T * (D1 * C + 2 * D1 * D1 + 3 * D2 * D1 + 3 * D2 * D2 + 6 * D2 * D2
                + 4 * 12 * D2 * D2 + 2 * D2 * D2 + 2 * D2 * Fo)
    return 3.0 * 2.0 * macs


def clip_flops(cfg, b_loc: int, b_glob: int) -> float:
    """Logits GEMM + dZ GEMM per rank-step."""
    D = out_features(cfg) * (cfg.T if cfg.seq2seq else 1)
    return 2.0 * (2.0 * b_loc * b_glob * D)
