"""megb200 — B200-native BrainEncoder + CLIPLoss training hot path (hand-written sm_100a kernels behind a C-ABI).

Precision modes (BASELINE.json north_star):
  "bf16"  bf16 activation storage, tcgen05 tensor-core contractions, fp32 accumulation/statistics   (default)
  "fp32"  fp32 storage, fp32 CUDA-core contractions — the check mode (1e-3 vs the fp32 reference)
"""
from __future__ import annotations

import os

import torch

_PRECISION = os.environ.get("MEGB200_PRECISION", "bf16")
_OPS_CACHE = {}
_OPS_OVERRIDE = None


def set_precision(p: str):
    global _PRECISION
    if p not in ("bf16", "fp32", "bf16_simt"):
        raise ValueError("precision must be 'bf16', 'fp32' or 'bf16_simt'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def get_ops(precision: str | None = None):
    """The op backend for a precision mode.  Always the CUDA library: there is no CPU / PyTorch fallback."""
    if _OPS_OVERRIDE is not None:
        return _OPS_OVERRIDE
    p = precision or _PRECISION
    if p not in _OPS_CACHE:
        from .ops import CudaOps
        if p == "bf16":
            _OPS_CACHE[p] = CudaOps(torch.bfloat16, "tc")
        elif p == "bf16_simt":            # bf16 storage on the fp32 CUDA-core engine: on-device cross-check of "bf16"
            _OPS_CACHE[p] = CudaOps(torch.bfloat16, "simt")
        elif p == "fp32":
            _OPS_CACHE[p] = CudaOps(torch.float32, "simt")
        else:
            raise ValueError(p)
    return _OPS_CACHE[p]


def _set_ops_override(ops):
    """Test hook: run the orchestration on another implementation of the op interface (tests/ only)."""
    global _OPS_OVERRIDE
    _OPS_OVERRIDE = ops
