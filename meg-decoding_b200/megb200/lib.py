"""ctypes binding of libmegb200.so (the C-ABI declared in include/megb200.h).

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``); there is no fallback: if the
shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmegb200.so")
# developer builds only (csrc/build.sh with MEGB200_EXTRA_NVCC_FLAGS writes /tmp/libmegb200_dev.so): never set in production
if os.environ.get("MEGB200_DEV_LIB"):
    LIB_PATH = os.environ["MEGB200_DEV_LIB"]

F32, BF16 = 0, 1
EPI_STORE, EPI_STORE_F32, EPI_GELU, EPI_GLU = 0, 1, 2, 3
SIMT, TC = 0, 1

_p, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> argtypes; mirrors include/megb200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "megb200_pack_x": [_p, _p, _i, _i, _i, _i, _i, _p],
    "megb200_pack_weight": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "megb200_pack_weight_multi": [_p, _i, _i, _p],
    "megb200_tapgemm": [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _i, _p, _i, _p, _i, _p, _i, _i, _p],
    "megb200_wgrad": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _sz, _i, _i, _p],
    "megb200_group_order": [_p, _i, _i, _p, _p],
    "megb200_colsum": [_p, _i, _i, _i, _i, _p, _i, _p, _i, _p],
    "megb200_sgemm": [_i, _i, _i, _i, _i, _f, _p, _i, _ll, _p, _i, _ll, _f, _p, _i, _ll, _i, _p, _sz, _p],
    "megb200_sa_softmax": [_p, _p, _p, _p, _i, _i, _p],
    "megb200_sa_softmax_bwd": [_p, _p, _p, _p, _i, _i, _p],
    "megb200_bn_stats": [_p, _ll, _i, _i, _p, _i, _p],
    "megb200_bn_finalize": [_p, _p, _p, _p, _p, _ll, _f, _f, _i, _p, _p, _p, _p, _i, _p],
    "megb200_bn_gelu": [_p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_bn_gelu_train": [_p, _p, _p, _p, _p, _p, _ll, _f, _f, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_bn_gelu_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_bn_gelu_bwd_reduce": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_bn_bwd_apply": [_p, _p, _p, _p, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_glu": [_p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_glu_bwd": [_p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "megb200_gelu": [_p, _p, _ll, _i, _p],
    "megb200_gelu_bwd": [_p, _p, _p, _p, _ll, _i, _i, _i, _p],
    "megb200_time_mean": [_p, _p, _i, _i, _i, _i, _i, _p],
    "megb200_time_mean_bwd": [_p, _p, _i, _i, _i, _i, _i, _p],
    "megb200_avgpool1d": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "megb200_avgpool1d_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "megb200_clip_prep_y": [_p, _i, _p, _p, _i, _i, _i, _i, _p],
    "megb200_rowsumsq": [_p, _p, _i, _ll, _i, _p],
    "megb200_clip_logits": [_p, _p, _p, _i, _i, _ll, _i, _i, _p],
    "megb200_clip_ce_partial": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "megb200_clip_ce_finish": [_p, _p, _p, _p, _p, _p, _i, _f, _p, _p, _i, _p, _i, _i, _i, _p],
    "megb200_clip_dz": [_p, _i, _p, _p, _p, _p, _p, _i, _i, _ll, _i, _i, _p],
    "megb200_retrieval_rank": [_p, _p, _p, _i, _i, _i, _p],
    "megb200_gather_segments": [_p, _p, _p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p],
    "megb200_robust_scale": [_p, _p, _ll, _i, _i, _f, _i, _p],
    "megb200_adam_multi": [_p, _i, _f, _f, _f, _f, _f, _p],
    "megb200_cast": [_p, _p, _ll, _i, _i, _p],
    "megb200_set_option": [_p, _ll],
    "megb200_clip_lse_merge": [_p, _i, _i, _p, _p],
    "megb200_clip_fused_stats": [_p, _p, _p, _p, _p, _i, _i, _ll, _i, _p, _p, _p, _i, _i, _p],
    "megb200_clip_fused_loss": [_p, _p, _p, _i, _i, _f, _p, _p],
    "megb200_clip_fused_coef": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _ll, _i, _f, _i, _i, _p, _i, _p, _p, _i, _i, _p],
    "megb200_ipc_export": [_p, _p, _p],
    "megb200_ipc_import": [_p, _p],
    "megb200_ipc_release": [_p],
    "megb200_copy_async": [_p, _p, _sz, _p],
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or meg-decoding_b200/csrc/build.sh).  There is no CPU or PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    lib.megb200_version.restype = C.c_int
    lib.megb200_version.argtypes = []
    lib.megb200_last_error.restype = C.c_char_p
    lib.megb200_last_error.argtypes = []
    lib.megb200_launch_count.restype = C.c_longlong
    lib.megb200_launch_count.argtypes = []
    lib.megb200_get_option.restype = C.c_longlong
    lib.megb200_get_option.argtypes = [C.c_char_p]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().megb200_last_error().decode("utf-8", "replace")
        kind = "argument error" if rc < 0 else f"cudaError {rc}"
        raise RuntimeError(f"{what}: {kind}: {msg}")


def launch_count() -> int:
    return int(load().megb200_launch_count())


def set_option(name: str, value: int) -> int:
    """Force a kernel variant (see include/megb200.h); returns the previous value."""
    lib = load()
    old = int(lib.megb200_get_option(name.encode()))
    check(lib.megb200_set_option(name.encode(), int(value)), "megb200_set_option")
    return old


def get_option(name: str) -> int:
    return int(load().megb200_get_option(name.encode()))


class options:
    """Context manager: ``with lib.options(tapgemm_pair_min_items=1): ...`` restores the previous values on exit."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)
