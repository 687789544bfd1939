"""The C-ABI library loads on a CPU-only box and exports every symbol include/megb200.h declares, with the
argument lists the ctypes binding assumes (no compute calls here)."""
import ctypes
import os
import re

from megb200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CT = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t}


def _parse_header():
    src = open(os.path.join(ROOT, "include", "megb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(megb200_\w+)\s*\(([^)]*)\)\s*;", src):
        args = []
        body = m.group(3).strip()
        if body != "void":
            for a in body.split(","):
                a = a.strip()
                if "*" in a:
                    args.append(ctypes.c_void_p)
                else:
                    ty = re.sub(r"\s+\w+$", "", a).replace("const ", "").strip()
                    args.append(_CT[ty])
        decls[m.group(2)] = args
    return decls


def test_header_symbols_exported(built_lib):
    decls = _parse_header()
    assert len(decls) >= 25
    for name in decls:
        assert hasattr(built_lib, name), f"{name} declared in include/megb200.h but not exported"


def test_ctypes_signatures_match_header(built_lib):
    decls = _parse_header()
    for name, argtypes in L.SIGNATURES.items():
        assert name in decls, name
        assert [a for a in argtypes] == decls[name], f"{name}: binding {argtypes} != header {decls[name]}"
    missing = set(decls) - set(L.SIGNATURES) - {"megb200_version", "megb200_last_error", "megb200_launch_count", "megb200_get_option"}
    assert not missing, f"header functions without a binding: {missing}"


def test_version_and_error_string(built_lib):
    assert built_lib.megb200_version() == 100
    assert isinstance(built_lib.megb200_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(built_lib):
    # null pointers are rejected before any CUDA call
    rc = built_lib.megb200_gelu(None, None, 8, L.BF16, None)
    assert rc == -1
    assert b"gelu" in built_lib.megb200_last_error()


def test_set_option_roundtrip_and_unknown_name(built_lib):
    assert L.get_option("tapgemm_pair_min_items") == -1
    with L.options(tapgemm_pair_min_items=1, clip_dz_large=1):
        assert L.get_option("tapgemm_pair_min_items") == 1 and L.get_option("clip_dz_large") == 1
    assert L.get_option("tapgemm_pair_min_items") == -1 and L.get_option("clip_dz_large") == 0
    assert built_lib.megb200_set_option(b"no_such_option", 1) == -1
    assert b"no_such_option" in built_lib.megb200_last_error()
