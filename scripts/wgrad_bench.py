"""Micro-benchmark of the weight-gradient kernels at config 2's layer shapes (B=256, T=360): the CTA-pair kernel
(tc_wgrad2_kernel) against the single-CTA kernel, per launch with CUDA events, L2 flushed between launches.
    python scripts/wgrad_bench.py [--ncu]     (--ncu: one launch per kernel and shape, for `ncu --set full`)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import torch  # noqa: E402

import megb200  # noqa: E402
from megb200 import lib  # noqa: E402

dev = torch.device("cuda:0")
ops = megb200.get_ops("bf16")
B, T = 256, 360
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ncu = "--ncu" in sys.argv
SHAPES = [("conv 320->320 k3", 320, 320, 3), ("conv0 270->320 k3", 320, 270, 3), ("conv2 320->640 k3", 640, 320, 3),
          ("head 320->640 k1", 640, 320, 1), ("head 640->1024 k1", 1024, 640, 1)]
if ncu:
    SHAPES = SHAPES[:1] + SHAPES[2:3]


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for name, M, C, taps in SHAPES:
    g = torch.Generator(device=dev).manual_seed(1)
    Cs, Ms = (C + 7) // 8 * 8, (M + 7) // 8 * 8
    dU = torch.randn(B, T, Ms, device=dev, generator=g).to(torch.bfloat16)
    X = torch.randn(B, T, Cs, device=dev, generator=g).to(torch.bfloat16)
    dW = torch.zeros(1, M, C, taps, device=dev)
    fl = 2.0 * B * T * M * C * taps
    res = {}
    for pair in (1, 0):
        with lib.options(wgrad_pair=pair):
            if ncu:
                dW.zero_()
                ops.wgrad(dU, M, X, C, taps, 1, None, None, dW)
                torch.cuda.synchronize()
                continue
            t = timeit(lambda: ops.wgrad(dU, M, X, C, taps, 1, None, None, dW))
            res[pair] = (t, lib.get_option("last_wgrad_splits"))
    if not ncu:
        print(f"{name:22s} pair {res[1][0] * 1e3:6.1f} us ({fl / res[1][0] / 1e9:5.0f} TF/s, {res[1][1]} splits)   "
              f"single {res[0][0] * 1e3:6.1f} us ({fl / res[0][0] / 1e9:5.0f} TF/s, {res[0][1]} splits)", flush=True)
