"""Micro-benchmark of the CLIP contraction kernels at data-parallel shapes on ONE GPU: By = world * Bz global targets
against Bz local embeddings (what a rank of an N-GPU job runs).  Prints time and TFLOP/s per kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import torch  # noqa: E402

import megb200  # noqa: E402

dev = torch.device("cuda:0")
ops = megb200.get_ops("bf16")
Bz, D = 256, 1024 * 360
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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


for world in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    By = world * Bz
    g = torch.Generator(device=dev).manual_seed(1)
    Yp = torch.randn(By, D, device=dev, generator=g).to(torch.bfloat16)
    Zp = torch.randn(Bz, D, device=dev, generator=g).to(torch.bfloat16)
    raw = torch.zeros(By, Bz, device=dev)
    ldc = (By + 7) // 8 * 8
    coef = (torch.randn(Bz, ldc, device=dev, generator=g) * 0.01).to(torch.bfloat16)
    cvec = torch.randn(Bz, device=dev, generator=g)
    gs = torch.ones(1, device=dev)
    dZ = torch.empty_like(Zp)
    fl = 2.0 * By * Bz * D

    def f_logits():
        raw.zero_()
        ops.clip_logits(Yp, Zp, raw)

    t1 = timeit(f_logits)
    t2 = timeit(lambda: ops.clip_dz(coef, Yp, Zp, cvec, gs, dZ))
    # correctness spot check against torch on a slice of d
    sl = slice(0, 4096)
    ref = coef[:, :By].float() @ Yp[:, sl].float() - Zp[:, sl].float() * cvec[:, None]
    err = ((dZ[:, sl].float() - ref).norm() / ref.norm()).item()
    refl = Yp[:16].float() @ Zp.float().T
    f_logits()
    errl = ((raw[:16] - refl).norm() / refl.norm()).item()
    print(f"CLIP world={world} By={By}: logits {t1 * 1e3:.0f} us ({fl / t1 / 1e9:.0f} TF/s, err {errl:.1e})  "
          f"dz {t2 * 1e3:.0f} us ({fl / t2 / 1e9:.0f} TF/s, {(By + 2 * Bz) * D * 2 / t2 / 1e9:.2f} TB/s min-traffic, err {err:.1e})",
          flush=True)
    del Yp, Zp, raw, coef, dZ
