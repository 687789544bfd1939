"""BASELINE.json configs[3] ("GOD-style ... large batch B x B logits sweep to 16k") on ONE GPU: CLIPLoss forward +
backward alone, through the public module, D = 512 static embeddings, B = 1k .. 16k, the never-materialised form
(megb200_clip_fused_*) next to the materialised pipeline.  Time per call (CUDA events, L2 flushed between calls), peak
extra device memory of the call, loss agreement of the two forms and against an f64 torch restatement of loss.py:86-107
on the same stored (bf16) operands.  One JSON line per batch size on stdout.

    python scripts/clip_sweep.py [B ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import torch  # noqa: E402

import megb200  # noqa: E402
from megb200 import lib, synth  # noqa: E402
from meg_decoding.utils.loss import CLIPLoss  # noqa: E402

dev = torch.device("cuda:0")
megb200.set_precision("bf16")
D = 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lossf = CLIPLoss(synth.Args(reduction="mean", init_temperature=5.1)).to(dev)


def run(B, fused):
    g = torch.Generator(device=dev).manual_seed(B)
    Y = torch.randn(B, D, device=dev, generator=g)
    Z = (0.2 * Y + torch.randn(B, D, device=dev, generator=g)).to(torch.bfloat16).requires_grad_(True)
    out = {}
    with lib.options(clip_fused=1 if fused else 0):
        def step():
            Z.grad = None
            lossf.temp.grad = None
            loss = lossf(Y, Z)
            loss.backward()
            return loss
        step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss = step()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out["ms"] = sorted(ts)[len(ts) // 2]
        out["peak_extra_mb"] = (torch.cuda.max_memory_allocated(dev) - base) / 1e6
        out["loss"] = float(loss)
        out["dz_norm"] = float(Z.grad.float().norm())
        out["dtemp"] = float(lossf.temp.grad)
    return out, Y, Z


for B in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]:
    fu, Y, Z = run(B, True)
    ma, _, _ = run(B, False)
    # f64 restatement of loss.py:86-107 on the stored operands (rows = Y rounded to bf16 as the kernels contract them)
    y = Y.to(torch.bfloat16).double()
    z = Z.detach().double()
    yn = y / Y.double().norm(dim=1, keepdim=True)            # the norm is taken from the unrounded targets
    zn = z / z.norm(dim=1, keepdim=True)
    ref = None
    if B <= 8192:
        L = (yn @ zn.T) * float(torch.exp(lossf.temp.detach().double()))
        t = torch.arange(B, device=dev)
        ref = float((torch.nn.functional.cross_entropy(L, t) + torch.nn.functional.cross_entropy(L.T, t)) / 2)
        del L
    flops = 2.0 * B * B * D
    line = {"B": B, "D": D, "fused_ms": fu["ms"], "materialised_ms": ma["ms"], "fused_peak_extra_mb": fu["peak_extra_mb"],
            "materialised_peak_extra_mb": ma["peak_extra_mb"], "logits_matrix_mb_f32": B * B * 4 / 1e6,
            "loss_fused": fu["loss"], "loss_materialised": ma["loss"], "loss_f64_restatement": ref,
            "loss_rel_fused_vs_materialised": abs(fu["loss"] - ma["loss"]) / abs(ma["loss"]),
            "dz_norm_rel": abs(fu["dz_norm"] - ma["dz_norm"]) / ma["dz_norm"],
            "dtemp_rel": abs(fu["dtemp"] - ma["dtemp"]) / abs(ma["dtemp"]),
            "fused_tflops_3gemm": 3 * flops / (fu["ms"] * 1e-3) / 1e12}
    print(json.dumps(line), flush=True)
    del Y, Z
