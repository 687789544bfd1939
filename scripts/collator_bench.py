"""N2: time of the device collator (megb200_robust_scale) at the Gwilliams batch shape next to the reference's CPU form
(one sklearn RobustScaler fit per sample, oracle port) on this host."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from megb200.data import robust_scale_  # noqa: E402
from oracle import collator_restatement as CR  # noqa: E402

B, C, T = 256, 208, 360
g = np.random.Generator(np.random.PCG64(0))
X = (g.standard_normal((B, C, T)) * 7).astype(np.float32)
Xd = torch.from_numpy(X).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(6):
    W = Xd.clone()
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    robust_scale_(W, 60, 20.0, True)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
gpu_ms = sorted(ts)[len(ts) // 2]
t0 = time.perf_counter()
ref = CR.collate_brain(X[:32], 60, 20.0, True)
cpu_ms = (time.perf_counter() - t0) * 1e3 * (B / 32)
err = np.abs(W[:32].cpu().numpy() - ref).max() / 20.0
print(f"COLLATOR B={B} C={C} T={T}: device {gpu_ms * 1e3:.0f} us ({2 * X.nbytes / gpu_ms / 1e6:.0f} GB/s of in+out), "
      f"numpy port on this host {cpu_ms:.0f} ms per batch (1 thread, extrapolated from 32 samples); max err/clamp {err:.1e}")
