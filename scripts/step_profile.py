"""In-step time of every megb200 op (CUDA events around each C-ABI call inside real training steps; warm L2, real
neighbours — unlike ncu's serialised cold-cache per-launch times).  Usage: python scripts/step_profile.py [config] [B]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import torch  # noqa: E402

import megb200  # noqa: E402
from megb200 import synth  # noqa: E402
from tests import helpers as H  # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.get_config(sys.argv[1] if len(sys.argv) > 1 else "gwilliams")
if len(sys.argv) > 2:
    cfg["B"] = int(sys.argv[2])
enc, lossf = H.make_modules(cfg, dev)
enc.train()
enc.drop_center = synth.drop_center(cfg)
params = list(enc.parameters()) + list(lossf.parameters())
X, Y, subj = synth.make_batch(cfg, seed=0)
X, Y, subj = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev), torch.from_numpy(subj)
ops = megb200.get_ops()
records = []
SKIP = {"workspace", "empty"}


def nbytes(args):
    n = 0
    for a in args:
        if torch.is_tensor(a):
            n += a.numel() * a.element_size()
    return n


def wrap(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        shp = "x".join(str(s) for s in a[0].shape) if a and torch.is_tensor(a[0]) else ""
        records.append((name, shp, nbytes(a) + nbytes(k.values()), e0, e1))
        return r
    return w


def step():
    for p in params:
        p.grad = None
    loss = lossf(Y, enc(X, subj))
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
for name in dir(ops):
    if name.startswith("_") or name in SKIP:
        continue
    fn = getattr(ops, name)
    if callable(fn):
        setattr(ops, name, wrap(name, fn))
N = 5
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(N):
    step()
b.record()
torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0, 0])
for name, shp, nb, e0, e1 in records:
    t = tot[(name, shp)]
    t[0] += e0.elapsed_time(e1)
    t[1] += 1
    t[2] += nb
step_ms = a.elapsed_time(b) / N
covered = sum(t[0] for t in tot.values()) / N
print(f"step {step_ms:.3f} ms (with per-op events), ops cover {covered:.3f} ms")
for (name, shp), (ms, n, nb) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {ms / N * 1e3:8.1f} us/step  n={n // N:3d}  avg {ms / n * 1e3:7.1f} us  {nb / n / (ms / n * 1e-3) / 1e12:5.2f} TB/s(args)  {name} {shp}")
