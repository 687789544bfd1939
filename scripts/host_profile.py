"""Where does the HOST time of one training step go?  cProfile over a few steps of the bench workload (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "meg-decoding_b200")]
import torch  # noqa: E402

import megb200  # noqa: E402
from megb200 import synth  # noqa: E402
from tests import helpers as H  # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.get_config(sys.argv[1] if len(sys.argv) > 1 else "gwilliams")
enc, lossf = H.make_modules(cfg, dev)
enc.train()
enc.drop_center = synth.drop_center(cfg)
params = list(enc.parameters()) + list(lossf.parameters())
X, Y, subj = synth.make_batch(cfg, seed=0)
X, Y, subj = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev), torch.from_numpy(subj)


def step():
    for p in params:
        p.grad = None
    lossf.prefetch_targets(Y)
    loss = lossf(Y, enc(X, subj))
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print(f"host enqueue {host * 1e3:.2f} ms/step, with sync {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(22)
