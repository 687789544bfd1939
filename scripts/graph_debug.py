"""Find the call that invalidates stream capture: wraps every megb200 op, checks the capture status after each."""
import os
import sys

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
X, Y, subj = synth.make_batch(cfg, seed=0)
X, Y = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
subj_d = torch.from_numpy(subj).to(dev).to(torch.int32)
center = torch.zeros(1, dtype=torch.int64, device=dev)
enc.drop_center = center
params = list(enc.parameters()) + list(lossf.parameters())
ops = megb200.get_ops()
cudart = torch.cuda.cudart()
state = {"bad": None, "n": 0}


def status():
    try:
        return torch.cuda.is_current_stream_capturing()
    except Exception as e:
        return f"ERR {e}"[:80]


def wrap(name, fn):
    def w(*a, **k):
        r = fn(*a, **k)
        state["n"] += 1
        st = status()
        if st is not True and state["bad"] is None and state["armed"]:
            state["bad"] = (state["n"], name, st)
            print("CAPTURE BROKEN after op", state["n"], name, st, flush=True)
        return r
    return w


def step():
    for p in params:
        p.grad = None
    loss = lossf(Y, enc(X, subj_d))
    loss.backward()
    return loss.detach()


state["armed"] = False
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
for name in dir(ops):
    if not name.startswith("_") and callable(getattr(ops, name)):
        setattr(ops, name, wrap(name, getattr(ops, name)))
for p in params:
    p.grad = None
g = torch.cuda.CUDAGraph()
state["armed"] = True
try:
    with torch.cuda.graph(g):
        loss = step()
    print("capture OK, ops:", state["n"])
    g.replay()
    torch.cuda.synchronize()
    print("replay OK loss", float(loss))
except Exception as e:
    print("capture FAILED:", str(e)[:300])
    print("first bad:", state["bad"], "ops seen", state["n"])
