""""Next" row N4: the optimizer step of the reference's loop (``torch.optim.Adam`` over the encoder's and the loss's
parameters, ``train.py:160-162``) as ONE kernel launch (``megb200_adam_multi``) instead of torch's multi-tensor sequence
(~0.9 ms per step at the Gwilliams shape against a 7.9 ms forward+backward).

Same constructor arguments and update rule as ``torch.optim.Adam`` (L2 ``weight_decay``; ``amsgrad`` / ``maximize`` /
``capturable`` are not supported — the reference uses the defaults).  Complex parameters (the spatial-attention ``z``)
are updated through their real view, exactly as torch does.  Parameters whose ``.grad`` is None are skipped (subjects
absent from the batch, like torch).  CUDA f32 parameters only: anything else raises.
"""
from __future__ import annotations

import torch

from . import lib

__all__ = ["Adam"]


def _real(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("megb200.optim.Adam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._static = {}      # group index -> per-parameter constants (pointers of p / exp_avg / exp_avg_sq / counter)
        self._tables = {}      # (group index, gradient pointers) -> device table

    # The kernel works from a table of raw device pointers into ``self.state``; anything that replaces the state tensors
    # (load_state_dict, add_param_group, unpickling) must drop the cached pointers, and ``step`` re-validates them.
    def _invalidate(self):
        self._static = {}
        self._tables = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._invalidate()

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_static"):
            self._invalidate()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._invalidate()

    @staticmethod
    def _usable(t, p):
        return (torch.is_tensor(t) and t.shape == p.shape and t.dtype == p.dtype and t.device == p.device
                and t.is_contiguous())

    def _prepare(self, gi, group):
        """Per-parameter state: moments like the parameter, one int64 step counter each (device).  Existing state (a
        resumed run: ``load_state_dict`` of this class's or ``torch.optim.Adam``'s checkpoint) is adopted, not zeroed."""
        ps = group["params"]
        for p in ps:
            if not p.is_cuda or _real(p).dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("megb200.optim.Adam: parameters must be contiguous f32 (or complex64) CUDA tensors")
        dev = ps[0].device
        counts = []
        for p in ps:
            st = self.state[p]
            if self._usable(st.get("exp_avg"), p) and self._usable(st.get("exp_avg_sq"), p):
                counts.append(int(float(st.get("step", 0))))       # torch.optim.Adam stores a float tensor
            else:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                counts.append(0)
        steps = torch.tensor(counts, dtype=torch.int64).to(dev)
        info = []
        for i, p in enumerate(ps):
            st = self.state[p]
            st["step"] = steps[i]                       # 0-dim view: torch.optim.Adam's state key, device-resident
            pr = _real(p)
            info.append((pr.data_ptr(), _real(st["exp_avg"]).data_ptr(), _real(st["exp_avg_sq"]).data_ptr(), pr.numel(),
                         steps.data_ptr() + 8 * i))
        self._static[gi] = (dev, steps, info)
        return self._static[gi]

    def _still_valid(self, gi, group):
        """The cached pointers still name the live parameter / moment tensors (someone may have replaced state[p] or
        p.data behind our back; a stale table would make the kernel write freed memory)."""
        ent = self._static.get(gi)
        if ent is None or len(ent[2]) != len(group["params"]):
            return False
        for p, (pp, m, v, _, _) in zip(group["params"], ent[2]):
            st = self.state.get(p)
            if st is None or p.data_ptr() != pp or st["exp_avg"].data_ptr() != m or st["exp_avg_sq"].data_ptr() != v:
                return False
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = group["params"]
            if not ps:
                continue
            if not self._still_valid(gi, group):
                self._tables = {k: v for k, v in self._tables.items() if k[0] != gi}
                self._prepare(gi, group)
            dev, _, info = self._static[gi]
            gptrs = []
            for p in ps:
                g = p.grad
                if g is None:
                    gptrs.append(0)
                    continue
                if g.is_sparse:
                    raise RuntimeError("megb200.optim.Adam does not support sparse gradients")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                gptrs.append(g.data_ptr())              # a complex gradient's real view starts at the same address
            key = (gi, tuple(gptrs))
            ent = self._tables.get(key)
            if ent is None:
                rows = [(pp, gp, m, v, n, sp) for (pp, m, v, n, sp), gp in zip(info, gptrs) if gp != 0]
                if len(self._tables) > 16:
                    self._tables.clear()
                tab = torch.tensor(rows, dtype=torch.int64).pin_memory().to(dev, non_blocking=True) if rows else None
                ent = self._tables[key] = (tab, len(rows))
            tab, n = ent
            if n == 0:
                continue
            with torch.cuda.device(dev):
                lib.call("megb200_adam_multi", tab.data_ptr(), n, float(group["lr"]), float(group["betas"][0]),
                         float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                         torch._C._cuda_getCurrentRawStream(dev.index))
        return loss
