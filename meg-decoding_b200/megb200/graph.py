"""One training step replayed from a CUDA graph.

The hot path launches ~145 kernels per step through the C-ABI; enqueueing them from Python costs 5-7 ms of host time
against ~8.6 ms of device time at the Gwilliams shape, so a busy host (or a faster GPU step) makes the loop launch
bound.  ``GraphedTrainStep`` records ``zero grads -> BrainEncoder.forward -> CLIPLoss -> backward`` once -- the same
kernels, launched through the same C-ABI calls while the stream is capturing -- and replays the recording: one
``cudaGraphLaunch`` per step.

What makes the step capturable (nothing in it may depend on host values that change between steps):
  * inputs live in static device buffers that ``__call__`` refills (``copy_`` from pinned or device memory);
  * the subject indices are a device tensor (the drop-in module routes on the device; no subject's gradient is None);
  * the spatial-dropout centre (models.py:236 draws one per forward call) is a device scalar the mask is gathered by,
    refilled before every replay from the same NumPy RNG the reference uses;
  * tensor maps, packed-weight tables and shared-memory attributes are created by the warm-up steps, so the captured
    region contains kernel launches and memsets only.

Data parallel (``megb200.parallel.init()`` done): the recording includes the hot path's exchange step -- the gather of
the packed targets on a forked side stream (copy-engine pulls from peer-mapped slabs between two 4-byte fences, or an
NCCL all-gather; either overlaps the encoder forward inside the graph), ONE all-gather of the softmax partials and
the in-place SUM all-reduce of the flat gradient pool -- captured in ``thread_local`` error mode because the NCCL
watchdog thread polls events while the stream captures.  Every rank must build the step at the same point of the
program (the warm-up steps run real collectives).
"""
from __future__ import annotations

from typing import Optional

import os

import numpy as np
import torch

from . import lib, parallel

__all__ = ["GraphedTrainStep", "ModuleGraphs", "accelerate"]


class _GradSync:
    """Starts the all-reduce of the gradients that are final early (head + conv blocks 4..1, ~70 % of the bytes) on a
    side stream while the backward pass of block 0 and the subject block still runs (EncoderEngine.backward calls
    ``early``); ``join`` makes the main stream wait for it.  Opt-in (MEGB200_DP_EARLY_BUCKET=1): measured on 2 x B200 the split is
    slightly SLOWER than one all-reduce at the end (7.71 vs 7.59 ms per step) -- the NCCL kernel takes SMs from the
    persistent block-0 kernels it overlaps, which costs more than the ~0.1 ms of all-reduce it hides."""

    def __init__(self, dev):
        self.dev = dev
        self.side = torch.cuda.Stream(device=dev)
        self.enabled = os.environ.get("MEGB200_DP_EARLY_BUCKET", "0") == "1"
        self.pending = False

    def early(self, flat):
        self.pending = False
        if not self.enabled:
            return
        cur = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            parallel.allreduce_flat(flat)
        self.pending = True

    def join(self):
        if self.pending:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)


class GraphedTrainStep:
    def __init__(self, encoder, loss_fn, X, Y, subject_idxs, drop_center: Optional[int] = None, warmup: int = 3):
        self.dist = parallel.context()
        self.enc, self.lossf = encoder, loss_fn
        dev = next(encoder.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the modules on a CUDA device")
        self.dev = dev
        self.C = int(X.shape[1])
        self.S = int(encoder.num_subjects)
        self.X = torch.empty(tuple(X.shape), dtype=torch.float32, device=dev)
        # targets: dense f32 / bf16 (B, F, T) | (B, F), or megb200.data.PackedTargets (no per-step pack pass at all)
        if hasattr(Y, "rows"):
            from .data import PackedTargets
            self.Y = PackedTargets(torch.empty_like(Y.rows, device=dev), torch.empty_like(Y.sumsq, device=dev),
                                   Y.feature_shape)
        else:
            self.Y = torch.empty(tuple(Y.shape), dtype=torch.bfloat16 if Y.dtype == torch.bfloat16 else torch.float32,
                                 device=dev)
        self.subj = torch.zeros(int(X.shape[0]), dtype=torch.int32, device=dev)
        self.center = torch.zeros(1, dtype=torch.int64, device=dev)
        self._one = torch.ones(1, dtype=torch.float32, device=dev)
        self.params = [p for p in list(encoder.parameters()) + list(loss_fn.parameters()) if p.requires_grad]
        self._user_center = encoder.drop_center
        self._side = torch.cuda.Stream(device=dev) if self.dist is not None else None
        self._sync = _GradSync(dev) if self.dist is not None else None
        self.load(X, Y, subject_idxs, drop_center)

        # warm-up on a side stream (torch's capture recipe); module buffers (BatchNorm running statistics, counters) are
        # restored afterwards so that building the graph has no side effect on the model
        saved = {k: v.clone() for k, v in list(encoder.named_buffers()) + list(loss_fn.named_buffers())}
        encoder.drop_center = self.center
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._step_eager()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for p in self.params:
                p.grad = None
            n0 = lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            undo = self._arm_capture_probe() if os.environ.get("MEGB200_GRAPH_DEBUG") else None
            try:
                mode = {} if self.dist is None else {"capture_error_mode": "thread_local"}
                with torch.cuda.graph(self.graph, **mode):
                    self.loss = self._step_eager()
            finally:
                if undo:
                    undo()
            self.launches_per_step = lib.launch_count() - n0       # megb200 kernels recorded into the graph
            # the gradient tensors the recording writes; re-attached after every replay in case the caller dropped them
            # (optimizer.zero_grad(set_to_none=True), an eager step in between, ...)
            self._static_grads = [p.grad for p in self.params]
        finally:
            encoder.drop_center = self._user_center
        with torch.no_grad():
            for k, v in list(encoder.named_buffers()) + list(loss_fn.named_buffers()):
                v.copy_(saved[k])

    @staticmethod
    def _arm_capture_probe():
        """MEGB200_GRAPH_DEBUG=1: report the first megb200 op after which the stream is no longer capturing."""
        import sys

        import megb200
        ops = megb200.get_ops()
        seen = {"n": 0, "bad": False}
        saved = {}

        def wrap(name, fn):
            def w(*a, **k):
                r = fn(*a, **k)
                seen["n"] += 1
                if not seen["bad"]:
                    try:
                        ok = torch.cuda.is_current_stream_capturing()
                    except Exception as e:      # an invalidated capture raises here
                        ok = repr(e)[:120]
                    if ok is not True:
                        seen["bad"] = True
                        print(f"[megb200.graph] capture broken after op #{seen['n']} {name}: {ok}", file=sys.stderr, flush=True)
                return r
            return w

        for name in dir(ops):
            if not name.startswith("_") and callable(getattr(ops, name)):
                saved[name] = getattr(ops, name)
                setattr(ops, name, wrap(name, saved[name]))

        def undo():
            for name in saved:
                try:
                    delattr(ops, name)
                except AttributeError:
                    pass
        return undo

    def _step_eager(self):
        """zero grads -> forward -> CLIP loss -> backward, WITHOUT the autograd engine: the drop-in modules wrap exactly
        these engine calls in two autograd nodes (models.py::_EncoderFn, loss.py::_ClipFn); calling the engines directly
        records the same kernels but keeps the autograd worker thread and its end-of-backward stream synchronisation
        (which touches the legacy stream when an earlier eager step left gradient accumulators alive) out of the capture."""
        enc, lossf = self.enc, self.lossf
        with torch.no_grad():
            sa = enc.subject_block.spatial_attention
            mask = None
            if enc.training:
                mask = sa.drop_masks.index_select(0, self.center).reshape(-1)
                nbt = []
                for k in range(5):
                    cb = getattr(enc.conv_blocks, f"conv{k}")
                    nbt += [cb.batchnorm0.num_batches_tracked, cb.batchnorm1.num_batches_tracked]
                torch._foreach_add_(nbt, 1)
            params = enc._param_list()
            P = enc._param_dict(params)
            eng = enc._engine()
            ceng = lossf._engine(eng.ops.act_dtype)
            prepared = None
            if self.dist is not None:
                # fork: pack + all-gather the targets while the encoder forward runs (joined before the logits)
                cur = torch.cuda.current_stream(self.dev)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    prepared = ceng.prepare_targets(self.Y, self.dist)
            Zi, saved = eng.forward(P, self.X, self.subj, mask, enc.training)
            Fs, Fo = Zi.shape[2], enc.F
            Zl = Zi if Fs == Fo else Zi[:, :, :Fo].contiguous()
            if self.dist is not None:
                torch.cuda.current_stream(self.dev).wait_stream(self._side)
            loss, _, csaved = ceng.forward(self.Y, Zl, lossf.temp.detach(), lossf.reduction, self.dist, False,
                                           prepared=prepared, reduce_loss=False)
            dZ, dtemp = ceng.backward(csaved, self._one)
            if Fs != Fo:
                dZ = torch.nn.functional.pad(dZ, (0, Fs - Fo))
            G = eng.backward(P, saved, dZ.contiguous(), sync=self._sync)
            if self.dist is not None and "_flat" in G:
                # the gradients are views of one flat f32 buffer: SUM all-reduce it in place; the loss contribution and
                # the temperature gradient ride in its two spare floats (one collective for everything)
                G["_extra"][0:1].copy_(loss.reshape(1))
                G["_extra"][1:2].copy_(dtemp.reshape(1))
                parallel.allreduce_flat(G["_flat_late"] if self._sync.pending else G["_flat"])
                self._sync.join()
                loss, dtemp = G["_extra"][0], G["_extra"][1:2]
            for p, g in zip(params, enc._grad_list(G, None)):
                p.grad = g if g.shape == p.shape else g.reshape(p.shape)
            lossf.temp.grad = dtemp.reshape(lossf.temp.shape)
            if self.dist is not None and "_flat" not in G:
                loss = self.dist.all_reduce_sum(loss.reshape(1)).reshape(())
                parallel.allreduce_grads(self.params)           # part of the recording: replay returns summed gradients
        return loss.detach()

    def load(self, X, Y, subject_idxs, drop_center: Optional[int] = None):
        """Refill the static inputs (async on the current stream; pinned host tensors do not block)."""
        self.X.copy_(X, non_blocking=True)
        if hasattr(self.Y, "rows"):
            self.Y.rows.copy_(Y.rows, non_blocking=True)
            self.Y.sumsq.copy_(Y.sumsq, non_blocking=True)
        else:
            self.Y.copy_(Y, non_blocking=True)
        from .engine import validate_subject_idxs
        s = validate_subject_idxs(torch.as_tensor(subject_idxs), self.S)     # IndexError like ModuleList; CPU or CUDA
        self.subj.copy_(s.to(torch.int32), non_blocking=True)
        if self.enc.training:
            c = drop_center if drop_center is not None else self._user_center
            if c is None:
                c = int(np.random.randint(self.C))                                       # models.py:236
            self.center.fill_(int(c))

    def __call__(self, X=None, Y=None, subject_idxs=None, drop_center: Optional[int] = None):
        """Run one step on a new batch (or on the batch already loaded when called without arguments).  Returns the
        loss as a device scalar that the NEXT call overwrites; parameter ``.grad`` tensors are rewritten in place."""
        if X is not None:
            self.load(X, Y, subject_idxs, drop_center)
        self.graph.replay()
        for p, g in zip(self.params, self._static_grads):
            if p.grad is not g:
                p.grad = g
        return self.loss



# ================================================================================================ the unchanged loop, graphed
class _SpecLoss(torch.autograd.Function):
    """The loss tensor ``CLIPLoss.forward`` returns in graph mode.  Its value and the gradients of every parameter were
    already produced by the replayed recording (the reference loop always calls ``loss.backward()`` next: the backward
    pass is run speculatively, with upstream gradient 1); ``backward`` hands those gradients to autograd scaled by the
    actual upstream gradient.  Fresh tensors every time, so gradient accumulation over several backward calls works."""

    @staticmethod
    def forward(ctx, runner, *params):
        ctx.runner, ctx.token = runner, runner.token
        return runner.loss_out.clone()

    @staticmethod
    def backward(ctx, g):
        r = ctx.runner
        if ctx.token != r.token:
            raise RuntimeError("megb200 graph mode: backward() of a loss whose step has been overwritten by a newer forward "
                               "(the recorded buffers hold one step); call loss.backward() before the next batch")
        scaled = torch._foreach_mul(r.static_grads, g.reshape(()).to(torch.float32))
        if r.absent:                       # subjects without a sample in this batch: grad None, as the reference / eager path
            scaled = [None if i in r.absent else t for i, t in enumerate(scaled)]
        return (None,) + tuple(scaled)


class ModuleGraphs:
    """The reference's UNCHANGED training loop (train.py:185-215: ``Z = brain_encoder(X, subject_idxs)``, ``loss =
    loss_func(Y, Z)``, ``classifier(Z, Y)``, ``loss.item()``, ``optimizer.zero_grad(); loss.backward(); optimizer.step()``)
    at CUDA-graph speed: the ~145 kernel launches of a step cost 6-10 ms of host time when enqueued one by one from
    Python, against ~7 ms of device time.  Installed on an encoder (``megb200.graph.accelerate(enc)``, or
    ``MEGB200_MODULE_GRAPHS=1`` / ``args.cuda_graphs`` at construction), it lets the first steps of a new input shape
    run eagerly (real training steps; they also create tensor maps and kernel attributes), then records

      G1  BrainEncoder.forward                                   replayed by ``brain_encoder(X, subject_idxs)``
      G2  CLIPLoss forward + backward, BrainEncoder backward     replayed by ``loss_func(Y, Z)``

    -- the same C-ABI launches, captured -- and ``loss.backward()`` only hands the recorded gradients to autograd
    (_SpecLoss).  Assumptions, all checked: training mode with gradients enabled, ``loss_func`` receives the Z just
    returned, ``fast=True`` / no ``return_logits``, one backward per forward.  Anything else (eval, no_grad, other
    shapes, other uses of Z) takes the eager path; Z carries no autograd history in graph mode, so using it in any
    differentiable way other than ``loss_func(Y, Z)`` raises.  Subjects absent from a batch receive ``grad = None`` as in
    the reference (the recording computes zeros for them; they are dropped when the gradients are handed out).  Data parallel works (collectives are part of G2) but the target gather is then
    not overlapped with the forward; GraphedTrainStep is the data-parallel fast path."""

    WARMUP = 3

    def __init__(self, encoder):
        self.enc = encoder
        self.key, self.eager_done = None, 0
        self.g1 = self.g2 = None
        self.token = 0
        self.absent = set()
        self.enabled = True
        self.stats = {"eager_steps": 0, "graph_steps": 0, "captures": 0}

    # ------------------------------------------------------------------ plumbing
    def note_eager_backward(self):
        """Called by the eager autograd node after a complete eager step (forward + loss + backward): only such steps count
        as warm-up, because they are what creates the tensor maps / descriptor tables the recordings must not create."""
        self.eager_done += 1

    def _reset(self, key):
        self.key, self.eager_done = key, 0
        self.g1 = self.g2 = None
        self.lossf = None
        self.static_grads = self.loss_out = None

    def _capture(self, fn):
        dev = self.dev
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        mode = {} if parallel.context() is None else {"capture_error_mode": "thread_local"}
        kw = dict(pool=self.pool) if self.pool is not None else {}
        with torch.cuda.graph(g, **kw, **mode):
            out = fn()
        if self.pool is None:
            self.pool = g.pool()
        self.stats["captures"] += 1
        return g, out

    # ------------------------------------------------------------------ G1: encoder forward
    def encoder_forward(self, X, subject_idxs):
        """Returns Z (the reference's (B, F, T) view or (B, F)), or None when this call has to run eagerly."""
        enc = self.enc
        if not self.enabled or not torch.is_tensor(X) or not X.is_cuda:
            return None
        key = (tuple(X.shape), str(X.device), megb200_precision(), id(enc._engine()))
        if key != self.key:
            self._reset(key)
        if self.eager_done < self.WARMUP:
            self.stats["eager_steps"] += 1
            return None
        from .engine import validate_subject_idxs
        dev = X.device
        if self.g1 is None:
            self.dev, self.pool = dev, None
            self.X = torch.empty(tuple(X.shape), dtype=torch.float32, device=dev)
            self.subj = torch.zeros(int(X.shape[0]), dtype=torch.int32, device=dev)
            self.center = torch.zeros(1, dtype=torch.int64, device=dev)
            self._one = torch.ones(1, dtype=torch.float32, device=dev)
        self.X.copy_(X, non_blocking=True)
        s = validate_subject_idxs(torch.as_tensor(subject_idxs), enc.num_subjects)
        self.subj.copy_(s.to(torch.int32), non_blocking=True)
        present = set(int(v) for v in (torch.unique(s) if s.is_cuda else s).tolist())
        # positions (in _SpecLoss's parameter list) of the subject layers that see no sample in this batch
        self.absent = {3 + k for k in range(enc.num_subjects) if k not in present}
        c = enc.drop_center
        if torch.is_tensor(c):
            self.center.copy_(c.reshape(1))
        else:
            self.center.fill_(int(c) if c is not None else int(np.random.randint(int(X.shape[1]))))   # models.py:236
        if self.g1 is None:
            self.g1, _ = self._capture(self._record_forward)
            self.g2 = None
        self.g1.replay()
        self.token += 1
        self.stats["graph_steps"] += 1
        Zi = self.Zi
        Z = Zi.permute(0, 2, 1) if enc.seq2seq else Zi[:, 0, : enc.F]
        Z = _GraphZ.apply(Z, self.enc.conv_final2.bias)       # a leash: any differentiable use other than the loss raises
        Z._megb200_runner = (self, self.token)
        return Z

    def _record_forward(self):
        enc = self.enc
        with torch.no_grad():
            sa = enc.subject_block.spatial_attention
            mask = sa.drop_masks.index_select(0, self.center).reshape(-1)
            nbt = []
            for k in range(5):
                cb = getattr(enc.conv_blocks, f"conv{k}")
                nbt += [cb.batchnorm0.num_batches_tracked, cb.batchnorm1.num_batches_tracked]
            torch._foreach_add_(nbt, 1)
            self.params_enc = enc._param_list()
            self.P = enc._param_dict(self.params_enc)
            self.eng = enc._engine()
            self.Zi, self.saved = self.eng.forward(self.P, self.X, self.subj, mask, True)

    # ------------------------------------------------------------------ G2: loss forward + whole backward
    def loss_forward(self, lossf, x, token):
        if token != self.token:
            raise RuntimeError("megb200 graph mode: loss_func(Y, Z) received the Z of an earlier step")
        if self.g2 is None or self.lossf is not lossf or self._ykey != self._target_key(x):
            self.lossf, self._ykey = lossf, self._target_key(x)
            Y = lossf._as_targets(x, self.dev)
            if hasattr(Y, "rows"):
                from .data import PackedTargets
                self.Y = PackedTargets(torch.empty_like(Y.rows), torch.empty_like(Y.sumsq), Y.feature_shape)
            else:
                self.Y = torch.empty_like(Y)
            self._load_targets(Y)
            self.g2, _ = self._capture(self._record_backward)
        else:
            self._load_targets(lossf._as_targets(x, self.dev))
        self.g2.replay()
        params = [p for p in self.params_enc] + [lossf.temp]
        return _SpecLoss.apply(self, *params)

    @staticmethod
    def _target_key(x):
        if hasattr(x, "rows"):
            return ("packed", tuple(x.rows.shape), x.rows.dtype)
        return ("dense", tuple(x.shape), torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32)

    def _load_targets(self, Y):
        if hasattr(self.Y, "rows"):
            self.Y.rows.copy_(Y.rows, non_blocking=True)
            self.Y.sumsq.copy_(Y.sumsq, non_blocking=True)
        else:
            self.Y.copy_(Y, non_blocking=True)

    def _record_backward(self):
        enc, lossf, dist_ctx = self.enc, self.lossf, parallel.context()
        with torch.no_grad():
            ceng = lossf._engine(self.eng.ops.act_dtype)
            Zi = self.Zi
            Fs, Fo = Zi.shape[2], enc.F
            Zl = Zi if Fs == Fo else Zi[:, :, :Fo].contiguous()
            loss, _, csaved = ceng.forward(self.Y, Zl, lossf.temp.detach(), lossf.reduction, dist_ctx, False,
                                           reduce_loss=False)
            dZ, dtemp = ceng.backward(csaved, self._one)
            if Fs != Fo:
                dZ = torch.nn.functional.pad(dZ, (0, Fs - Fo))
            G = self.eng.backward(self.P, self.saved, dZ.contiguous())
            if dist_ctx is not None:
                if "_flat" in G:
                    G["_extra"][0:1].copy_(loss.reshape(1))
                    G["_extra"][1:2].copy_(dtemp.reshape(1))
                    parallel.allreduce_flat(G["_flat"])
                    loss, dtemp = G["_extra"][0], G["_extra"][1:2]
                else:
                    raise RuntimeError("megb200 graph mode: gradient pool spilled in data parallel")
            grads = []
            for p, g in zip(self.params_enc, enc._grad_list(G, None)):
                grads.append(g if g.shape == p.shape else g.reshape(p.shape))
            grads.append(dtemp.reshape(lossf.temp.shape))
            self.static_grads = grads
            self.loss_out = loss.reshape(())


class _GraphZ(torch.autograd.Function):
    """Identity on the recorded embeddings that ties them to autograd: differentiating through Z in graph mode (anything
    but ``loss_func(Y, Z)``, which never backpropagates through this node) is an error, not silently-zero gradients."""

    @staticmethod
    def forward(ctx, Z, anchor):
        return Z.view_as(Z)

    @staticmethod
    def backward(ctx, g):
        raise RuntimeError("megb200 graph mode (ModuleGraphs): the embeddings Z are only differentiable through "
                           "loss_func(Y, Z); disable it with megb200.graph.accelerate(enc, False) for other uses")


def megb200_precision():
    import megb200
    return megb200.get_precision()


def accelerate(encoder, on: bool = True):
    """Install (or remove) CUDA-graph replay of the unchanged training loop on a drop-in BrainEncoder.  Returns the
    ModuleGraphs runner (``runner.stats`` counts eager / graph steps and captures)."""
    if not on:
        encoder._runner = None
        return None
    if getattr(encoder, "_runner", None) is None:
        encoder._runner = ModuleGraphs(encoder)
    return encoder._runner
