"""Drop-in ``meg_decoding.models`` for the training hot path: same constructor / forward signatures and
``state_dict`` layout as the reference (``meg_decoding/models.py:167-383``), computed by the megb200 kernels.

    brain_encoder = BrainEncoder(args).to(device)        # train.py:147
    Z = brain_encoder(X, subject_idxs)                   # train.py:201

The sub-modules (``nn.Conv1d`` / ``nn.BatchNorm1d``) are parameter containers only: they give the reference's
parameter names, shapes, init distribution, ``.to()`` / ``.train()`` / ``state_dict()`` behaviour; the arithmetic is
one explicit kernel sequence (``megb200.engine.EncoderEngine``) inside a single autograd node.  Without the built
CUDA library the forward raises — there is no PyTorch fallback.

Intentional differences from the reference, all documented in DESIGN.md: ``z`` is a properly registered complex
Parameter on every device (the reference loses it on CUDA, models.py:186-188); the spatial-dropout centre can be
pinned through ``drop_center`` (the reference draws ``np.random.randint`` per call, models.py:236 — still the default).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

import megb200
from megb200.engine import EncoderEngine, validate_subject_idxs
from meg_decoding.utils.layout import ch_locations_2d


def get_model(args):
    """models.py:18-30; only the BrainEncoder branch is on the accelerated path."""
    if args.model == "brain_encoder":
        return BrainEncoder(args)
    if args.model == "brain_endcoder_seq2static":            # the reference's spelling (models.py:21)
        return BrainEncoderSeq2Static(args)
    raise ValueError("no model named {} is prepared by megb200 (only 'brain_encoder' is on the hot path)".format(args.model))


class SpatialDropout(nn.Module):
    """models.py:223-241.  Holds the sensor layout; the mask is folded into the attention weights by the engine."""

    def __init__(self, loc, d_drop):
        super().__init__()
        self.loc = loc
        self.d_drop = d_drop
        self.num_channels = loc.shape[0]

    def all_masks(self):
        """masks[c] = mask used when sensor c is the centre: 0 where ||loc - loc_c|| < d_drop (strict), else 1."""
        d = (self.loc[None, :, :] - self.loc[:, None, :]).norm(dim=-1)
        return torch.where(d < self.d_drop, 0.0, 1.0).to(torch.float32)


class SpatialAttention(nn.Module):
    """models.py:167-220."""

    def __init__(self, args):
        super().__init__()
        K = args.K
        loc = torch.as_tensor(ch_locations_2d(args), dtype=torch.float32)
        kk = torch.arange(K).repeat_interleave(K).to(torch.float32)
        ll = torch.arange(K).repeat(K).to(torch.float32)
        self.z = nn.Parameter(torch.rand(size=(args.D1, K ** 2), dtype=torch.cfloat))
        phi = 2 * torch.pi * (torch.einsum("k,x->kx", kk, loc[:, 0]) + torch.einsum("l,y->ly", ll, loc[:, 1]))
        # like the reference these are plain (non-persistent) tensors: not part of state_dict
        self.register_buffer("cos", torch.cos(phi).contiguous(), persistent=False)
        self.register_buffer("sin", torch.sin(phi).contiguous(), persistent=False)
        self.spatial_dropout = SpatialDropout(loc, args.d_drop)
        self.register_buffer("drop_masks", self.spatial_dropout.all_masks().contiguous(), persistent=False)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """A checkpoint the reference wrote while running on CUDA has no ``z`` key: there ``nn.Parameter(...).to(device)``
        turns z into a plain tensor that is neither trained nor saved (models.py:186-188).  Such a checkpoint loads
        here with z left at its current value and a warning, instead of a missing-key error.  torch hands this hook only
        the keys under its own prefix, so whether the dict is a checkpoint of the subject block at all (rather than a
        partial non-strict load of other keys, which stays silent) is noted by ``BrainEncoder.load_state_dict``."""
        if prefix + "z" not in state_dict:
            if getattr(self, "_warn_missing_z", True):
                import warnings
                warnings.warn("megb200: checkpoint has no '{}z' (written by the reference on CUDA, where z is not a "
                              "registered parameter); keeping the current z".format(prefix))
            state_dict = dict(state_dict)
            state_dict[prefix + "z"] = self.z.detach()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class SubjectBlock(nn.Module):
    """models.py:244-273."""

    def __init__(self, args):
        super().__init__()
        self.num_subjects = args.num_subjects
        self.D1 = args.D1
        self.K = args.K
        self.spatial_attention = SpatialAttention(args)
        self.conv = nn.Conv1d(self.D1, self.D1, kernel_size=1, stride=1)
        self.subject_layer = nn.ModuleList(
            [nn.Conv1d(self.D1, self.D1, kernel_size=1, bias=False, stride=1) for _ in range(self.num_subjects)])


class ConvBlock(nn.Module):
    """models.py:276-322 (dilation 1 everywhere, as the reference code runs: :289,297,305)."""

    def __init__(self, k, D1, D2, ks=3):
        super().__init__()
        self.k = k
        self.D2 = D2
        self.in_channels = D1 if k == 0 else D2
        self.ks = ks
        self.conv0 = nn.Conv1d(self.in_channels, D2, kernel_size=ks, padding="same")
        self.batchnorm0 = nn.BatchNorm1d(num_features=D2)
        self.conv1 = nn.Conv1d(D2, D2, kernel_size=ks, padding="same")
        self.batchnorm1 = nn.BatchNorm1d(num_features=D2)
        self.conv2 = nn.Conv1d(D2, 2 * D2, kernel_size=ks, padding="same")


class _EncoderFn(torch.autograd.Function):
    """One autograd node for the whole encoder: forward and backward are explicit kernel sequences."""

    @staticmethod
    def forward(ctx, module, X, subj_dev, mask, present, *params):
        P = module._param_dict(params)
        eng = module._engine()
        Z, saved = eng.forward(P, X, subj_dev, mask, module.training)
        ctx.module, ctx.saved, ctx.P, ctx.present = module, saved, P, present
        ctx.was_training = module.training
        ctx.nparams = len(params)
        return Z

    @staticmethod
    def backward(ctx, dZ):
        module = ctx.module
        if not ctx.was_training:
            raise NotImplementedError("megb200: backward through an eval-mode BrainEncoder forward is not implemented "
                                      "(the reference scripts evaluate under torch.no_grad())")
        eng = module._engine()
        ops = eng.ops
        if dZ.dtype != ops.act_dtype or not dZ.is_contiguous():
            dZ = dZ.to(ops.act_dtype).contiguous()
        G = eng.backward(ctx.P, ctx.saved, dZ)
        grads = module._grad_list(G, ctx.present)
        ctx.saved = None
        if module._runner is not None:
            module._runner.note_eager_backward()
        return (None, None, None, None, None) + tuple(grads)


def _arg(args, name, default=None):
    try:
        return args.get(name, default) if hasattr(args, "get") else getattr(args, name, default)
    except Exception:
        return default


class BrainEncoder(nn.Module):
    """models.py:341-383.

    ``args.conv_dilations`` (optional megb200 extension; 5 x 3 integers or ``"paper"``) enables the dilated block of the
    paper, which the reference keeps as an alternative formulation (tests/test_modules/models.py:267-312) while its
    ``ConvBlock`` runs with dilation 1 (models.py:289,297,305 commented out) -- the default here too."""

    pooled = False                       # BrainEncoderSeq2Static: pooling between the blocks

    def _kernel_sizes(self, args):
        return [3] * 5

    def __init__(self, args):
        super().__init__()
        self.num_subjects = args.num_subjects
        self.D1 = args.D1
        self.D2 = args.D2
        self.F = args.F if not args.preprocs["last4layers"] else 1024
        self.K = args.K
        self.dataset_name = args.dataset
        self.seq2seq = bool(args.seq2seq)
        if self.F % 8 != 0:
            raise ValueError("megb200: the embedding width F must be a multiple of 8")

        self.subject_block = SubjectBlock(args)
        self.conv_blocks = nn.Sequential()
        self.ks = self._kernel_sizes(args)
        for k in range(5):
            self.conv_blocks.add_module(f"conv{k}", ConvBlock(k, self.D1, self.D2, ks=self.ks[k]))
            if self.pooled:                                                      # models.py:483-487 (no parameters)
                if k < 4:
                    self.conv_blocks.add_module(f"avgpool{k}", nn.AvgPool1d(3, stride=2))
                else:
                    self.conv_blocks.add_module(f"globalavgpool{k}", nn.AdaptiveAvgPool1d(output_size=1))
        dil = _arg(args, "conv_dilations")
        if isinstance(dil, str):
            if dil != "paper":
                raise ValueError(f"unknown conv_dilations {dil!r}")
            dil = [(2 ** (2 * k % 5), 2 ** (2 * k + 1 % 5), 2) for k in range(1, 6)]   # tests/test_modules/models.py:275-297
        self.dilations = [tuple(int(x) for x in row) for row in dil] if dil is not None else None
        self.conv_final1 = nn.Conv1d(self.D2, 2 * self.D2, kernel_size=1)
        self.conv_final2 = nn.Conv1d(2 * self.D2, self.F, kernel_size=1)

        try:
            self.precision = args.get("precision") if hasattr(args, "get") else getattr(args, "precision", None)
        except Exception:
            self.precision = None
        # dtype of the returned embeddings: None = the precision mode's storage type (bf16 by default).  The reference's
        # evaluation / inference scripts call ``Z.detach().cpu().numpy()`` (evaluate.py:340, examples/inference.py:105),
        # which numpy cannot do for bf16: set ``args.output_dtype = "float32"`` (or ``enc.output_dtype = torch.float32``)
        # there -- one extra cast pass (megb200_cast) over Z, outside autograd (eval / no_grad only).
        od = _arg(args, "output_dtype")
        self.output_dtype = getattr(torch, od) if isinstance(od, str) else od
        self.drop_center = None          # set to an int to pin the spatial-dropout centre (tests, data parallel)
        self._eng = None
        # CUDA-graph replay of the unchanged training loop (megb200.graph.ModuleGraphs): opt-in
        self._runner = None
        import os
        if _arg(args, "cuda_graphs") or os.environ.get("MEGB200_MODULE_GRAPHS", "0") == "1":
            from megb200.graph import accelerate
            accelerate(self)

    # ------------------------------------------------------------------ plumbing
    def _engine(self) -> EncoderEngine:
        ops = megb200.get_ops(self.precision)
        if self._eng is None or self._eng.ops is not ops:
            C = self.subject_block.spatial_attention.cos.shape[1]
            self._eng = EncoderEngine(ops, C=C, D1=self.D1, D2=self.D2, F=self.F, S=self.num_subjects,
                                      seq2seq=self.seq2seq, dilations=self.dilations, ks=self.ks, pooled=self.pooled)
        return self._eng

    def _param_list(self):
        sb = self.subject_block
        ps = [sb.spatial_attention.z, sb.conv.weight, sb.conv.bias]
        ps += [m.weight for m in sb.subject_layer]
        for k in range(5):
            cb = getattr(self.conv_blocks, f"conv{k}")
            ps += [cb.conv0.weight, cb.conv0.bias, cb.batchnorm0.weight, cb.batchnorm0.bias,
                   cb.conv1.weight, cb.conv1.bias, cb.batchnorm1.weight, cb.batchnorm1.bias,
                   cb.conv2.weight, cb.conv2.bias]
        ps += [self.conv_final1.weight, self.conv_final1.bias, self.conv_final2.weight, self.conv_final2.bias]
        return ps

    def _param_dict(self, params):
        """Flat autograd inputs -> the engine's parameter dict (f32, contiguous; BN buffers by reference)."""
        S, D1 = self.num_subjects, self.D1
        sa = self.subject_block.spatial_attention
        it = iter(params)
        z = next(it).detach()
        zr = torch.view_as_real(z)
        P = {"z_re": zr[..., 0].contiguous(), "z_im": zr[..., 1].contiguous(), "cos": sa.cos, "sin": sa.sin}
        P["Wc"] = next(it).detach().reshape(D1, D1)
        P["bc"] = next(it).detach()
        P["Ws"] = torch.stack([next(it).detach().reshape(D1, D1) for _ in range(S)])
        P["blocks"] = []
        for k in range(5):
            cb = getattr(self.conv_blocks, f"conv{k}")
            d = {}
            for i, bn in (("0", cb.batchnorm0), ("1", cb.batchnorm1)):
                d["w" + i], d["b" + i] = next(it).detach(), next(it).detach()
                d["g" + i], d["be" + i] = next(it).detach(), next(it).detach()
                d["rm" + i], d["rv" + i] = bn.running_mean, bn.running_var
            d["w2"], d["b2"] = next(it).detach(), next(it).detach()
            P["blocks"].append(d)
        P["wf1"], P["bf1"], P["wf2"], P["bf2"] = (next(it).detach() for _ in range(4))
        return P

    def _grad_list(self, G, present):
        """Engine gradient dict -> gradients in ``_param_list`` order (reference shapes / dtypes)."""
        S, D1 = self.num_subjects, self.D1
        out = [torch.complex(G["z_re"], G["z_im"]), G["Wc"].view(D1, D1, 1), G["bc"]]
        for s in range(S):
            # the reference leaves grad=None for subjects absent from the batch
            out.append(G["Ws"][s].view(D1, D1, 1) if (present is None or s in present) else None)
        for k in range(5):
            g = G["blocks"][k]
            out += [g["w0"], g["b0"], g["g0"], g["be0"], g["w1"], g["b1"], g["g1"], g["be1"], g["w2"], g["b2"]]
        out += [G["wf1"], G["bf1"], G["wf2"], G["bf2"]]
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, X, subject_idxs):
        return self._forward(X, subject_idxs)

    def _forward(self, X, subject_idxs):
        sa = self.subject_block.spatial_attention
        dev = sa.cos.device
        C = sa.cos.shape[1]
        assert X.shape[1] == C                                                  # models.py:233
        if self._runner is not None and self.training and torch.is_grad_enabled() and dev.type == "cuda":
            Z = self._runner.encoder_forward(X.to(dev) if torch.is_tensor(X) else X, subject_idxs)
            if Z is not None:
                return Z
        S = self.num_subjects
        present = None
        subject_idxs = torch.as_tensor(subject_idxs)
        # ModuleList semantics (models.py:270-272): IndexError outside [-S, S), negative indices wrap.  The reference
        # keeps subject_idxs on the CPU (train.py:199 moves X and Y only); a device-resident index tensor is validated
        # too -- one small device->host read -- because the grouped kernels index weights and shared memory with it.
        # (megb200.graph.GraphedTrainStep validates in load(), outside the captured region.)
        subject_idxs = validate_subject_idxs(subject_idxs, S)
        present = set(int(s) for s in (torch.unique(subject_idxs) if subject_idxs.is_cuda else subject_idxs).tolist())
        subj = subject_idxs.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        X = X.to(device=dev, dtype=torch.float32).contiguous()
        mask = None
        if self.training:
            # models.py:236 — one centre per forward call from the global NumPy RNG unless pinned
            c = self.drop_center if self.drop_center is not None else int(np.random.randint(C))
            if torch.is_tensor(c):      # device-resident centre (megb200.graph: the step is replayed from a CUDA graph)
                mask = sa.drop_masks.index_select(0, c.reshape(1).to(torch.int64)).reshape(-1)
            else:
                mask = sa.drop_masks[c]
            nbt = []
            for k in range(5):
                cb = getattr(self.conv_blocks, f"conv{k}")
                nbt += [cb.batchnorm0.num_batches_tracked, cb.batchnorm1.num_batches_tracked]
            torch._foreach_add_(nbt, 1)            # one launch for the ten counters instead of ten
        params = self._param_list()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            Zi = _EncoderFn.apply(self, X, subj, mask, present, *params)
        else:
            Zi, _ = self._engine().forward(self._param_dict(params), X, subj, mask, self.training)
            if self.output_dtype is not None and self.output_dtype != Zi.dtype:
                Zo = torch.empty(Zi.shape, dtype=self.output_dtype, device=Zi.device)
                self._engine().ops.cast(Zi, Zo)
                Zi = Zo
        # PC layout [B,T,F] -> the reference's (B, F, T) (a view) or (B, F)
        if self.seq2seq:
            return Zi.permute(0, 2, 1)
        return Zi[:, 0, : self.F]

    def load_state_dict(self, state_dict, strict=True, assign=False):
        # see SpatialAttention._load_from_state_dict: warn about an absent z only for dicts that cover the subject block
        sa = self.subject_block.spatial_attention
        sa._warn_missing_z = any(k.startswith("subject_block.") for k in state_dict)
        try:
            return super().load_state_dict(state_dict, strict=strict, assign=assign)
        finally:
            sa._warn_missing_z = True


class BrainEncoderSeq2Static(BrainEncoder):
    """models.py:465-512 (``args.model == "brain_endcoder_seq2static"``, the GOD image-decoding configs): the same
    SubjectBlock / ConvBlock / 1x1 head, but block k convolves with kernel size ``args.ConvBlocks.ks[k]`` (3, 5 or 11 in
    the reference's configs), ``AvgPool1d(3, stride=2)`` follows blocks 0..3 and ``AdaptiveAvgPool1d(1)`` block 4, so the
    head runs on one time step.  Returns (B, F) for ``seq2seq=False`` (the mean over the single step, :510) and
    (B, F, 1) otherwise.  Same state_dict keys as the reference (the pooling layers have no parameters)."""

    pooled = True

    def _kernel_sizes(self, args):
        cb = _arg(args, "ConvBlocks")
        ks = cb["ks"] if isinstance(cb, dict) else cb.ks
        ks = [int(k) for k in ks]
        if len(ks) != 5:
            raise ValueError("args.ConvBlocks.ks must list five kernel sizes")
        return ks


class Classifier(nn.Module):
    """models.py:386-460 — top-1 / top-10 retrieval accuracy: for every speech segment, is the matching brain embedding
    the most (one of the 10 most) cosine-similar among the batch?  Metric only (``no_grad``), called every step
    (train.py:205-206).  "Next" row N1 of the scope table: the reference's O(B^2) Python loop of dot products becomes the
    tcgen05 logits contraction plus one rank kernel (megb200_retrieval_rank); same return values."""

    def __init__(self, args):
        super().__init__()
        self.factor = 1
        try:
            self.normalize_image_features = args.get("normalize_image_features", False) if hasattr(args, "get") else \
                getattr(args, "normalize_image_features", False)
        except Exception:
            self.normalize_image_features = False
        self._eng = None

    @torch.no_grad()
    def forward(self, Z: torch.Tensor, Y: torch.Tensor, test=False, top_k=None):
        from megb200.engine import ClipEngine
        batch_size = Z.size(0)
        z = Z.detach()
        if z.dtype not in (torch.bfloat16, torch.float32):
            z = z.float()
        if z.dim() == 3:
            Zi = z.permute(0, 2, 1)
            if not Zi.is_contiguous():
                Zi = Zi.contiguous()
        else:
            Zi = z.contiguous().unsqueeze(1)
        ops = megb200.get_ops("bf16" if Zi.dtype == torch.bfloat16 and megb200.get_precision() != "bf16_simt"
                              else ("bf16_simt" if Zi.dtype == torch.bfloat16 else "fp32"))
        if self._eng is None or self._eng.ops is not ops:
            self._eng = ClipEngine(ops)
        Yf = Y.detach().to(device=Zi.device, dtype=torch.float32).contiguous()
        rank = self._eng.retrieval_ranks(Yf, Zi)
        top1 = (rank == 0).float().mean().item()
        top10 = (rank < min(10, batch_size)).float().mean().item()
        if top_k is None:
            return top1, top10
        return top1, top10, (rank < min(int(top_k), batch_size)).float().mean().item()
