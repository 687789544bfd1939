"""GPU tier, op level: every C-ABI kernel against the oracle's plain-torch restatement of the same op on identical
seeded inputs.  fp32 storage is compared tightly; bf16 storage within one output rounding.  The contraction ops are
run on both engines (SIMT fp32 CUDA cores and tcgen05 tensor cores)."""
import pytest
import torch

from megb200 import lib as L
from megb200.ops import pad8, pad16, pad64
from oracle.ops_restatement import TorchOps

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops(dtype, backend):
    from megb200.ops import CudaOps
    return CudaOps(dtype, backend)


def _rand(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).float()


def _act(gen, dtype, B, T, C, Cs=None):
    """Activation tensor [B,T,Cs] with zero pad channels; returns (cpu tensor in `dtype`, cuda copy)."""
    Cs = Cs or pad8(C)
    t = torch.zeros(B, T, Cs)
    t[:, :, :C] = _rand(gen, B, T, C)
    t = t.to(dtype)
    return t, t.to(DEV)


def _close(got, ref, dtype, what, atol_scale=1.0):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    scale = float(ref.abs().max()) + 1e-30
    tol = (2e-5 if dtype == torch.float32 else 1.2e-2) * atol_scale
    err = float((got - ref).abs().max()) / scale
    assert err <= tol, f"{what}: max abs err / max|ref| = {err:.3e} > {tol:.1e}"


CONTRACTION_MODES = [(torch.float32, "simt"), (torch.bfloat16, "simt"), (torch.bfloat16, "tc")]
EW_DTYPES = [torch.float32, torch.bfloat16]


# ------------------------------------------------------------------------------------------------ packing
@pytest.mark.parametrize("dtype", EW_DTYPES)
def test_pack_x_and_weight(dtype):
    gen = torch.Generator().manual_seed(0)
    ops, ref = _ops(dtype, "simt"), TorchOps(dtype)
    X = _rand(gen, 3, 20, 45)
    Xp = ops.empty(3, 45, 24, device=DEV)
    ops.pack_x(X.to(DEV), Xp)
    Xr = ref.empty(3, 45, 24, device="cpu")
    ref.pack_x(X, Xr)
    _close(Xp, Xr, dtype, "pack_x")
    for mode, (O, I, Kt) in ((0, (30, 20, 3)), (1, (30, 20, 3)), (0, (270, 208, 1))):
        w = _rand(gen, 2, O, I, Kt)
        N, K = (O, I) if mode == 0 else (I, O)
        out = ops.empty(2, pad16(N), Kt * pad64(K), device=DEV)
        ops.pack_weight(w.to(DEV), out, mode)
        outr = ref.empty(2, pad16(N), Kt * pad64(K), device="cpu")
        ref.pack_weight(w, outr, mode)
        _close(out, outr, dtype, f"pack_weight mode {mode}")


# ------------------------------------------------------------------------------------------------ contractions
TAPGEMM_CASES = [
    # B, T, Ca, N, taps, dil, G, bias, res
    (3, 40, 20, 30, 1, 1, 4, True, False),        # grouped 1x1 (subject GEMM), ragged dims
    (2, 40, 30, 32, 3, 1, 1, True, False),        # block-0 conv0
    (2, 40, 32, 32, 3, 1, 1, True, True),         # residual conv
    (2, 200, 32, 64, 3, 2, 1, True, False),       # two M tiles per sample, dilation 2, GLU conv width
    (2, 360, 270, 320, 3, 1, 1, True, False),     # real widths: K chunk with OOB channels, N = 2 x 160
    (2, 360, 320, 640, 3, 1, 1, True, False),
    (1, 360, 640, 1024, 1, 1, 1, True, False),    # head conv, N tile 256
    (3, 130, 208, 270, 1, 1, 5, True, False),     # subject GEMM at real widths (N = 272 -> 2 x 144)
    (2, 129, 320, 320, 3, 1, 1, False, True),     # dgrad-like: no bias, residual, T = 128 + 1
    # --- shapes that cross the persistent CTA-pair kernels' work threshold NATIVELY (>= 2 x 148 tile items), i.e. the
    # kernels config 2 runs (tc_tapgemm2h_kernel / tc_tapgemm2_kernel), with no option forced
    (64, 360, 320, 320, 3, 1, 1, False, True),    # conv data gradient: STORE + residual, pair + halo (384 items)
    (51, 360, 320, 320, 3, 2, 1, True, False),    # dilation 2 halo, odd number of M tiles (153: last pair half empty)
    (50, 360, 640, 320, 1, 1, 1, False, False),   # head data gradient 640 -> 320: plain pair kernel (taps = 1)
]

# which tensor-core kernel each natively-large case must have launched (megb200_get_option("last_tapgemm_kernel"))
PAIR_EXPECT = {(64, 360, 320, 320, 3, 1): 2, (51, 360, 320, 320, 3, 2): 2, (50, 360, 640, 320, 1, 1): 1,
               (80, 200, 320, 320, 11, 1): 1, (64, 360, 320, 320, 3, 128): 1}      # wide kernel / long halo: plain pair


@pytest.mark.parametrize("dtype,backend", CONTRACTION_MODES)
@pytest.mark.parametrize("case", TAPGEMM_CASES)
def test_tapgemm(dtype, backend, case):
    B, T, Ca, N, taps, dil, G, has_bias, has_res = case
    gen = torch.Generator().manual_seed(1)
    ops, ref = _ops(dtype, backend), TorchOps(dtype)
    A, Ad = _act(gen, dtype, B, T, Ca)
    W = torch.zeros(G, pad16(N), taps, pad64(Ca))
    W[:, :N, :, :Ca] = _rand(gen, G, N, taps, Ca, scale=(Ca * taps) ** -0.5)
    W = W.view(G, pad16(N), taps * pad64(Ca)).to(dtype)
    bias = _rand(gen, G, N) if has_bias else None
    Res, Resd = _act(gen, dtype, B, T, N) if has_res else (None, None)
    group = torch.randint(0, G, (B,), generator=gen, dtype=torch.int32) if G > 1 else None
    Out = ops.empty(B, T, pad8(N), device=DEV)
    ops.tapgemm(Ad, Ca, W.to(DEV), N, taps, dil, None if bias is None else bias.to(DEV), Resd,
                None if group is None else group.to(DEV), Out)
    Outr = ref.empty(B, T, pad8(N), device="cpu")
    ref.tapgemm(A, Ca, W, N, taps, dil, bias, Res, group, Outr)
    torch.cuda.synchronize()
    _close(Out, Outr, dtype, f"tapgemm {case} {backend}")
    if backend == "tc" and case[:6] in PAIR_EXPECT and L.get_option("tapgemm_pair_min_items") < 0 \
            and not L.get_option("tapgemm_force_1cta"):
        assert L.get_option("last_tapgemm_kernel") == PAIR_EXPECT[case[:6]], "the product (CTA-pair) kernel did not run"


EPI_CASES = [
    # B, T, Ca, N, taps, epi
    (2, 40, 32, 32, 3, "store_f32"),
    (2, 360, 320, 320, 3, "store_f32"),
    (5, 200, 320, 320, 3, "store_f32"),      # odd number of M tiles: the CTA-pair kernel's last pair is half empty
    (2, 40, 32, 64, 3, "glu"),              # D2 = 32
    (2, 360, 320, 640, 3, "glu"),           # D2 = 320: tile = 80 a-channels + 80 gates
    (3, 130, 32, 64, 1, "gelu"),
    (2, 360, 320, 640, 1, "gelu"),
    (1, 360, 640, 1024, 1, "gelu"),
    # --- natively above the CTA-pair threshold: the product forms of every fused epilogue (config 2's kernels)
    (64, 360, 320, 320, 3, "store_f32"),     # conv1: f32 store + residual + BatchNorm statistics, pair + halo
    (64, 360, 270, 320, 3, "store_f32_nores"),  # block-0 conv0: K = 270 (OOB channels in the last k chunk), no residual
    (50, 360, 320, 640, 3, "glu"),           # conv2 + GLU, pair + halo (750 items)
    (64, 360, 320, 640, 1, "gelu"),          # head conv 1, pair
    (32, 360, 640, 1024, 1, "gelu"),         # head conv 2: 1024 wide, N tile 256, pair
]
EPI_PAIR_EXPECT = {(64, 360, 320, 320, 3): 2, (64, 360, 270, 320, 3): 2, (50, 360, 320, 640, 3): 2,
                   (64, 360, 320, 640, 1): 1, (32, 360, 640, 1024, 1): 1}


@pytest.mark.parametrize("dtype,backend", CONTRACTION_MODES)
@pytest.mark.parametrize("case", EPI_CASES)
def test_tapgemm_fused_epilogues(dtype, backend, case):
    B, T, Ca, N, taps, epi = case
    gen = torch.Generator().manual_seed(11)
    ops, ref = _ops(dtype, backend), TorchOps(dtype)
    A, Ad = _act(gen, dtype, B, T, Ca)
    W = torch.zeros(1, pad16(N), taps, pad64(Ca))
    W[:, :N, :, :Ca] = _rand(gen, 1, N, taps, Ca, scale=(Ca * taps) ** -0.5)
    W = W.view(1, pad16(N), taps * pad64(Ca)).to(dtype)
    bias = _rand(gen, 1, N)
    Res, Resd = (None, None)
    if epi == "store_f32_nores":
        epi = "store_f32"
        Out, Outr = torch.empty(B, T, pad8(N), device=DEV), torch.empty(B, T, pad8(N))
        Out2 = Out2r = None
    elif epi == "store_f32":
        Res, Resd = _act(gen, dtype, B, T, N)
        Out, Outr = torch.empty(B, T, pad8(N), device=DEV), torch.empty(B, T, pad8(N))
        Out2 = Out2r = None
    elif epi == "glu":
        Out, Outr = ops.empty(B, T, pad8(N // 2), device=DEV), ref.empty(B, T, pad8(N // 2), device="cpu")
        Out2, Out2r = ops.empty(B, T, pad8(N), device=DEV), ref.empty(B, T, pad8(N), device="cpu")
    else:
        Out, Outr = ops.empty(B, T, pad8(N), device=DEV), ref.empty(B, T, pad8(N), device="cpu")
        Out2, Out2r = ops.empty(B, T, pad8(N), device=DEV), ref.empty(B, T, pad8(N), device="cpu")
    stats = torch.zeros(2, N, device=DEV) if epi == "store_f32" else None
    statsr = torch.zeros(2, N) if epi == "store_f32" else None
    ops.tapgemm(Ad, Ca, W.to(DEV), N, taps, 1, bias.to(DEV), Resd, None, Out, epi=epi, Out2=Out2, stats=stats)
    ref.tapgemm(A, Ca, W, N, taps, 1, bias, Res, None, Outr, epi=epi, Out2=Out2r, stats=statsr)
    torch.cuda.synchronize()
    if stats is not None:      # BatchNorm batch statistics fused into the epilogue
        _close(stats, statsr, torch.float32, f"tapgemm fused BN statistics {backend}", atol_scale=50)
    _close(Out, Outr, torch.float32 if epi == "store_f32" else dtype, f"tapgemm {epi} Out {backend}",
           atol_scale=5 if epi == "store_f32" else 1)
    if Out2 is not None:
        _close(Out2, Out2r, dtype, f"tapgemm {epi} Out2 {backend}")
    if backend == "tc" and case[:5] in EPI_PAIR_EXPECT:
        assert L.get_option("last_tapgemm_kernel") == EPI_PAIR_EXPECT[case[:5]], "the product (CTA-pair) kernel did not run"


# Every hand-written form of the tap GEMM on every shape above that it accepts, forced in-process through
# megb200_set_option (the size thresholds would otherwise keep small shapes on the single-CTA kernel).
TAPGEMM_VARIANTS = {
    "pair_halo": dict(tapgemm_pair_min_items=1),                              # tc_tapgemm2h_kernel for k > 1
    "pair_plain": dict(tapgemm_pair_min_items=1, tapgemm_pair_halo=0),        # tc_tapgemm2_kernel for every k
    "pair_plain_ksub1": dict(tapgemm_pair_min_items=1, tapgemm_pair_halo=0, tapgemm_ksub=1),   # one k chunk per stage
    "single": dict(tapgemm_force_1cta=1),                                     # tc_tapgemm_kernel even above the threshold
    "single_halo": dict(tapgemm_force_1cta=1, tapgemm_halo=1),                # tc_tapgemm_halo_kernel
    "single_direct_epi": dict(tapgemm_force_1cta=1, tapgemm_direct_epi=1),    # per-thread global stores
}


@pytest.mark.parametrize("variant", list(TAPGEMM_VARIANTS))
@pytest.mark.parametrize("case", [c for c in TAPGEMM_CASES if c[6] == 1 and c[0] <= 8])
def test_tapgemm_forced_variants(variant, case):
    with L.options(**TAPGEMM_VARIANTS[variant]):
        test_tapgemm(torch.bfloat16, "tc", case)
        if variant.startswith("pair") and case[3] % 32 == 0:
            assert L.get_option("last_tapgemm_kernel") in (1, 2)


@pytest.mark.parametrize("variant", list(TAPGEMM_VARIANTS))
@pytest.mark.parametrize("case", [c for c in EPI_CASES if c[0] <= 8])
def test_tapgemm_fused_epilogues_forced_variants(variant, case):
    with L.options(**TAPGEMM_VARIANTS[variant]):
        test_tapgemm_fused_epilogues(torch.bfloat16, "tc", case)


WGRAD_CASES = [
    # B, T, M, C, taps, dil, G
    (3, 40, 32, 30, 3, 1, 1),
    (4, 40, 30, 20, 1, 1, 4),        # grouped (subject) wgrad
    (3, 200, 64, 32, 3, 2, 1),
    (4, 360, 320, 320, 3, 1, 1),     # real widths: 3 m tiles, c tile 320 = 192 + 128
    (3, 360, 640, 320, 3, 1, 1),
    (2, 360, 1024, 640, 1, 1, 1),    # head: two c tiles
    (5, 130, 270, 208, 1, 1, 3),     # grouped at real widths
    (20, 72, 320, 270, 3, 1, 1),     # more samples than one CTA range; c = 270 (OOB channels)
    # --- config 2's split plans (B = 256): one wave of CTAs, several samples per CTA
    (256, 360, 320, 320, 3, 1, 1),   # 9 tiles x 16 splits
    (256, 360, 640, 320, 3, 1, 1),   # 15 tiles x 9 splits (floor: one wave)
    (256, 360, 1024, 640, 1, 1, 1),  # head: 16 tiles x 9 splits
    (256, 360, 270, 208, 1, 1, 27),  # grouped subject wgrad at the Gwilliams shape
]


@pytest.mark.parametrize("dtype,backend", CONTRACTION_MODES)
@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad(dtype, backend, case):
    B, T, M, C, taps, dil, G = case
    gen = torch.Generator().manual_seed(2)
    ops, ref = _ops(dtype, backend), TorchOps(dtype)
    dU, dUd = _act(gen, dtype, B, T, M)
    X, Xd = _act(gen, dtype, B, T, C)
    group = torch.randint(0, G, (B,), generator=gen, dtype=torch.int32) if G > 1 else None
    order = None
    if group is not None:
        order = torch.empty(B, dtype=torch.int32, device=DEV)
        ops.group_order(group.to(DEV), G, order)
        assert torch.equal(order.cpu().long(), torch.sort(group.long(), stable=True).indices)   # bit-exact
    dW = torch.zeros(G, M, C, taps, device=DEV)
    ops.wgrad(dUd, M, Xd, C, taps, dil, None if group is None else group.to(DEV), order, dW)
    dWr = torch.zeros(G, M, C, taps)
    ref.wgrad(dU, M, X, C, taps, dil, group, None, dWr)
    torch.cuda.synchronize()
    # fp32 accumulation of the same stored operands: only the summation order differs
    got, want = dW.cpu(), dWr
    err = float((got - want).abs().max()) / (float(want.abs().max()) + 1e-30)
    assert torch.isfinite(got).all() and err < 2e-4, f"wgrad {case} {backend}: {err:.3e}"
    if backend == "tc" and B == 256 and G == 1:
        assert L.get_option("last_wgrad_splits") > 1, "config 2's split plan was not exercised"
    if backend == "tc" and G == 1:
        assert L.get_option("last_wgrad_kernel") == L.get_option("wgrad_pair"), "unexpected weight-gradient kernel"


@pytest.mark.parametrize("case", [c for c in WGRAD_CASES if c[6] == 1 and c[0] <= 20] +
                         [(256, 360, 320, 320, 3, 1, 1), (256, 360, 640, 320, 3, 1, 1), (64, 360, 1024, 640, 1, 1, 1)])
def test_wgrad_cta_pair_kernel_forced(case):
    """The CTA-pair weight-gradient kernel (cta_group::2, M = (tap, input channel) atoms; opt-in: measured slower than the
    single-CTA kernel, see profiles/r2_wgrad_ncu.md) on the ungrouped shapes."""
    with L.options(wgrad_pair=1):
        test_wgrad(torch.bfloat16, "tc", case)
        assert L.get_option("last_wgrad_kernel") == 1


@pytest.mark.parametrize("splits", [1, 2, 5])
def test_wgrad_forced_splits(splits):
    with L.options(wgrad_splits=splits):
        test_wgrad(torch.bfloat16, "tc", (20, 72, 320, 270, 3, 1, 1))
        assert L.get_option("last_wgrad_splits") == splits


@pytest.mark.parametrize("dtype", EW_DTYPES)
def test_colsum_and_sgemm(dtype):
    gen = torch.Generator().manual_seed(3)
    ops, ref = _ops(dtype, "simt"), TorchOps(dtype)
    U, Ud = _act(gen, dtype, 5, 33, 30)
    for group in (None, torch.tensor([2, 0, 2, 1, 0], dtype=torch.int32)):
        G = 1 if group is None else 3
        out = torch.zeros(G, 30, device=DEV)
        ops.colsum(Ud, 30, None if group is None else group.to(DEV), out)
        outr = torch.zeros(G, 30)
        ref.colsum(U, 30, group, outr)
        _close(out, outr, torch.float32, "colsum", atol_scale=5)
    if dtype == torch.float32:
        for tA, tB, beta, K, batch in ((False, False, 0.0, 130, 3), (True, False, 0.0, 130, 3), (False, True, 1.0, 130, 3),
                                       (True, True, 0.5, 130, 3), (True, False, 0.0, 3000, 1), (False, True, 1.0, 2048, 2)):
            M, N = 270, 45          # the long-K cases take the split-K path
            A = _rand(gen, batch, *((K, M) if tA else (M, K)))
            Bm = _rand(gen, *((N, K) if tB else (K, N)))           # 2-D: broadcast over the batch
            C0 = _rand(gen, batch, M, N)
            Cd = C0.to(DEV)
            ops.matmul(A.to(DEV), Bm.to(DEV), Cd, transA=tA, transB=tB, alpha=0.7, beta=beta)
            Cr = C0.clone()
            ref.matmul(A, Bm, Cr, transA=tA, transB=tB, alpha=0.7, beta=beta)
            _close(Cd, Cr, torch.float32, f"sgemm tA={tA} tB={tB}", atol_scale=5)


def test_sa_softmax_fwd_bwd():
    gen = torch.Generator().manual_seed(4)
    ops, ref = _ops(torch.float32, "simt"), TorchOps()
    a = _rand(gen, 30, 20, scale=9.0)
    mask = (torch.rand(20, generator=gen) > 0.2).float()
    for m in (mask, None):
        W, Wm = torch.empty(30, 20, device=DEV), torch.empty(30, 20, device=DEV)
        ops.sa_softmax(a.to(DEV), None if m is None else m.to(DEV), W, Wm)
        Wr, Wmr = torch.empty(30, 20), torch.empty(30, 20)
        ref.sa_softmax(a, m, Wr, Wmr)
        _close(W, Wr, torch.float32, "sa_softmax W")
        _close(Wm, Wmr, torch.float32, "sa_softmax Wm")
        g = _rand(gen, 30, 20)
        da, dar = torch.empty(30, 20, device=DEV), torch.empty(30, 20)
        ops.sa_softmax_bwd(W, g.to(DEV), None if m is None else m.to(DEV), da)
        ref.sa_softmax_bwd(Wr, g, m, dar)
        _close(da, dar, torch.float32, "sa_softmax_bwd")


# ------------------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("dtype", EW_DTYPES)
@pytest.mark.parametrize("u_f32", [True, False])
@pytest.mark.parametrize("shape", [(3, 40, 30), (2, 360, 320), (70, 37, 320)])
def test_batchnorm_gelu_chain(dtype, u_f32, shape):
    B, T, C = shape
    gen = torch.Generator().manual_seed(5)
    ops, ref = _ops(dtype, "simt"), TorchOps(dtype)
    U, Ud = _act(gen, torch.float32 if u_f32 else dtype, B, T, C)      # the pre-BN tensor may stay in f32
    n = B * T
    gamma, beta = torch.rand(C, generator=gen) + 0.5, _rand(gen, C, scale=0.2)
    for training in (1, 0):
        rm, rv = _rand(gen, C, scale=0.1), torch.rand(C, generator=gen) + 0.5
        rmd, rvd = rm.to(DEV), rv.to(DEV)
        sums, sumsr = torch.zeros(2, C, device=DEV), torch.zeros(2, C)
        ops.bn_stats(Ud, C, sums)
        ref.bn_stats(U, C, sumsr)
        _close(sums, sumsr, torch.float32, "bn_stats", atol_scale=5)
        st = {k: torch.empty(C, device=DEV) for k in ("scale", "shift", "mean", "rstd")}
        str_ = {k: torch.empty(C) for k in st}
        ops.bn_finalize(sums, gamma.to(DEV), beta.to(DEV), rmd, rvd, n, 1e-5, 0.1, training, st["scale"], st["shift"],
                        st["mean"], st["rstd"])
        ref.bn_finalize(sumsr, gamma, beta, rm, rv, n, 1e-5, 0.1, training, str_["scale"], str_["shift"], str_["mean"],
                        str_["rstd"])
        for k in st:
            _close(st[k], str_[k], torch.float32, f"bn_finalize {k}", atol_scale=20)
        _close(rmd, rm, torch.float32, "running_mean", atol_scale=20)
        _close(rvd, rv, torch.float32, "running_var", atol_scale=20)
        H = ops.empty(B, T, pad8(C), device=DEV)
        Hr = ref.empty(B, T, pad8(C), device="cpu")
        # feed both sides the same statistics so that only the kernel under test differs
        stc = {k: v.cpu() for k, v in st.items()}
        ops.bn_gelu(Ud, st["scale"], st["shift"], H, C)
        ref.bn_gelu(U, stc["scale"], stc["shift"], Hr, C)
        _close(H, Hr, dtype, "bn_gelu")
        if training:
            dH, dHd = _act(gen, dtype, B, T, C)
            dY, dYr = ops.empty(B, T, pad8(C), device=DEV), ref.empty(B, T, pad8(C), device="cpu")
            s2, s2r = torch.zeros(2, C, device=DEV), torch.zeros(2, C)
            ops.bn_gelu_bwd_reduce(dHd, Ud, st["scale"], st["shift"], st["mean"], st["rstd"], dY, s2, C)
            ref.bn_gelu_bwd_reduce(dH, U, stc["scale"], stc["shift"], stc["mean"], stc["rstd"], dYr, s2r, C)
            _close(dY, dYr, dtype, "bn_gelu_bwd_reduce dY")
            _close(s2, s2r, torch.float32, "bn_gelu_bwd_reduce sums", atol_scale=500 if dtype == torch.bfloat16 else 20)
            dU, dUr = ops.empty(B, T, pad8(C), device=DEV), ref.empty(B, T, pad8(C), device="cpu")
            db = torch.zeros(C, device=DEV)
            ops.bn_bwd_apply(dY, Ud, gamma.to(DEV), st["mean"], st["rstd"], s2, n, dU, C, db)
            ref.bn_bwd_apply(dY.cpu(), U, gamma, stc["mean"], stc["rstd"], s2.cpu(), n, dUr, C)
            _close(dU, dUr, dtype, "bn_bwd_apply")
            # fused conv-bias gradient = column sums of the dU just written (analytically ~0: compare absolutely)
            want = dU.float().cpu()[:, :, :C].sum((0, 1))
            assert float((db.cpu() - want).abs().max()) <= 1e-3 * float(dU.float().abs().max()) * (n ** 0.5)
            # --- fused forms used by the engine: one-launch train-mode BN+GELU, reduce without dY, apply with recompute
            rm2, rv2 = _rand(gen, C, scale=0.1), torch.rand(C, generator=gen) + 0.5
            rm2d, rv2d = rm2.to(DEV), rv2.to(DEV)
            st2 = {k: torch.full((C,), float("nan"), device=DEV) for k in st}
            H2 = ops.empty(B, T, pad8(C), device=DEV)
            ops.bn_gelu_train(Ud, sums, gamma.to(DEV), beta.to(DEV), rm2d, rv2d, n, 1e-5, 0.1, st2["scale"], st2["shift"],
                              st2["mean"], st2["rstd"], H2, C)
            str2 = {k: torch.empty(C) for k in st}
            ref.bn_finalize(sums.cpu(), gamma, beta, rm2, rv2, n, 1e-5, 0.1, 1, str2["scale"], str2["shift"], str2["mean"],
                            str2["rstd"])
            for k in st:
                _close(st2[k], str2[k], torch.float32, f"bn_gelu_train {k}", atol_scale=20)
            _close(rm2d, rm2, torch.float32, "bn_gelu_train running_mean", atol_scale=20)
            _close(rv2d, rv2, torch.float32, "bn_gelu_train running_var", atol_scale=20)
            _close(H2, Hr, dtype, "bn_gelu_train H")
            s3 = torch.zeros(2, C, device=DEV)
            ops.bn_gelu_bwd_reduce(dHd, Ud, st["scale"], st["shift"], st["mean"], st["rstd"], None, s3, C)
            # same arithmetic either way; the per-channel atomics commute only up to f32 rounding
            _close(s3, s2.cpu(), torch.float32, "bn_gelu_bwd_reduce sums without dY", atol_scale=20)
            dU2, dU2r = ops.empty(B, T, pad8(C), device=DEV), ref.empty(B, T, pad8(C), device="cpu")
            db2 = torch.zeros(C, device=DEV)
            ops.bn_gelu_bwd_apply(dHd, Ud, st["scale"], st["shift"], gamma.to(DEV), st["mean"], st["rstd"], s2, n, dU2, C, db2)
            ref.bn_gelu_bwd_apply(dH, U, stc["scale"], stc["shift"], gamma, stc["mean"], stc["rstd"], s2.cpu(), n, dU2r, C)
            _close(dU2, dU2r, dtype, "bn_gelu_bwd_apply")
            want2 = dU2.float().cpu()[:, :, :C].sum((0, 1))
            assert float((db2.cpu() - want2).abs().max()) <= 1e-3 * float(dU2.float().abs().max()) * (n ** 0.5)


@pytest.mark.parametrize("dtype", EW_DTYPES)
@pytest.mark.parametrize("C", [30, 320])
def test_glu_gelu_timemean(dtype, C):
    gen = torch.Generator().manual_seed(6)
    ops, ref = _ops(dtype, "simt"), TorchOps(dtype)
    B, T = 3, 50
    V, Vd = _act(gen, dtype, B, T, 2 * C)
    out, outr = ops.empty(B, T, pad8(C), device=DEV), ref.empty(B, T, pad8(C), device="cpu")
    ops.glu(Vd, out, C)
    ref.glu(V, outr, C)
    _close(out, outr, dtype, "glu")
    d, dd = _act(gen, dtype, B, T, C)
    dV, dVr = ops.empty(*V.shape, device=DEV), ref.empty(*V.shape, device="cpu")
    fused = C % 8 == 0
    db, dbr = (torch.zeros(2 * C, device=DEV), torch.zeros(2 * C)) if fused else (None, None)
    ops.glu_bwd(dd, Vd, dV, C, db)
    ref.glu_bwd(d, V, dVr, C, dbr)
    _close(dV, dVr, dtype, "glu_bwd")
    if fused:
        _close(db, dV.float().cpu()[:, :, :2 * C].sum((0, 1)), torch.float32, "glu_bwd dbias", atol_scale=20)
    P, Pd = _act(gen, dtype, B, T, pad8(C))
    Q, Qr = ops.empty(*P.shape, device=DEV), ref.empty(*P.shape, device="cpu")
    ops.gelu(Pd, Q)
    ref.gelu(P, Qr)
    _close(Q, Qr, dtype, "gelu")
    g, gd = _act(gen, dtype, B, T, pad8(C))
    dP, dPr = ops.empty(*P.shape, device=DEV), ref.empty(*P.shape, device="cpu")
    db, dbr = torch.zeros(C, device=DEV), torch.zeros(C)
    ops.gelu_bwd(gd, Pd, dP, C, db)
    ref.gelu_bwd(g, P, dPr, C, dbr)
    _close(dP, dPr, dtype, "gelu_bwd")
    # the fused column sums are taken before the rounding of dP: compare with one rounding of slack in bf16
    _close(db, dP.float().cpu()[:, :, :C].sum((0, 1)), torch.float32, "gelu_bwd dbias",
           atol_scale=20 if dtype == torch.float32 else 400)
    Fs = pad8(C)
    Zm, Zmr = ops.empty(B, 1, Fs, device=DEV), ref.empty(B, 1, Fs, device="cpu")
    ops.time_mean(Pd, Zm, C)
    ref.time_mean(P, Zmr, C)
    _close(Zm, Zmr, dtype, "time_mean")
    dZ, dZr = ops.empty(B, T, Fs, device=DEV), ref.empty(B, T, Fs, device="cpu")
    ops.time_mean_bwd(Zm, dZ, C)
    ref.time_mean_bwd(Zm.cpu(), dZr, C)
    _close(dZ, dZr, dtype, "time_mean_bwd")


@pytest.mark.parametrize("dtype", EW_DTYPES)
@pytest.mark.parametrize("shape", [(3, 100, 30), (2, 49, 320), (5, 3, 32), (4, 200, 320)])
def test_avgpool_fwd_bwd(dtype, shape):
    """AvgPool1d(3, 2) over time (BrainEncoderSeq2Static, models.py:483-485) and its backward; even and odd lengths (the
    last position of an even-length sequence belongs to no window and gets zero gradient)."""
    B, T, C = shape
    gen = torch.Generator().manual_seed(14)
    ops, ref = _ops(dtype, "simt"), TorchOps(dtype)
    X, Xd = _act(gen, dtype, B, T, C)
    To = (T - 3) // 2 + 1
    out, outr = ops.empty(B, To, pad8(C), device=DEV), ref.empty(B, To, pad8(C), device="cpu")
    ops.avgpool(Xd, out, C)
    ref.avgpool(X, outr, C)
    _close(out, outr, dtype, "avgpool")
    g, gd = _act(gen, dtype, B, To, C)
    dX, dXr = ops.empty(B, T, pad8(C), device=DEV), ref.empty(B, T, pad8(C), device="cpu")
    ops.avgpool_bwd(gd, dX, C)
    ref.avgpool_bwd(g, dXr, C)
    _close(dX, dXr, dtype, "avgpool_bwd")


WIDE_TAP_CASES = [
    # B, T, Ca, N, taps, dil: kernel sizes 5 / 11 (BrainEncoderSeq2Static) and the paper's dilations (some beyond T)
    (2, 200, 320, 320, 11, 1), (3, 100, 32, 64, 5, 1), (2, 360, 320, 320, 3, 32), (2, 360, 320, 320, 3, 2048),
    (80, 200, 320, 320, 11, 1),          # above the CTA-pair threshold: k = 11 does not fit the halo form -> plain pair
    (64, 360, 320, 320, 3, 128),         # halo rows would exceed one box -> plain pair
]


@pytest.mark.parametrize("dtype,backend", CONTRACTION_MODES)
@pytest.mark.parametrize("case", WIDE_TAP_CASES)
def test_tapgemm_and_wgrad_wide_kernels_and_dilations(dtype, backend, case):
    B, T, Ca, N, taps, dil = case
    test_tapgemm(dtype, backend, (B, T, Ca, N, taps, dil, 1, True, True if Ca == N else False))
    if B <= 8:
        test_wgrad(dtype, backend, (B, T, N, Ca, taps, dil, 1))


# ------------------------------------------------------------------------------------------------ CLIP
CLIP_CASES = [
    # By (speech rows, all ranks), Bz (local brains), off, F, T
    (6, 6, 0, 64, 40),
    (5, 5, 0, 48, 1),          # static targets, ragged batch
    (24, 8, 8, 64, 24),        # one-sided data-parallel shard: rank 1 of 3
    (16, 16, 0, 1024, 360),    # the real contraction length D = 368 640
    (300, 150, 150, 512, 1),   # GOD-like: several tiles in both directions
    (256, 256, 0, 1024, 72),   # config 2's batch (single rank): small-K dZ form at By = 256
    (2048, 256, 512, 1024, 36),  # 8-rank shape, rank 2: By > 512 -> tc_clip_dz_kernel (large-K form), 16 i-tiles
    (2048, 2048, 0, 512, 1),   # GOD-like global batch on one rank
]
CLIP_DZ_EXPECT = {(256, 256): 0, (2048, 256): 1, (2048, 2048): 1}


@pytest.mark.parametrize("dtype,backend", CONTRACTION_MODES)
@pytest.mark.parametrize("case", CLIP_CASES)
def test_clip_pipeline(dtype, backend, case):
    By, Bz, off, Fo, T = case
    gen = torch.Generator().manual_seed(7)
    ops, ref = _ops(dtype, backend), TorchOps(dtype)
    D = Fo * T
    Y = _rand(gen, By, Fo, T) if T > 1 else _rand(gen, By, Fo)
    Z = (_rand(gen, Bz, D) * 0.3).to(dtype)
    temp = torch.tensor([5.1])
    # -- prep + norms
    Yp, Ypr = ops.empty(By, D, device=DEV), ref.empty(By, D, device="cpu")
    ysq, ysqr = torch.zeros(By, device=DEV), torch.zeros(By)
    ops.clip_prep_y(Y.to(DEV), Yp, ysq)
    ref.clip_prep_y(Y, Ypr, ysqr)
    _close(Yp, Ypr, dtype, "clip_prep_y")
    _close(ysq, ysqr, torch.float32, "ysq", atol_scale=5)
    zsq, zsqr = torch.zeros(Bz, device=DEV), torch.zeros(Bz)
    Zd = Z.to(DEV)
    ops.rowsumsq(Zd, zsq)
    ref.rowsumsq(Z, zsqr)
    _close(zsq, zsqr, torch.float32, "zsq", atol_scale=5)
    # -- logits contraction (same stored operands on both sides: fp32 accumulation order only)
    raw, rawr = torch.zeros(By, Bz, device=DEV), torch.zeros(By, Bz)
    ops.clip_logits(Yp, Zd, raw)
    ref.clip_logits(Yp.cpu(), Z, rawr)
    _close(raw, rawr, torch.float32, f"clip_logits {backend}", atol_scale=20)
    # -- softmax / CE stage on identical inputs
    rawc, ysqc, zsqc = raw.cpu(), ysq.cpu(), zsq.cpu()
    Lg, Lgr = torch.empty(By, Bz, device=DEV), torch.empty(By, Bz)
    st = {k: torch.empty(n, device=DEV) for k, n in (("rm", By), ("rs", By), ("cl", Bz))}
    str_ = {k: torch.empty(v.numel()) for k, v in st.items()}
    ops.clip_ce_partial(raw, ysq, zsq, temp.to(DEV), Lg, st["rm"], st["rs"], st["cl"])
    ref.clip_ce_partial(rawc, ysqc, zsqc, temp, Lgr, str_["rm"], str_["rs"], str_["cl"])
    _close(Lg, Lgr, torch.float32, "logits", atol_scale=5)
    _close(st["rm"] + torch.log(st["rs"]), str_["rm"] + torch.log(str_["rs"]), torch.float32, "row lse", atol_scale=5)
    _close(st["cl"], str_["cl"], torch.float32, "col lse", atol_scale=5)
    # complete the row LSE as a single rank would see it (other ranks' brains contribute nothing here)
    row_lse = torch.empty(By, device=DEV)
    ops.clip_lse_merge(torch.stack([st["rm"], st["rs"]]).view(1, 2, By).contiguous(), row_lse)
    _close(row_lse, st["rm"] + torch.log(st["rs"]), torch.float32, "clip_lse_merge (single rank)", atol_scale=5)
    wscale = 0.5 / By
    out2, out2r = torch.zeros(2, device=DEV), torch.zeros(2)
    coef, coefr = ops.empty(Bz, pad8(By), device=DEV), ref.empty(Bz, pad8(By), device="cpu")
    cvec, cvecr = torch.zeros(Bz, device=DEV), torch.zeros(Bz)
    ops.clip_ce_finish(Lg, row_lse, st["cl"], ysq, zsq, temp.to(DEV), off, wscale, out2, coef, cvec)
    ref.clip_ce_finish(Lg.cpu(), row_lse.cpu(), st["cl"].cpu(), ysqc, zsqc, temp, off, wscale, out2r, coefr, cvecr)
    _close(out2, out2r, torch.float32, "loss/dtemp", atol_scale=50)
    _close(coef[:, :By], coefr[:, :By], dtype, "coef")
    _close(cvec, cvecr, torch.float32, "cvec", atol_scale=50)
    # -- dZ contraction
    gs = torch.tensor([0.7])
    dZ, dZr = ops.empty(Bz, D, device=DEV), ref.empty(Bz, D, device="cpu")
    ops.clip_dz(coef, Yp, Zd, cvec, gs.to(DEV), dZ)
    ref.clip_dz(coef.cpu(), Yp.cpu(), Z, cvec.cpu(), gs, dZr)
    torch.cuda.synchronize()
    _close(dZ, dZr, dtype, f"clip_dz {backend}")
    if backend == "tc" and (By, Bz) in CLIP_DZ_EXPECT and not L.get_option("clip_dz_large"):
        assert L.get_option("last_clip_dz_kernel") == CLIP_DZ_EXPECT[(By, Bz)]


FUSED_CLIP_CASES = [
    # By (speech rows, all ranks), Bz (local brains), off, D
    (6, 6, 0, 48),              # one ragged tile, one (partial) k chunk
    (300, 150, 150, 512),       # several tiles both ways, data-parallel shard
    (2048, 256, 512, 512),      # 8-rank GOD shard
    (2048, 2048, 0, 512),       # GOD global batch on one rank
    (1000, 1000, 0, 1024),      # ragged last tiles, 16 k chunks
]


@pytest.mark.parametrize("case", FUSED_CLIP_CASES)
def test_clip_fused_never_materialised(case):
    """The large-batch form of the loss (logits tiles live in TMEM only; backward recomputes them) against the op
    restatement of the materialised pipeline on the same stored operands: log-sum-exps, loss, d/dtemp, the coefficient
    matrix (in two column chunks), cvec and dZ."""
    By, Bz, off, D = case
    dtype = torch.bfloat16
    gen = torch.Generator().manual_seed(13)
    ops, ref = _ops(dtype, "tc"), TorchOps(dtype)
    Yp = _rand(gen, By, D).to(dtype)
    Z = (_rand(gen, Bz, D) * 0.3).to(dtype)
    Z[: min(Bz, By - off)] += 0.05 * Yp[off:off + Bz].float().to(dtype)[: min(Bz, By - off)]      # positives stand out a little
    temp = torch.tensor([3.0])
    ysq, zsq = (Yp.float() ** 2).sum(1), (Z.float() ** 2).sum(1)
    # ---- reference: the materialised pipeline
    raw = torch.zeros(By, Bz)
    ref.clip_logits(Yp, Z, raw)
    Lg, rm, rs, cl = torch.empty(By, Bz), torch.empty(By), torch.empty(By), torch.empty(Bz)
    ref.clip_ce_partial(raw, ysq, zsq, temp, Lg, rm, rs, cl)
    row_lse_r = rm + torch.log(rs)
    wscale = 0.5 / By
    out2r, coefr, cvecr = torch.zeros(2), ref.empty(Bz, pad8(By), device="cpu"), torch.zeros(Bz)
    ref.clip_ce_finish(Lg, row_lse_r, cl, ysq, zsq, temp, off, wscale, out2r, coefr, cvecr)
    # ---- fused
    Ypd, Zd, ysqd, zsqd, td = Yp.to(DEV), Z.to(DEV), ysq.to(DEV), zsq.to(DEV), temp.to(DEV)
    row_part = torch.full(((Bz + 127) // 128, 2, By), float("nan"), device=DEV)
    col_part = torch.full(((By + 127) // 128, 2, Bz), float("nan"), device=DEV)
    diag = torch.zeros(Bz, device=DEV)
    ops.clip_fused_stats(Ypd, Zd, ysqd, zsqd, td, off, row_part, col_part, diag)
    row_lse, col_lse = torch.empty(By, device=DEV), torch.empty(Bz, device=DEV)
    ops.clip_lse_merge(row_part, row_lse)
    ops.clip_lse_merge(col_part, col_lse)
    _close(row_lse, row_lse_r, torch.float32, "fused row lse", atol_scale=20)
    _close(col_lse, cl, torch.float32, "fused col lse", atol_scale=20)
    j = torch.arange(Bz)
    _close(diag, Lg[off + j, j], torch.float32, "fused positives", atol_scale=20)
    out2 = torch.zeros(2, device=DEV)
    ops.clip_fused_loss(row_lse, col_lse, diag, off, wscale, out2)
    cvec = torch.zeros(Bz, device=DEV)
    half = max(1, (Bz // 2 + 127) // 128 * 128)
    coef = ops.empty(Bz, pad8(By), device=DEV)
    for j0 in range(0, Bz, half):
        n = min(half, Bz - j0)
        ops.clip_fused_coef(Ypd, Zd, ysqd, zsqd, td, row_lse, col_lse, off, wscale, j0, coef[j0:j0 + n], cvec, out2)
    torch.cuda.synchronize()
    _close(out2, out2r, torch.float32, "fused loss / dtemp", atol_scale=100)
    _close(coef[:, :By], coefr[:, :By], dtype, "fused coef")
    _close(cvec, cvecr, torch.float32, "fused cvec", atol_scale=100)
    gs = torch.tensor([0.7])
    dZ, dZr = ops.empty(Bz, D, device=DEV), ref.empty(Bz, D, device="cpu")
    ops.clip_dz(coef, Ypd, Zd, cvec, gs.to(DEV), dZ)
    ref.clip_dz(coefr, Yp, Z, cvecr, gs, dZr)
    torch.cuda.synchronize()
    _close(dZ, dZr, dtype, "fused dZ")


@pytest.mark.parametrize("W,By", [(1, 37), (2, 512), (8, 2048)])
def test_clip_lse_merge(W, By):
    """The ranks' partial (max, sum exp) pairs -> complete per-row log-sum-exp (one all-gather + this kernel replace a MAX
    and a SUM all-reduce); a rank whose whole shard is far below the maximum must not produce inf/nan."""
    gen = torch.Generator().manual_seed(12)
    ops, ref = _ops(torch.float32, "simt"), TorchOps()
    parts = torch.empty(W, 2, By)
    parts[:, 0] = _rand(gen, W, By, scale=40.0)
    parts[:, 1] = torch.rand(W, By, generator=gen) * 255 + 1.0
    if W > 1:
        parts[0, 0, :5] = -150.0                     # exp underflows against the other ranks' maxima
    out, outr = torch.empty(By, device=DEV), torch.empty(By)
    ops.clip_lse_merge(parts.to(DEV), out)
    ref.clip_lse_merge(parts, outr)
    _close(out, outr, torch.float32, "clip_lse_merge", atol_scale=5)


@pytest.mark.parametrize("case", [c for c in CLIP_CASES if c[1] > 128 or c[0] > 512])
def test_clip_pipeline_dz_two_row_tiles_per_cta(case):
    """Large-K dZ kernel with two 128-row brain tiles per CTA sharing one staged target tile (what the 8-GPU speech shape
    runs: 256 local brains x 1440 column tiles), forced on the smaller shapes."""
    with L.options(clip_dz_large=1, clip_dz_jt=2):
        test_clip_pipeline(torch.bfloat16, "tc", case)
        assert L.get_option("last_clip_dz_kernel") == 1


@pytest.mark.parametrize("case", [c for c in CLIP_CASES if c[0] <= 512])
def test_clip_pipeline_large_k_dz_forced(case):
    """tc_clip_dz_kernel (the form every >= 4-GPU job runs) on the shapes the small-K form normally takes."""
    with L.options(clip_dz_large=1):
        test_clip_pipeline(torch.bfloat16, "tc", case)
        assert L.get_option("last_clip_dz_kernel") == 1
