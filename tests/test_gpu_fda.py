"""GPU parity of the fused FDA kernel through the C-ABI.
Oracle: fp64 / fp32 PyTorch restatement of Aligner + the confidence product (oracle/torch_oracle.py), itself
pinned bit-for-bit to the reference's Aligner / Network.forward (tests/golden/model_*.npz).

Tolerance.  north_star: "soft correspondences match within 1e-3 relative".  Written down here as
  (a) normwise:   max|got - ref| <= 1e-3 * max|ref|                       (the north_star bar)
  (b) allclose:   |got - ref|    <= 1e-3 * |ref| + 1e-4 * max|ref|         element-wise
and asserted 2x tighter than (a) for every input family, 10x tighter for network-like (post-ReLU) inputs.
The kernel evaluates every product with bf16 hi/lo-split operands (3 tensor-core MMAs, ~2^-17 per product) and
fp32 accumulation; the "peaked" family (logits of magnitude ~300, far beyond what the network produces) is where
that shows most, because a logit error of 2^-17*|q||k| moves near-tied softmax weights.
"""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as T
from dcl_testutil import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

from dcl_net_b200 import _lib as L                                    # noqa: E402
from dcl_net_b200.modules import Aligner, fda_align, fda_attention_map  # noqa: E402


def _check(got, want, what, tol_norm=5e-4, atol_rel=1e-4):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().item()
    err = (got - want).abs()
    assert err.max().item() <= tol_norm * scale, \
        f"{what}: normwise error {err.max().item() / scale:.3e} (limit {tol_norm:.1e}; north_star 1e-3)"
    excess = (err - 1e-3 * want.abs() - atol_rel * scale).max().item()
    assert excess <= 0, f"{what}: allclose(rtol=1e-3, atol={atol_rel:.0e}*scale) violated by {excess:.3e}"


@pytest.mark.parametrize("N,K", [(64, 64), (256, 64), (64, 128), (32, 16), (128, 128)])
def test_umma_probe(cuda_dev, N, K):
    """Pins the UMMA descriptor / TMEM conventions the FDA kernel relies on: one CTA, D = A B^T."""
    g = torch.Generator().manual_seed(N * 1000 + K)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
    D = torch.full((128, N), float("nan"), device=cuda_dev)
    a, b = A.to(cuda_dev), B.to(cuda_dev)
    L.check(L.load().dcl_debug_umma_gemm(N, K, L.ptr(a), L.ptr(b), L.ptr(D), 0, L.stream_ptr()), "umma probe")
    torch.cuda.synchronize()
    want = A.double() @ B.double().T
    assert rel_err(D, want) < 3e-5, f"UMMA probe rel err {rel_err(D, want):.3e}"


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,K", [(64, 64), (256, 64), (128, 128), (32, 16)])
def test_umma_pair_probe(cuda_dev, N, K, mode):
    """Pins the CTA-pair (cta_group::2, M=256) conventions: B split by rows over the pair, D rows per CTA."""
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A, B = torch.randn(256, K, generator=g), torch.randn(N, K, generator=g)
    D = torch.full((256, N), float("nan"), device=cuda_dev)
    a, b = A.to(cuda_dev), B.to(cuda_dev)
    L.check(L.load().dcl_debug_umma_pair_gemm(N, K, L.ptr(a), L.ptr(b), L.ptr(D), mode, L.stream_ptr()), "pair probe")
    torch.cuda.synchronize()
    want = A.double() @ B.double().T
    assert rel_err(D, want) < 3e-5, f"UMMA pair probe rel err {rel_err(D, want):.3e}"


def _inputs(seed, b, c, n, m, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "relu":  # post-ReLU features as in the network (BN eval ~ identity)
        ri1, ri2 = torch.randn(b, c, n, generator=g).relu(), torch.randn(b, c, m, generator=g).relu()
    elif kind == "peaked":  # large logits => near one-hot softmax, exercises the rescale path
        ri1, ri2 = 3.0 * torch.randn(b, c, n, generator=g), 3.0 * torch.randn(b, c, m, generator=g)
    else:  # increasing key norms => the running max keeps growing
        ri1 = torch.randn(b, c, n, generator=g).relu()
        ri2 = torch.randn(b, c, m, generator=g).relu() * torch.linspace(0.2, 2.5, m)[None, None, :]
    re2 = torch.randn(b, 256, m, generator=g)
    return ri1, ri2, re2


@pytest.mark.parametrize("kind", ["relu", "peaked", "growing"])
@pytest.mark.parametrize("b,c,n,m", [(1, 64, 128, 64), (2, 64, 128, 128), (2, 64, 256, 192), (3, 64, 1024, 1024),
                                     (2, 128, 128, 64), (2, 128, 1024, 1024), (1, 64, 128, 2048)])
def test_fda_align(cuda_dev, kind, b, c, n, m):
    ri1, ri2, re2 = _inputs(7 + n + m, b, c, n, m, kind)
    re_e, ri_e, lse = fda_align(ri1.to(cuda_dev), ri2.to(cuda_dev), re2.to(cuda_dev), return_lse=True)
    torch.cuda.synchronize()
    want_re, want_ri, a = T.fda_direction(ri1.double(), ri2.double(), re2.double())
    tol = dict(tol_norm=5e-4, atol_rel=5e-4) if kind == "peaked" else dict(tol_norm=1e-4, atol_rel=1e-4)
    _check(re_e, want_re, f"RE_embed {kind}", **tol)
    _check(ri_e, want_ri, f"RI_embed {kind}", **tol)
    if kind == "relu":
        # north_star's bar taken literally — 1e-3 RELATIVE, element by element — where it is well defined: RI_embed is
        # a convex combination of non-negative (post-ReLU) keys, so no element suffers cancellation
        w = want_ri
        pos = w > 0
        rel = ((ri_e.detach().double().cpu() - w).abs()[pos] / w[pos]).max().item()
        assert rel < 1e-3, f"RI_embed element-wise relative error {rel:.2e}"
        assert (ri_e.detach().cpu()[~pos] == 0).all()
    want_lse = torch.logsumexp(torch.bmm(ri2.double().transpose(1, 2), ri1.double()), dim=1)
    assert (lse.double().cpu() - want_lse).abs().max().item() < 1e-3 * max(1.0, want_lse.abs().max().item())


def test_fda_matches_fp32_reference_restatement(cuda_dev):
    """Against the fp32 restatement run on the GPU (what the reference computes, TF32 off)."""
    ri1, ri2, re2 = (t.to(cuda_dev) for t in _inputs(5, 4, 64, 1024, 1024, "relu"))
    re_e, ri_e = fda_align(ri1, ri2, re2)
    want_re, want_ri, _ = T.fda_direction(ri1, ri2, re2)
    _check(re_e, want_re, "RE_embed vs fp32", tol_norm=1e-4)
    _check(ri_e, want_ri, "RI_embed vs fp32", tol_norm=1e-4)


def test_aligner_module_and_attention_map(cuda_dev):
    ri1, ri2, re2 = _inputs(9, 2, 64, 256, 128, "relu")
    re_e, a = Aligner()(ri1.to(cuda_dev), ri2.to(cuda_dev), re2.to(cuda_dev))
    want_re, want_a = T.aligner(ri1.double(), ri2.double(), re2.double())
    _check(re_e, want_re, "Aligner RE_embed")
    assert a.shape == (2, 128, 256)
    assert (a.double().cpu() - want_a).abs().max().item() < 1e-4 * want_a.max().item()
    assert (a.sum(1) - 1).abs().max().item() < 1e-4  # column-stochastic over the m axis (Modules.py:167)


def test_aligner_golden(cuda_dev):
    """tests/golden/model_aligner.npz was produced by the reference's own Aligner (oracle/make_golden.py)."""
    gold = np.load(f"{GOLDEN}/model_aligner.npz")
    g = torch.Generator().manual_seed(int(gold["seed"]))
    ri1, ri2, re2 = (torch.randn(2, 64, 128, generator=g).relu(), torch.randn(2, 64, 192, generator=g).relu(),
                     torch.randn(2, 256, 192, generator=g))
    re_e, a = Aligner()(ri1.to(cuda_dev), ri2.to(cuda_dev), re2.to(cuda_dev))
    _check(re_e, torch.from_numpy(gold["RE_embed"]), "Aligner vs golden")
    assert np.abs(a.cpu().numpy()[:, ::16, ::16] - gold["A_sample"]).max() < 1e-4 * gold["A_sample"].max()


@pytest.mark.parametrize("b,c,n,m", [(2, 64, 128, 128), (2, 128, 256, 192), (3, 128, 1024, 1024)])
def test_fda_point_major_outputs(cuda_dev, b, c, n, m):
    """The point-major bf16 hi/lo images the fused tail consumes hold the same numbers as the fp32 outputs
    (hi + lo carries 16 mantissa bits), alone or next to them."""
    from dcl_net_b200.modules import fda_align_formats
    from dcl_net_b200.fused_tail import pm_unpack
    ri1, ri2, re2 = (x.to(cuda_dev) for x in _inputs(3 + n, b, c, n, m, "relu"))
    re_cm, ri_cm, re_pm, ri_pm, _ = fda_align_formats(ri1, ri2, re2, re_pm=True, ri_pm=True)
    only = fda_align_formats(ri1, ri2, re2, re_cm=False, ri_cm=False, re_pm=True, ri_pm=True)
    assert only[0] is None and only[1] is None
    assert torch.equal(only[2], re_pm) and torch.equal(only[3], ri_pm)
    for cm, pm, ch in ((re_cm, re_pm, 256), (ri_cm, ri_pm, c)):
        rows = pm_unpack(pm, b * n, ch)                       # (b*n, ch)
        want = cm.transpose(1, 2).reshape(b * n, ch)
        assert (rows - want).abs().max().item() <= 2.0 ** -16 * want.abs().max().item()


def test_fda_two_jobs_in_one_launch(cuda_dev):
    """Both directions of the dual FDA in one launch (grid.z = job) == two separate launches, bit for bit."""
    from dcl_net_b200 import modules as M
    b, c, n = 3, 128, 256
    lib = L.load()
    nbytes = lib.dcl_fda_workspace_bytes(b, c, 256, n, n)
    wss, sep = [], []
    for seed in (5, 6):
        ri1, ri2, re2 = (x.to(cuda_dev) for x in _inputs(seed, b, c, n, n, "relu"))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda_dev)
        L.check(lib.dcl_fda_pack(b, c, 256, n, n, L.ptr(ri1), L.ptr(ri2), L.ptr(re2), L.ptr(ws), ws.numel(),
                                 L.stream_ptr()), "pack")
        wss.append(ws)
        sep.append(M.fda_from_workspace(ws, b, c, n, n, True, True, True, True, True))
    both = M.fda_from_workspaces([(wss[0], True, True, True, True, True), (wss[1], True, True, True, True, True)],
                                 b, c, n, n)
    for one, two in zip(sep, both):
        for x, y in zip(one, two):
            assert torch.equal(x, y)


def test_fda_rejects_bad_shapes(cuda_dev):
    x = torch.zeros(1, 32, 128, device=cuda_dev)
    with pytest.raises(ValueError):
        fda_align(x, x, torch.zeros(1, 256, 128, device=cuda_dev))
    y = torch.zeros(1, 64, 100, device=cuda_dev)
    with pytest.raises(ValueError):
        fda_align(y, y, torch.zeros(1, 256, 100, device=cuda_dev))


def test_fda_linearity_in_values_full_size(cuda_dev):
    """Size-independent property at the BASELINE shape (B=32, N=M=1024): the output is linear in the
    value operand, and a constant value row comes back as that constant (softmax weights sum to 1)."""
    b, c, n, m = 32, 64, 1024, 1024
    g = torch.Generator().manual_seed(77)
    ri1, ri2 = torch.randn(b, c, n, generator=g).relu().to(cuda_dev), torch.randn(b, c, m, generator=g).relu().to(cuda_dev)
    v1, v2 = torch.randn(b, 256, m, generator=g).to(cuda_dev), torch.randn(b, 256, m, generator=g).to(cuda_dev)
    o1, _ = fda_align(ri1, ri2, v1)
    o2, _ = fda_align(ri1, ri2, v2)
    o12, _ = fda_align(ri1, ri2, 2.0 * v1 - 3.0 * v2)
    assert rel_err(o12, 2.0 * o1 - 3.0 * o2) < 1e-4
    ones, _ = fda_align(ri1, ri2, torch.full((b, 256, m), 0.75, device=cuda_dev))
    assert (ones - 0.75).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------------ fp16 P V form
@pytest.mark.parametrize("kind", ["relu", "peaked", "growing"])
@pytest.mark.parametrize("b,c,n,m", [(2, 64, 256, 192), (3, 64, 1024, 1024), (2, 128, 256, 64), (2, 128, 1024, 1024),
                                     (1, 64, 256, 2048)])
def test_fda_pv16_within_tolerance(cuda_dev, kind, b, c, n, m):
    """pv_fmt = 1 (the inference path's form): logits on split operands, P and the values rounded once to fp16, one
    MMA per P V product.  Bar (north_star): soft correspondences within 1e-3; measured ~1e-4 normwise."""
    from dcl_net_b200.modules import fda_align_formats
    ri1, ri2, re2 = _inputs(7 + n + m, b, c, n, m, kind)
    re_e, ri_e, _, _, lse = fda_align_formats(ri1.to(cuda_dev), ri2.to(cuda_dev), re2.to(cuda_dev), return_lse=True,
                                              pv_fmt=1)
    torch.cuda.synchronize()
    want_re, want_ri, _ = T.fda_direction(ri1.double(), ri2.double(), re2.double())
    _check(re_e, want_re, f"RE_embed {kind} pv16", tol_norm=1e-3, atol_rel=1e-3)
    _check(ri_e, want_ri, f"RI_embed {kind} pv16", tol_norm=1e-3, atol_rel=1e-3)
    # ... and for network-like inputs it is much tighter than the bar (the "peaked" family has logits ~300 and
    # near one-hot weights: a single fp16-rounded value dominates a row, 2^-11 of the value range)
    if kind != "peaked":
        scale = want_re.abs().max().item()
        assert (re_e.double().cpu() - want_re).abs().max().item() < 4e-4 * scale
    want_lse = torch.logsumexp(torch.bmm(ri2.double().transpose(1, 2), ri1.double()), dim=1)
    assert (lse.double().cpu() - want_lse).abs().max().item() < 1e-3 * max(1.0, want_lse.abs().max().item())


def test_fda_pv16_against_exactly_rounded_operands(cuda_dev):
    """With values that are exactly representable in fp16 and equal within every softmax row's support, the only
    difference to the split path is the rounding of P: constant value rows come back as the constant (numerator and
    denominator carry the same rounded weights)."""
    from dcl_net_b200.modules import fda_align_formats
    b, c, n, m = 2, 128, 256, 256
    ri1, ri2, _ = (t.to(cuda_dev) for t in _inputs(11, b, c, n, m, "relu"))
    const = torch.full((b, 256, m), 0.75, device=cuda_dev)
    out, _, _, _, _ = fda_align_formats(ri1, ri2, const, pv_fmt=1)
    assert (out - 0.75).abs().max().item() < 2e-5


@pytest.mark.parametrize("b,c,n,m", [(2, 64, 256, 128), (3, 128, 1024, 1024)])
def test_fda_pv16_point_major_outputs(cuda_dev, b, c, n, m):
    """PM16 outputs == fp16 rounding of the fp32 outputs of the same launch."""
    from dcl_net_b200.modules import fda_align_formats
    from dcl_net_b200.fused_tail import pm_unpack
    ri1, ri2, re2 = (x.to(cuda_dev) for x in _inputs(3 + n, b, c, n, m, "relu"))
    re_cm, ri_cm, re_pm, ri_pm, _ = fda_align_formats(ri1, ri2, re2, re_pm=True, ri_pm=True, pv_fmt=1)
    for cm, pm, ch in ((re_cm, re_pm, 256), (ri_cm, ri_pm, c)):
        assert pm.numel() == b * n * ch * 2
        rows = pm_unpack(pm, b * n, ch, L.FMT_F16)
        want = cm.transpose(1, 2).reshape(b * n, ch)
        assert torch.equal(rows, want.to(torch.float16).float())


def test_fda_pv16_needs_cta_pairs(cuda_dev):
    from dcl_net_b200.modules import fda_align_formats
    x = torch.zeros(1, 64, 128, device=cuda_dev)
    with pytest.raises(RuntimeError):
        fda_align_formats(x, x, torch.zeros(1, 256, 128, device=cuda_dev), pv_fmt=1)


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,n,m,kind", [(2, 64, 128, 128, "relu"), (3, 64, 256, 384, "relu"), (2, 128, 384, 128, "relu"),
                                          (2, 128, 256, 256, "sharp"), (1, 64, 1024, 1024, "relu")])
def test_fda_backward_fused_vs_fp64(cuda_dev, b, c, n, m, kind):
    """dcl_fda_bwd (csrc/fda_bwd.cu: A rebuilt on chip from lse, the five gradient products on tcgen05) against
    autograd through the reference's graph (models/Modules.py:166-169: bmm -> softmax -> bmm) evaluated in fp64.
    Bar: 5e-5 of each gradient's max, the figure the unfused backward was held to."""
    g = torch.Generator().manual_seed(100 + n + m)
    scale = 1.0 if kind == "relu" else 3.0        # "sharp": logits up to ~100, near one-hot attention rows
    ri1 = (torch.randn(b, c, n, generator=g).relu() * scale).to(cuda_dev).requires_grad_(True)
    ri2 = (torch.randn(b, c, m, generator=g).relu() * scale).to(cuda_dev).requires_grad_(True)
    re2 = torch.randn(b, 256, m, generator=g).to(cuda_dev).requires_grad_(True)
    ge, gi = torch.randn(b, 256, n, generator=g).to(cuda_dev), torch.randn(b, c, n, generator=g).to(cuda_dev)
    e, i = fda_align(ri1, ri2, re2)
    ((e * ge).sum() + (i * gi).sum()).backward()
    got = [t.grad.clone() for t in (ri1, ri2, re2)]
    x64 = [t.detach().double().requires_grad_(True) for t in (ri1, ri2, re2)]
    a = torch.softmax(torch.bmm(x64[1].transpose(1, 2), x64[0]), dim=1)
    ((torch.bmm(x64[2], a) * ge.double()).sum() + (torch.bmm(x64[1], a) * gi.double()).sum()).backward()
    # "sharp": the logits are recomputed by the same split product as in the forward (|S| 2^-17 ~ 8e-4 absolute at
    # |S| ~ 100), which moves individual attention weights by that relative amount: the forward's 5e-4 bar applies
    tol = 5e-5 if kind == "relu" else 5e-4
    for name, gg, t in zip(("d RI_1", "d RI_2", "d RE_2"), got, x64):
        assert rel_err(gg, t.grad) < tol, f"{name}: {rel_err(gg, t.grad):.2e}"


@pytest.mark.gpu
def test_fda_backward_fused_matches_unfused(cuda_dev):
    """Same gradients from the fused kernel and from the materialised-A backward, one output gradient absent."""
    from dcl_net_b200 import modules
    g = torch.Generator().manual_seed(7)
    ri1 = torch.randn(2, 64, 256, generator=g).relu().to(cuda_dev).requires_grad_(True)
    ri2 = torch.randn(2, 64, 256, generator=g).relu().to(cuda_dev).requires_grad_(True)
    re2 = torch.randn(2, 256, 256, generator=g).to(cuda_dev).requires_grad_(True)
    ge = torch.randn(2, 256, 256, generator=g).to(cuda_dev)
    res = []
    for fused in (True, False):
        modules.USE_FUSED_FDA_BACKWARD = fused
        try:
            e, _ = fda_align(ri1, ri2, re2)
            (e * ge).sum().backward()
        finally:
            modules.USE_FUSED_FDA_BACKWARD = True
        res.append([t.grad.clone() for t in (ri1, ri2, re2)])
        for t in (ri1, ri2, re2):
            t.grad = None
    for a, b_ in zip(*res):
        assert rel_err(a, b_) < 5e-5
