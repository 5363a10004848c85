"""GPU parity of the tensor-core pointwise-MLP path (csrc/pm_gemm.cu, fused_tail.py) — SURVEY.md §8 a12 (the MLP
stacks of the FDA section) and a10 (point features written as operand images).
Oracle: fp64 matmul / the fp32 PyTorch layers of the reference (oracle/torch_oracle.py).  The bf16 hi/lo split keeps
~2^-17 per product with fp32 accumulation: asserted at 2e-5 of the output scale per layer."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as T
from dcl_testutil import flat_bxyz, rel_err

pytestmark = pytest.mark.gpu

from dcl_net_b200 import _lib as L                                  # noqa: E402
from dcl_net_b200 import fused_tail as FT                            # noqa: E402
from dcl_net_b200.dcl_net import Network                             # noqa: E402
from dcl_net_b200.pointnet_sp import pointnet2_utils as pu_sp        # noqa: E402


@pytest.mark.parametrize("rows,c", [(128, 32), (256, 480), (1024, 64)])
def test_pm_pack_roundtrip(cuda_dev, rows, c):
    x = torch.randn(rows, c, generator=torch.Generator().manual_seed(rows + c)).to(cuda_dev)
    pm = FT.pm_pack_rows(x)
    back = FT.pm_unpack(pm, rows, c)
    assert (back - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    # channel-major packing gives the same image
    b = rows // 128
    x_cm = x.view(b, 128, c).transpose(1, 2).contiguous()
    assert torch.equal(FT.pm_pack_cm(x_cm), pm)


def _ref_layer(x, w, bias, relu, ps, pt):
    y = x.double() @ w.double().T
    if bias is not None:
        y = y + bias.double()
    if relu:
        y = y.clamp_min(0)
    if ps is not None:
        y = y * ps.double() + pt.double()
    return y


@pytest.mark.parametrize("rows,cin,cout,relu,post,split", [
    (256, 480, 256, True, False, 0), (128, 256, 64, True, False, 0), (384, 256, 128, False, False, 0),
    (256, 512, 512, True, True, 256), (128, 128, 1024, True, True, 0), (256, 256, 128, True, False, 128),
    (128, 32, 64, False, False, 0)])
def test_pm_gemm_layer(cuda_dev, rows, cin, cout, relu, post, split):
    g = torch.Generator().manual_seed(rows + cin + cout)
    x = torch.randn(rows, cin, generator=g).to(cuda_dev)
    w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(cuda_dev)
    bias = torch.randn(cout, generator=g).to(cuda_dev)
    ps = (torch.rand(cout, generator=g) + 0.5).to(cuda_dev) if post else None
    pt = torch.randn(cout, generator=g).to(cuda_dev) if post else None
    lay = FT.Layer(w, bias, relu, ps, pt)
    n_inst = 128
    out_pm = FT.pm_empty(rows, cout, cuda_dev)
    out_cm = torch.full((rows // n_inst, cout, n_inst), float("nan"), device=cuda_dev)
    pool_w = torch.rand(rows, generator=g).to(cuda_dev)
    pool_out = torch.full((rows // 32, cout), float("nan"), device=cuda_dev)
    prob = {"layer": lay, "out_pm": out_pm, "out_cm": out_cm, "rows_per_inst": n_inst, "pool_w": pool_w, "pool_out": pool_out}
    if split:
        prob.update(a0=FT.pm_pack_rows(x[:, :split]), a1=FT.pm_pack_rows(x[:, split:]), c0=split)
    else:
        prob.update(a0=FT.pm_pack_rows(x))
    FT.run_gemm([prob], rows)
    torch.cuda.synchronize()
    want = _ref_layer(x, w, bias, relu, ps, pt)
    scale = want.abs().max().item()
    got_pm = FT.pm_unpack(out_pm, rows, cout)
    assert (got_pm.double() - want).abs().max().item() <= 3e-5 * scale
    got_cm = out_cm.transpose(1, 2).reshape(rows, cout)
    assert (got_cm.double() - want).abs().max().item() <= 2e-5 * scale
    want_pool = (want * pool_w.double()[:, None]).view(rows // 32, 32, cout).sum(1)
    assert (pool_out.double() - want_pool).abs().max().item() <= 2e-5 * want_pool.abs().max().item()
    pooled = torch.empty(rows // n_inst, cout, device=cuda_dev)
    L.check(L.load().dcl_pm_pool_reduce(rows // n_inst, cout, n_inst // 32, L.ptr(pool_out), None, L.ptr(pooled), 0,
                                        L.stream_ptr()), "pool")
    assert rel_err(pooled, want_pool.view(rows // n_inst, n_inst // 32, cout).sum(1)) < 2e-5


@pytest.mark.parametrize("b,n,c", [(2, 128, 64), (3, 256, 128)])
def test_pm_gemm_writes_fda_operand_images(cuda_dev, b, n, c):
    """The disengage GEMMs write the FDA query / key / value operand images themselves; byte for byte they must be
    what dcl_fda_pack makes of the same layers' channel-major fp32 outputs."""
    import ctypes
    g = torch.Generator().manual_seed(b * n + c)
    rows = b * n
    x = torch.randn(rows, 256, generator=g).to(cuda_dev)
    a0 = FT.pm_pack_rows(x)
    lays = {name: FT.Layer((torch.randn(co, 256, generator=g) / 16).to(cuda_dev), torch.randn(co, generator=g).to(cuda_dev),
                           True, None, None) for name, co in (("q", c), ("k", c), ("p", 256))}
    lib = L.load()
    nbytes = lib.dcl_fda_workspace_bytes(b, c, 256, n, n)
    offs = (ctypes.c_size_t * 3)()
    L.check(lib.dcl_fda_workspace_layout(b, c, 256, n, n, ctypes.cast(offs, ctypes.c_void_p)), "layout")
    ws_direct = torch.zeros(nbytes, dtype=torch.uint8, device=cuda_dev)
    ws_packed = torch.zeros(nbytes, dtype=torch.uint8, device=cuda_dev)
    base = ws_direct.data_ptr()
    cm = {name: torch.empty(b, lay.cout, n, device=cuda_dev) for name, lay in lays.items()}
    FT.run_gemm([{"a0": a0, "layer": lays["q"], "out_cm": cm["q"], "rows_per_inst": n,
                  "out_qk": base + offs[0], "qk_tile_rows": 128},
                 {"a0": a0, "layer": lays["k"], "out_cm": cm["k"], "rows_per_inst": n,
                  "out_qk": base + offs[1], "qk_tile_rows": 64, "out_v": base + offs[2], "v_row0": 256, "v_rows": 256 + c}],
                rows)
    FT.run_gemm([{"a0": a0, "layer": lays["p"], "out_cm": cm["p"], "rows_per_inst": n,
                  "out_v": base + offs[2], "v_row0": 0, "v_rows": 256 + c}], rows)
    L.check(lib.dcl_fda_pack(b, c, 256, n, n, L.ptr(cm["q"]), L.ptr(cm["k"]), L.ptr(cm["p"]), L.ptr(ws_packed),
                             ws_packed.numel(), L.stream_ptr()), "pack")
    torch.cuda.synchronize()
    used = offs[2] + b * n * (256 + c) * 4
    assert torch.equal(ws_direct[:used], ws_packed[:used])


def test_pm_gemm_batched_problems(cuda_dev):
    """Several problems (different inputs and weights) in one launch."""
    g = torch.Generator().manual_seed(3)
    rows, cin, cout = 256, 64, 256
    xs = [torch.randn(rows, cin, generator=g).to(cuda_dev) for _ in range(3)]
    ws = [(torch.randn(cout, cin, generator=g) / 8).to(cuda_dev) for _ in range(5)]
    outs = [FT.pm_empty(rows, cout, cuda_dev) for _ in range(5)]
    FT.run_gemm([{"a0": FT.pm_pack_rows(xs[i % 3]), "layer": FT.Layer(ws[i], None, i % 2 == 0), "out_pm": outs[i]}
                 for i in range(5)], rows)
    for i in range(5):
        want = _ref_layer(xs[i % 3], ws[i], None, i % 2 == 0, None, None)
        assert (FT.pm_unpack(outs[i], rows, cout).double() - want).abs().max().item() <= 3e-5 * want.abs().max().item()


@pytest.mark.parametrize("b,n_per,m_per,c,col0,ctot", [(2, 128, 90, 32, 0, 480), (4, 256, 40, 128, 96, 480), (1, 128, 7, 256, 224, 480)])
def test_nn_interpolate_pm_equals_fp32(cuda_dev, b, n_per, m_per, c, col0, ctot):
    unknown = flat_bxyz(5, b, n_per, shuffle=False).to(cuda_dev)
    known = flat_bxyz(6, b, m_per).to(cuda_dev)
    feats = torch.randn(known.shape[0], c, generator=torch.Generator().manual_seed(7)).to(cuda_dev)
    want = pu_sp.nn_interpolate(unknown, known, feats)
    pm = torch.zeros(FT.pm_bytes(b * n_per, ctot), dtype=torch.uint8, device=cuda_dev)
    pu_sp.nn_interpolate_pm(unknown, known, feats, pm, ctot, col0)
    got = FT.pm_unpack(pm, b * n_per, ctot)
    assert (got[:, col0:col0 + c] - want).abs().max().item() <= 2.0 ** -16 * want.abs().max().item()
    assert float(got[:, :col0].abs().sum()) == 0 and float(got[:, col0 + c:].abs().sum()) == 0


class Cfg:
    def __init__(self, n):
        self.n_inp = self.n_tmp = n
        self.unit_voxel_extent = [0.006] * 3


@pytest.mark.parametrize("b,n,c_m", [(2, 128, 64), (4, 1024, 64), (2, 1024, 128)])
def test_fused_tail_equals_unfused_and_oracle(cuda_dev, b, n, c_m):
    torch.manual_seed(11)
    oracle_net = T.TailNetwork(mode="test", c_m=c_m).eval()
    # non-trivial BatchNorm statistics so that folding / the post-ReLU affine are really exercised
    gg = torch.Generator().manual_seed(12)
    for mod in oracle_net.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm3d)):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=gg))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=gg))
            mod.weight.data.copy_(0.5 + torch.rand(mod.num_features, generator=gg))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=gg))
    net = Network(Cfg(n), mode="test", c_m=c_m).eval()
    net.load_state_dict(oracle_net.state_dict(), strict=False)
    net = net.to(cuda_dev)
    g = torch.Generator().manual_seed(b * n)
    f_xc, f_yo = torch.randn(b * n, 480, generator=g).to(cuda_dev), torch.randn(b * n, 480, generator=g).to(cuda_dev)
    with torch.no_grad():
        got = net.forward_from_point_feats(f_xc, f_yo, b)
        assert net._fused_tail is not None, "the tensor-core path did not run"
        net.use_fused_tail = False
        unfused = net.forward_from_point_feats(f_xc, f_yo, b)
        want = oracle_net.to(cuda_dev)(f_xc, f_yo, b, n, n)
    for ref, name in ((unfused, "unfused"), (want, "oracle")):
        assert rel_err(got["F_Xo_p"], ref["F_Xo_p"]) < 1e-3, name
        assert rel_err(got["conf"], ref["conf"]) < 1e-3, name
        ang = T.rotation_angle_deg(got["rot_pred"].cpu(), ref["rot_pred"].cpu()).max().item()
        dt = (got["trans_pred"] - ref["trans_pred"]).abs().max().item()
        assert ang < 0.01 and dt < 1e-5, (name, ang, dt)


def test_nn_interpolate_vox_pm_equals_tensor2points_path(cuda_dev):
    """Voxel centres formed inside the kernels == Ops_tensor2points followed by the float-row kernels, bit for bit."""
    import types
    from dcl_net_b200.modules import Ops_tensor2points
    g = torch.Generator().manual_seed(21)
    b, n_per, c = 3, 128, 64
    unknown = flat_bxyz(22, b, n_per, shuffle=False, scale=0.3).to(cuda_dev)
    ind = torch.cat([torch.randint(0, b, (500, 1), generator=g), torch.randint(0, 16, (500, 3), generator=g)], 1).int()
    ind = torch.unique(ind, dim=0)
    ind = ind[torch.randperm(ind.shape[0], generator=g)].contiguous().to(cuda_dev)
    feats = torch.randn(ind.shape[0], c, generator=g).to(cuda_dev)
    ext, off = np.array([0.024, 0.024, 0.024]), np.array([-0.192, -0.192, -0.192])
    _, centres = Ops_tensor2points(types.SimpleNamespace(features=feats, indices=ind), off, ext)
    pm_a = torch.zeros(FT.pm_bytes(b * n_per, 64), dtype=torch.uint8, device=cuda_dev)
    pm_b = torch.zeros_like(pm_a)
    pu_sp.nn_interpolate_pm(unknown, centres.contiguous(), feats, pm_a, 64, 0)
    e32 = torch.as_tensor(ext, dtype=torch.float32).tolist()
    o32 = torch.as_tensor(off, dtype=torch.float32).tolist()
    pu_sp.nn_interpolate_vox_pm(unknown, ind, e32, o32, feats, pm_b, 64, 0)
    assert torch.equal(pm_a, pm_b)


@pytest.mark.parametrize("slabs", [False, True])
@pytest.mark.parametrize("b,n_per,sizes", [(3, 128, (500, 200, 60, 9)), (32, 256, (30000, 9000, 2500, 700)),
                                          (2, 128, (300, 0, 5, 1)), (1, 128, (4000,)), (40, 128, (60000,))])
def test_nn_interpolate_levels_equals_per_level_calls(cuda_dev, b, n_per, sizes, slabs):
    """The two-launch multi-level path (cluster bucket build + one search/interpolation launch) writes the same
    point-major image, bit for bit, as one nn_interpolate_vox_pm call per level — including an empty level, levels
    with fewer than three voxels per instance and shuffled voxel order; with `slabs` the search walks slabs of equal
    first voxel index instead of scanning the whole instance (40 instances x 64 slabs overflow the bucket table and
    take the full-scan fallback; 4000 voxels in one instance overflow the shared-memory staging)."""
    g = torch.Generator().manual_seed(sum(sizes) + b)
    unknown = flat_bxyz(23, b, n_per, shuffle=False, scale=0.3)
    # every fourth query sits exactly on a corner of the finest voxel grid: equidistant (up to rounding) from the
    # eight surrounding centres, i.e. distance ties across slabs
    corner = torch.randint(0, 65, (unknown.shape[0], 3), generator=g).float() * (0.6 / 64) - 0.3
    unknown[::4, 1:] = corner[::4]
    unknown = unknown.contiguous().to(cuda_dev)
    widths = (32, 64, 128, 256)[:len(sizes)]
    total = sum(widths)
    specs, col = [], 0
    pm_a = torch.zeros(FT.pm_bytes(b * n_per, total), dtype=torch.uint8, device=cuda_dev)
    pm_b = torch.zeros_like(pm_a)
    for li, (m, c) in enumerate(zip(sizes, widths)):
        side = 64 >> li
        ind = torch.cat([torch.randint(0, b, (m, 1), generator=g), torch.randint(0, side, (m, 3), generator=g)], 1).int()
        if m:
            ind = torch.unique(ind, dim=0)
            ind = ind[torch.randperm(ind.shape[0], generator=g)]
        ind = ind.contiguous().to(cuda_dev)
        feats = torch.randn(ind.shape[0], c, generator=g).to(cuda_dev)
        ext = [0.6 / side] * 3
        off = [-0.3] * 3
        specs.append((ind, ext, off, feats, col, side if slabs else 0))
        if ind.shape[0]:
            pu_sp.nn_interpolate_vox_pm(unknown, ind, ext, off, feats, pm_a, total, col)
        col += c
    pu_sp.nn_interpolate_vox_levels_pm(unknown, specs, pm_b, total)
    if 0 in [s[0].shape[0] for s in specs]:
        # an empty level: every query gets (inf, 0) neighbours -> weights NaN -> skip that column range
        rows_a, rows_b = FT.pm_unpack(pm_a, b * n_per, total), FT.pm_unpack(pm_b, b * n_per, total)
        c0 = 0
        for spec, c in zip(specs, widths):
            ind, col0 = spec[0], spec[4]
            if ind.shape[0]:
                # bitwise: instances with no voxel in a level produce NaN weights in both paths
                assert torch.equal(rows_a[:, col0:col0 + c].contiguous().view(torch.int32),
                                   rows_b[:, col0:col0 + c].contiguous().view(torch.int32))
    else:
        assert torch.equal(pm_a, pm_b)


@pytest.mark.parametrize("rows,cin,cout,nprob", [(2048, 96, 512, 5), (4096, 160, 256, 3), (1024, 480, 128, 8), (2048, 64, 64, 2)])
def test_pm_gemm_persistent_many_tiles(cuda_dev, rows, cin, cout, nprob):
    """More work units than CTA pairs and k-block counts that are not multiples of the ring depth: exercises the
    persistent tile loop, the ring wrap-around across tiles, the TMEM accumulator ping-pong and the multicast pairing."""
    g = torch.Generator().manual_seed(rows + cin + cout + nprob)
    xs = [torch.randn(rows, cin, generator=g).to(cuda_dev) for _ in range(2)]
    pms = [FT.pm_pack_rows(x) for x in xs]
    ws = [(torch.randn(cout, cin, generator=g) / cin ** 0.5).to(cuda_dev) for _ in range(nprob)]
    bs = [torch.randn(cout, generator=g).to(cuda_dev) for _ in range(nprob)]
    outs = [FT.pm_empty(rows, cout, cuda_dev) for _ in range(nprob)]
    cms = [torch.empty(rows // 128, cout, 128, device=cuda_dev) for _ in range(nprob)]
    FT.run_gemm([{"a0": pms[i % 2], "layer": FT.Layer(ws[i], bs[i], True), "out_pm": outs[i], "out_cm": cms[i],
                  "rows_per_inst": 128} for i in range(nprob)], rows)
    torch.cuda.synchronize()
    for i in range(nprob):
        want = _ref_layer(xs[i % 2], ws[i], bs[i], True, None, None)
        scale = want.abs().max().item()
        assert (FT.pm_unpack(outs[i], rows, cout).double() - want).abs().max().item() <= 3e-5 * scale, i
        assert (cms[i].transpose(1, 2).reshape(rows, cout).double() - want).abs().max().item() <= 2e-5 * scale, i


# ------------------------------------------------------------------------------------------------ PM16 (fp16) format
def _f16(x):
    return x.to(torch.float16).to(torch.float32)


@pytest.mark.parametrize("rows,c", [(128, 32), (256, 480), (1024, 64)])
def test_pm16_pack_roundtrip(cuda_dev, rows, c):
    x = torch.randn(rows, c, generator=torch.Generator().manual_seed(rows + c)).to(cuda_dev)
    x[0, 0], x[1, 1] = 1e6, -1e6                                    # beyond the fp16 range: clamped, not inf
    pm = FT.pm_pack_rows(x, L.FMT_F16)
    assert pm.numel() == rows * c * 2
    back = FT.pm_unpack(pm, rows, c, L.FMT_F16)
    assert torch.equal(back, _f16(x.clamp(-65504, 65504)))
    b = rows // 128
    x_cm = x.view(b, 128, c).transpose(1, 2).contiguous()
    assert torch.equal(FT.pm_pack_cm(x_cm, L.FMT_F16), pm)


@pytest.mark.parametrize("rows,cin,cout,relu,post,split", [
    (256, 480, 256, True, False, 0), (128, 256, 64, True, False, 0), (384, 256, 128, False, False, 0),
    (256, 512, 512, True, True, 256), (128, 128, 1024, True, True, 0), (256, 256, 128, True, False, 128),
    (128, 32, 64, False, False, 0), (4096, 512, 1024, True, True, 0)])
def test_pm16_gemm_layer(cuda_dev, rows, cin, cout, relu, post, split):
    """PM16 operands: X rounded once to fp16, W as fp16 hi + lo, 2 MMAs.  Against the fp64 product of the SAME rounded
    X the kernel must be fp32-faithful (the weight split keeps ~2^-22): the only precision given up is the rounding
    of the stored activation, which is asserted separately on the written PM16 image (half an fp16 ulp)."""
    g = torch.Generator().manual_seed(rows + cin + cout)
    x = torch.randn(rows, cin, generator=g).to(cuda_dev)
    w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(cuda_dev)
    bias = torch.randn(cout, generator=g).to(cuda_dev)
    ps = (torch.rand(cout, generator=g) + 0.5).to(cuda_dev) if post else None
    pt = torch.randn(cout, generator=g).to(cuda_dev) if post else None
    lay = FT.Layer(w, bias, relu, ps, pt, fmt=L.FMT_F16)
    n_inst = 128
    out_pm = FT.pm_empty(rows, cout, cuda_dev, L.FMT_F16)
    out_cm = torch.full((rows // n_inst, cout, n_inst), float("nan"), device=cuda_dev)
    pool_w = torch.rand(rows, generator=g).to(cuda_dev)
    pool_out = torch.full((rows // 32, cout), float("nan"), device=cuda_dev)
    prob = {"layer": lay, "out_pm": out_pm, "out_cm": out_cm, "rows_per_inst": n_inst, "pool_w": pool_w, "pool_out": pool_out}
    if split:
        prob.update(a0=FT.pm_pack_rows(x[:, :split], L.FMT_F16), a1=FT.pm_pack_rows(x[:, split:], L.FMT_F16), c0=split)
    else:
        prob.update(a0=FT.pm_pack_rows(x, L.FMT_F16))
    FT.run_gemm([prob], rows)
    torch.cuda.synchronize()
    want = _ref_layer(_f16(x), w, bias, relu, ps, pt)
    scale = want.abs().max().item()
    got_cm = out_cm.transpose(1, 2).reshape(rows, cout)
    assert (got_cm.double() - want).abs().max().item() <= 4e-6 * scale
    got_pm = FT.pm_unpack(out_pm, rows, cout, L.FMT_F16)
    assert torch.equal(got_pm, _f16(got_cm)), "the PM16 image is the fp16 rounding of the fp32 result"
    want_pool = (want * pool_w.double()[:, None]).view(rows // 32, 32, cout).sum(1)
    assert (pool_out.double() - want_pool).abs().max().item() <= 4e-6 * want_pool.abs().max().item()
    # and against the unrounded input: the price of the format, well inside the 1e-3 bar of the path
    exact = _ref_layer(x, w, bias, relu, ps, pt)
    assert (got_cm.double() - exact).abs().max().item() <= 3e-4 * exact.abs().max().item()


@pytest.mark.parametrize("b,n,c", [(2, 256, 64), (3, 256, 128)])
def test_pm16_gemm_writes_fda_operand_images(cuda_dev, b, n, c):
    """fp16 path: query / key images stay bf16 hi/lo, the value image is one fp16 image per chunk; byte for byte what
    dcl_fda_pack_fmt(pv_fmt=1) makes of the same layers' fp32 outputs."""
    import ctypes
    g = torch.Generator().manual_seed(b * n + c)
    rows = b * n
    x = torch.randn(rows, 256, generator=g).to(cuda_dev)
    a0 = FT.pm_pack_rows(x, L.FMT_F16)
    lays = {name: FT.Layer((torch.randn(co, 256, generator=g) / 16).to(cuda_dev), torch.randn(co, generator=g).to(cuda_dev),
                           True, None, None, fmt=L.FMT_F16) for name, co in (("q", c), ("k", c), ("p", 256))}
    lib = L.load()
    nbytes = lib.dcl_fda_workspace_bytes(b, c, 256, n, n)
    offs = (ctypes.c_size_t * 3)()
    L.check(lib.dcl_fda_workspace_layout(b, c, 256, n, n, ctypes.cast(offs, ctypes.c_void_p)), "layout")
    ws_direct = torch.zeros(nbytes, dtype=torch.uint8, device=cuda_dev)
    ws_packed = torch.zeros(nbytes, dtype=torch.uint8, device=cuda_dev)
    base = ws_direct.data_ptr()
    cm = {name: torch.empty(b, lay.cout, n, device=cuda_dev) for name, lay in lays.items()}
    FT.run_gemm([{"a0": a0, "layer": lays["q"], "out_cm": cm["q"], "rows_per_inst": n,
                  "out_qk": base + offs[0], "qk_tile_rows": 128},
                 {"a0": a0, "layer": lays["k"], "out_cm": cm["k"], "rows_per_inst": n,
                  "out_qk": base + offs[1], "qk_tile_rows": 64, "out_v": base + offs[2], "v_row0": 256, "v_rows": 256 + c}],
                rows)
    FT.run_gemm([{"a0": a0, "layer": lays["p"], "out_cm": cm["p"], "rows_per_inst": n,
                  "out_v": base + offs[2], "v_row0": 0, "v_rows": 256 + c}], rows)
    L.check(lib.dcl_fda_pack_fmt(b, c, 256, n, n, L.ptr(cm["q"]), L.ptr(cm["k"]), L.ptr(cm["p"]), L.ptr(ws_packed),
                                 ws_packed.numel(), 1, L.stream_ptr()), "pack")
    torch.cuda.synchronize()
    used = offs[2] + b * n * (256 + c) * 2
    assert torch.equal(ws_direct[:used], ws_packed[:used])


def test_pm16_interpolation_image_equals_fp16_of_fp32(cuda_dev):
    """The multi-level 3-NN interpolation writing a PM16 image == fp16 rounding of the fp32 point features (same
    search, same fma order; the bf16 hi/lo image is not the comparand: rounding it again to fp16 double-rounds)."""
    import types
    from dcl_net_b200.modules import Ops_GetPointFeat_spconv
    import bench
    b = 2
    batch = bench.make_host_batch(5, b, pin=False)
    getter = Ops_GetPointFeat_spconv(scale_lists=[2, 4, 6, 8], unit_voxel_extent=np.array([0.006] * 3),
                                     voxel_num_limit=[64, 64, 64])
    ids = torch.arange(b, device=cuda_dev).repeat_interleave(bench.N_PTS)
    lv = lambda side: [types.SimpleNamespace(features=f.to(cuda_dev), indices=i.to(cuda_dev)) for f, i in batch[side]]
    args = (batch["points_inp"].to(cuda_dev), ids, lv("inp"), batch["points_tmp"].to(cuda_dev), ids, lv("tmp"))
    h_a, h_b = getter.forward_pm_pair(*args, fmt=L.FMT_F16)
    rows = b * bench.N_PTS
    with torch.no_grad():
        f_a, f_b = getter(args[0], ids, *args[2]), getter(args[3], ids, *args[5])
    for full, half in ((f_a, h_a), (f_b, h_b)):
        assert half.numel() == rows * 480 * 2
        assert torch.equal(FT.pm_unpack(half, rows, 480, L.FMT_F16), _f16(full))
