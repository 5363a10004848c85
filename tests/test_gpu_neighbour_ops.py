"""GPU parity of the neighbourhood ops — through the C-ABI, against
(1) the C oracle (oracle/neighbour_oracle.c) and (2) the reference's own kernels (oracle/_ref).
Bar: indices, squared distances and forward values bit-exact; backward (atomic reorder) 1e-6 rel."""
import numpy as np
import pytest
import torch

from oracle import cpu_oracle, ref_kernels
from dcl_testutil import cad_like_cloud, flat_bxyz, np_t, rel_err, uniform_cloud

pytestmark = pytest.mark.gpu

from dcl_net_b200.pointnet_lib import pointnet2_utils as pu      # noqa: E402
from dcl_net_b200.pointnet_sp import pointnet2_utils as pu_sp    # noqa: E402
from dcl_net_b200 import _lib as L                               # noqa: E402

HAVE_REF = ref_kernels.available()


def _eq(a, b, what):
    a, b = np_t(a) if torch.is_tensor(a) else a, np_t(b) if torch.is_tensor(b) else b
    assert a.shape == b.shape, what
    assert np.array_equal(a, b, equal_nan=True), f"{what}: {np.sum(a != b)} of {a.size} differ"


CLOUDS = [("uniform", uniform_cloud), ("cad_dup", cad_like_cloud)]


@pytest.mark.parametrize("kind,gen", CLOUDS)
@pytest.mark.parametrize("b,n,m", [(2, 1000, 128), (3, 1024, 256), (2, 4096, 512), (1, 16384, 256), (2, 37, 20),
                                   (1, 20000, 64), (2, 5, 5), (1, 1, 1), (2, 300, 1)])
def test_fps(cuda_dev, kind, gen, b, n, m):
    xyz = gen(11 + n, b, n)
    got = pu.furthest_point_sample(xyz.to(cuda_dev), m)
    assert got.dtype == torch.int32
    _eq(got, cpu_oracle.furthest_point_sample(xyz.numpy(), m), f"FPS vs C oracle ({kind})")
    if HAVE_REF:
        _eq(got, ref_kernels.furthest_point_sample(xyz.to(cuda_dev), m), f"FPS vs reference kernel ({kind})")


def test_fps_temp_side_effect(cuda_dev):
    """temp is left holding the final min-distances (observable through the C-ABI)."""
    xyz = cad_like_cloud(5, 2, 2048).to(cuda_dev)
    temp = torch.full((2, 2048), 1e10, device=cuda_dev)
    out = torch.empty(2, 64, dtype=torch.int32, device=cuda_dev)
    L.check(L.load().dcl_lib_furthest_point_sampling_kernel_launcher(2, 2048, 64, L.ptr(xyz), L.ptr(temp), L.ptr(out),
                                                                     L.stream_ptr()), "fps")
    _, temp_o = cpu_oracle.furthest_point_sample(np_t(xyz), 64, return_temp=True)
    _eq(temp, temp_o, "FPS temp")


@pytest.mark.parametrize("kind,gen", CLOUDS)
@pytest.mark.parametrize("b,n,m,r,ns", [(2, 4096, 256, 0.05, 32), (2, 1023, 100, 0.2, 16), (1, 16384, 512, 0.05, 32),
                                        (2, 2048, 64, 1e-4, 8), (1, 3000, 130, 0.3, 64), (2, 10, 10, 0.5, 3)])
def test_ball_query(cuda_dev, kind, gen, b, n, m, r, ns):
    xyz = gen(21 + n, b, n)
    if kind == "cad_dup":
        r = r * 0.4
    new_xyz = xyz[:, torch.randperm(n, generator=torch.Generator().manual_seed(1))[:m]].contiguous()
    got = pu.ball_query(r, ns, xyz.to(cuda_dev), new_xyz.to(cuda_dev))
    _eq(got, cpu_oracle.ball_query(r, ns, xyz.numpy(), new_xyz.numpy()), "ball_query vs C oracle")
    if HAVE_REF:
        _eq(got, ref_kernels.ball_query(r, ns, xyz.to(cuda_dev), new_xyz.to(cuda_dev)), "ball_query vs reference")


def test_ball_query_empty_balls(cuda_dev):
    xyz = uniform_cloud(3, 2, 500)
    far = xyz[:, :40] + 10.0
    got = pu.ball_query(0.1, 8, xyz.to(cuda_dev), far.contiguous().to(cuda_dev))
    assert int(got.abs().sum()) == 0  # rows stay zero (pointnet2_utils.py:261)


@pytest.mark.parametrize("kind,gen", CLOUDS)
@pytest.mark.parametrize("b,n,m", [(2, 2048, 512), (1, 16384, 1024), (3, 100, 2), (2, 77, 1), (2, 513, 1500)])
def test_three_nn(cuda_dev, kind, gen, b, n, m):
    unknown, known = gen(31 + n, b, n), gen(32 + m, b, m)
    dist, idx = pu.three_nn(unknown.to(cuda_dev), known.to(cuda_dev))
    d2_o, i_o = cpu_oracle.three_nn(unknown.numpy(), known.numpy())
    _eq(idx, i_o, "three_nn idx vs C oracle")
    _eq(dist, np.sqrt(d2_o), "three_nn dist vs C oracle")
    if HAVE_REF:
        d2_r, i_r = ref_kernels.three_nn(unknown.to(cuda_dev), known.to(cuda_dev))
        _eq(idx, i_r, "three_nn idx vs reference")
        _eq(dist, torch.sqrt(d2_r), "three_nn dist vs reference")


@pytest.mark.parametrize("kind,gen", CLOUDS)
@pytest.mark.parametrize("b,n,m,k", [(2, 1024, 1024, 16), (1, 4096, 1024, 16), (2, 300, 200, 1), (2, 300, 200, 3),
                                     (1, 200, 50, 32), (1, 64, 10, 16), (1, 128, 300, 40), (1, 50, 400, 200)])
def test_knn(cuda_dev, kind, gen, b, n, m, k):
    unknown, known = gen(41 + n, b, n), gen(42 + m, b, m)
    dist, idx = pu.knn(k, unknown.to(cuda_dev), known.to(cuda_dev))
    d2_o, i_o = cpu_oracle.knn(k, unknown.numpy(), known.numpy())
    _eq(idx, i_o, "knn idx vs C oracle")
    _eq(dist, np.sqrt(d2_o), "knn dist vs C oracle")
    if HAVE_REF:
        d2_r, i_r = ref_kernels.knn(k, unknown.to(cuda_dev), known.to(cuda_dev))
        _eq(idx, i_r, "knn idx vs reference")
        _eq(dist, torch.sqrt(d2_r), "knn dist vs reference")


@pytest.mark.parametrize("b,c,n,npoint,ns", [(2, 16, 4096, 128, 32), (2, 128, 16384, 64, 32), (1, 5, 1000, 33, 7),
                                             (2, 3, 100, 10, 1), (1, 2, 60000, 50, 16)])
def test_grouping_fwd_bwd(cuda_dev, b, c, n, npoint, ns):
    g = torch.Generator().manual_seed(51)
    feats = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, npoint, ns), generator=g, dtype=torch.int32)
    f = feats.to(cuda_dev).requires_grad_(True)
    out = pu.grouping_operation(f, idx.to(cuda_dev))
    _eq(out, cpu_oracle.grouping_operation(feats.numpy(), idx.numpy()), "grouping fwd vs C oracle")
    go = torch.randn(out.shape, generator=g)
    out.backward(go.to(cuda_dev))
    want = cpu_oracle.grouping_operation_grad(go.numpy(), idx.numpy(), n)
    assert rel_err(f.grad, torch.from_numpy(want)) < 1e-6
    if HAVE_REF:
        _eq(out, ref_kernels.grouping_operation(feats.to(cuda_dev), idx.to(cuda_dev)), "grouping fwd vs reference")
        assert rel_err(f.grad, ref_kernels.grouping_operation_grad(go.to(cuda_dev), idx.to(cuda_dev), n)) < 1e-6


@pytest.mark.parametrize("b,c,n,npoint", [(2, 64, 16384, 1024), (3, 7, 999, 100), (1, 1, 10, 10)])
def test_gather_fwd_bwd(cuda_dev, b, c, n, npoint):
    g = torch.Generator().manual_seed(61)
    feats = torch.randn(b, c, n, generator=g)
    idx = torch.randint(0, n, (b, npoint), generator=g, dtype=torch.int32)
    f = feats.to(cuda_dev).requires_grad_(True)
    out = pu.gather_operation(f, idx.to(cuda_dev))
    _eq(out, cpu_oracle.gather_operation(feats.numpy(), idx.numpy()), "gather fwd")
    go = torch.randn(out.shape, generator=g)
    out.backward(go.to(cuda_dev))
    assert rel_err(f.grad, torch.from_numpy(cpu_oracle.gather_operation_grad(go.numpy(), idx.numpy(), n))) < 1e-6
    if HAVE_REF:
        _eq(out, ref_kernels.gather_operation(feats.to(cuda_dev), idx.to(cuda_dev)), "gather fwd vs reference")


@pytest.mark.parametrize("b,c,m,n", [(2, 128, 1024, 4096), (1, 32, 1024, 16384), (2, 5, 33, 100), (1, 3, 70000, 64)])
def test_three_interpolate_fwd_bwd(cuda_dev, b, c, m, n):
    g = torch.Generator().manual_seed(71)
    feats = torch.randn(b, c, m, generator=g)
    idx = torch.randint(0, m, (b, n, 3), generator=g, dtype=torch.int32)
    w = torch.rand(b, n, 3, generator=g)
    w = w / w.sum(-1, keepdim=True)
    f = feats.to(cuda_dev).requires_grad_(True)
    out = pu.three_interpolate(f, idx.to(cuda_dev), w.to(cuda_dev))
    _eq(out, cpu_oracle.three_interpolate(feats.numpy(), idx.numpy(), w.numpy()), "three_interpolate fwd vs C oracle")
    go = torch.randn(out.shape, generator=g)
    out.backward(go.to(cuda_dev))
    want = cpu_oracle.three_interpolate_grad(go.numpy(), idx.numpy(), w.numpy(), m)
    assert rel_err(f.grad, torch.from_numpy(want)) < 1e-6
    if HAVE_REF:
        _eq(out, ref_kernels.three_interpolate(feats.to(cuda_dev), idx.to(cuda_dev), w.to(cuda_dev)),
            "three_interpolate fwd vs reference")
        assert rel_err(f.grad, ref_kernels.three_interpolate_grad(go.to(cuda_dev), idx.to(cuda_dev), w.to(cuda_dev), m)) < 1e-6


# ------------------------------------------------------------------ pointnet_sp (flat, batch-id aware)
def _sp_case(seed, b, n_per, m_per, dup=False, empty_batch=False, shuffle=True):
    unknown = flat_bxyz(seed, b, n_per, shuffle=False)
    known = flat_bxyz(seed + 1, b, m_per, shuffle=shuffle)
    if dup:
        known[:, 1:] = (known[:, 1:] * 64).round() / 64
        unknown[:, 1:] = (unknown[:, 1:] * 64).round() / 64
    if empty_batch and b > 1:
        known = known[known[:, 0] != 1].contiguous()
    return unknown, known


@pytest.mark.parametrize("b,n_per,m_per,dup,empty", [(4, 256, 300, False, False), (32, 1024, 800, False, False),
                                                     (3, 100, 2, False, False), (4, 128, 50, True, False),
                                                     (4, 128, 50, False, True), (1, 1000, 1, False, False),
                                                     (2, 64, 5000, True, False)])
def test_sp_three_nn(cuda_dev, b, n_per, m_per, dup, empty):
    unknown, known = _sp_case(81, b, n_per, m_per, dup, empty)
    dist, idx = pu_sp.three_nn(unknown.to(cuda_dev), known.to(cuda_dev))
    d2_o, i_o = cpu_oracle.sp_three_nn(unknown.numpy(), known.numpy())
    _eq(idx, i_o, "sp three_nn idx (segmented) vs C oracle")
    _eq(dist, np.sqrt(d2_o), "sp three_nn dist (segmented) vs C oracle")
    dist_f, idx_f = pu_sp.three_nn_full_scan(unknown.to(cuda_dev), known.to(cuda_dev))
    _eq(idx_f, i_o, "sp three_nn idx (full scan) vs C oracle")
    _eq(dist_f, np.sqrt(d2_o), "sp three_nn dist (full scan) vs C oracle")
    if HAVE_REF:
        d2_r, i_r = ref_kernels.sp_three_nn(unknown.to(cuda_dev), known.to(cuda_dev))
        _eq(idx, i_r, "sp three_nn idx vs reference")
        _eq(dist, torch.sqrt(d2_r), "sp three_nn dist vs reference")


def test_sp_three_nn_non_integer_batch_ids(cuda_dev):
    """Batch ids are compared as floats (interpolate_gpu.cu:35); ids that are not small integers take the
    on-device fallback and must still match."""
    unknown, known = _sp_case(91, 3, 50, 40)
    unknown[:, 0] = unknown[:, 0] * 0.5 - 1.0
    known[:, 0] = known[:, 0] * 0.5 - 1.0
    dist, idx = pu_sp.three_nn(unknown.to(cuda_dev), known.to(cuda_dev))
    d2_o, i_o = cpu_oracle.sp_three_nn(unknown.numpy(), known.numpy())
    _eq(idx, i_o, "sp three_nn idx, fractional ids")
    _eq(dist, np.sqrt(d2_o), "sp three_nn dist, fractional ids")


@pytest.mark.parametrize("c", [32, 64, 128, 256, 7])
def test_sp_three_interpolate_fwd_bwd(cuda_dev, c):
    g = torch.Generator().manual_seed(101)
    m, n = 700, 2048
    feats = torch.randn(m, c, generator=g)
    idx = torch.randint(0, m, (n, 3), generator=g, dtype=torch.int32)
    w = torch.rand(n, 3, generator=g)
    f = feats.to(cuda_dev).requires_grad_(True)
    out = pu_sp.three_interpolate(f, idx.to(cuda_dev), w.to(cuda_dev))
    _eq(out, cpu_oracle.sp_three_interpolate(feats.numpy(), idx.numpy(), w.numpy()), "sp interpolate fwd vs C oracle")
    go = torch.randn(n, c, generator=g)
    out.backward(go.to(cuda_dev))
    want = cpu_oracle.sp_three_interpolate_grad(go.numpy(), idx.numpy(), w.numpy(), m)
    assert rel_err(f.grad, torch.from_numpy(want)) < 1e-6
    if HAVE_REF:
        _eq(out, ref_kernels.sp_three_interpolate(feats.to(cuda_dev), idx.to(cuda_dev), w.to(cuda_dev)),
            "sp interpolate fwd vs reference")
        assert rel_err(f.grad, ref_kernels.sp_three_interpolate_grad(go.to(cuda_dev), idx.to(cuda_dev), w.to(cuda_dev), m)) < 1e-6


@pytest.mark.parametrize("b,n_per,m_per,c", [(4, 256, 300, 32), (8, 1024, 130, 128), (2, 100, 2, 64), (3, 50, 40, 7)])
def test_sp_nn_interpolate_fused_equals_unfused(cuda_dev, b, n_per, m_per, c):
    """Fused search+weights+interpolation vs the three-step chain of models/Modules.py:213-226 (kernel ->
    torch element-wise weights -> kernel).  Neighbours are identical; the weights go through torch's own
    element-wise kernels in the chain and through explicit IEEE ops in the fused kernel, so values agree to
    rounding (<= 1e-6 relative), not necessarily bit for bit."""
    unknown, known = _sp_case(111, b, n_per, m_per)
    feats = torch.randn(known.shape[0], c, generator=torch.Generator().manual_seed(3))
    u, k, f = unknown.to(cuda_dev), known.to(cuda_dev), feats.to(cuda_dev)
    dist, idx = pu_sp.three_nn(u, k)
    recip = 1.0 / (dist + 1e-8)
    w = recip / torch.sum(recip, dim=1, keepdim=True)
    want = pu_sp.three_interpolate(f, idx, w)
    out = torch.full((u.shape[0], c + 8), -1.0, device=cuda_dev)
    pu_sp.nn_interpolate(u, k, f, out, 4 if c % 4 == 0 else 1)
    col0 = 4 if c % 4 == 0 else 1
    assert rel_err(out[:, col0:col0 + c], want) < 1e-6, "fused nn_interpolate"
    assert float(out[:, :col0].min()) == -1.0 and float(out[:, col0 + c:].max()) == -1.0
