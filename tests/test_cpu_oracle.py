"""CPU: the C oracle (oracle/neighbour_oracle.c) against (a) independent numpy statements of each op's
semantics, (b) the documented tie-break rules, (c) fixtures produced by the reference's own kernels on the
GPU box (tests/golden/neighbour_ref.npz, tools/make_neighbour_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import cpu_oracle as O
from dcl_testutil import GOLDEN, cad_like_cloud, flat_bxyz, uniform_cloud


def d2_matrix(a, b):
    """fp32 squared distances with the reference's contraction order, vectorised via float64 fma emulation
    only where exactness is not asserted; for exact checks the oracle itself is the statement."""
    d = a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)
    return (d * d).sum(-1)


def test_opt_n_threads():
    for n, t in [(1, 1), (2, 2), (3, 2), (1000, 512), (1024, 1024), (1025, 1024), (16384, 1024), (37, 32), (5, 4)]:
        assert O.opt_n_threads(n) == t


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def fps_by_rule(xyz, m):
    """FPS with the tie rule stated in SURVEY.md D6: among equal maxima minimise (bitrev(k mod T), k)."""
    n = xyz.shape[0]
    T = O.opt_n_threads(n)
    bits = T.bit_length() - 1
    key = np.array([bitrev(k % T, bits) * (n + 1) + k for k in range(n)])
    temp = np.full(n, 1e10, np.float32)
    idx = [0]
    xyz64 = xyz.astype(np.float64)
    for _ in range(1, m):
        p = xyz[idx[-1]]
        dx, dy, dz = (xyz[:, 0] - p[0]).astype(np.float32), (xyz[:, 1] - p[1]).astype(np.float32), (xyz[:, 2] - p[2]).astype(np.float32)
        # fmaf(dz,dz,fmaf(dx,dx,dy*dy)) emulated: inner products are exact in float64, one rounding per fma
        inner = (dx.astype(np.float64) * dx + (dy * dy).astype(np.float32).astype(np.float64)).astype(np.float32)
        d = (dz.astype(np.float64) * dz + inner.astype(np.float64)).astype(np.float32)
        temp = np.minimum(d, temp)
        cand = np.flatnonzero(temp == temp.max())
        idx.append(int(cand[np.argmin(key[cand])]))
    return np.array(idx, np.int32)


@pytest.mark.parametrize("n,m", [(1000, 40), (1024, 64), (300, 30), (2500, 50)])
def test_fps_tie_break_rule(n, m):
    """Clouds with many duplicated points: the literal block simulation equals the closed-form rule."""
    xyz = cad_like_cloud(n, 1, n, n_unique=max(8, n // 6))[0].numpy()
    assert np.array_equal(O.furthest_point_sample(xyz[None], m)[0], fps_by_rule(xyz, m))


def test_fps_basic_properties():
    xyz = uniform_cloud(0, 3, 777).numpy()
    idx, temp = O.furthest_point_sample(xyz, 50, return_temp=True)
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 50
        d = d2_matrix(xyz[b], xyz[b][idx[b][:-1]]).min(1)  # the last pick is never used as a seed
        assert np.allclose(temp[b], d, rtol=1e-5, atol=1e-9)


def test_ball_query_semantics():
    xyz = uniform_cloud(1, 2, 600).numpy()
    new_xyz = xyz[:, :70].copy()
    new_xyz[:, -5:] += 10.0  # empty balls
    r, ns = 0.12, 6
    idx = O.ball_query(r, ns, xyz, new_xyz)
    for b in range(2):
        d = d2_matrix(new_xyz[b], xyz[b])
        for i in range(70):
            hits = np.flatnonzero(d[i] < np.float32(r) * np.float32(r) * (1 - 1e-6))
            sure = hits[:ns]
            if len(hits) == 0:
                assert (idx[b, i] == 0).all()
                continue
            row = idx[b, i]
            k = min(len(sure), ns)
            assert np.array_equal(row[:k], sure[:k]) or np.all(np.diff(row[:k]) > 0)
            assert (row[k:] == row[0]).all() or len(hits) >= ns


def test_three_nn_and_knn_agree_with_sorting():
    u, k = cad_like_cloud(2, 2, 200).numpy(), cad_like_cloud(3, 2, 90).numpy()
    d3, i3 = O.three_nn(u, k)
    dk, ik = O.knn(5, u, k)
    assert np.array_equal(i3, ik[:, :, :3]) and np.array_equal(d3, dk[:, :, :3])
    for b in range(2):
        d = d2_matrix(u[b], k[b])
        order = np.argsort(d, axis=1, kind="stable")[:, :5]
        exact_ties = np.take_along_axis(d, order, 1)
        # wherever the float64 distances are separated beyond fp32 noise the orders must agree
        sep = np.diff(exact_ties, axis=1) > 1e-9
        agree = (order[:, :4] == ik[b][:, :4]) | ~sep
        assert agree.all()
    assert np.all(np.diff(dk, axis=2) >= 0)


def test_knn_unfilled_slots():
    u, k = uniform_cloud(4, 1, 10).numpy(), uniform_cloud(5, 1, 3).numpy()
    d, i = O.knn(6, u, k)
    assert np.isinf(d[:, :, 3:]).all() and (i[:, :, 3:] == 0).all()  # (float)1e40 -> inf, idx 0


def test_sp_three_nn_matches_batched():
    b, n_per, m_per = 3, 40, 25
    u = flat_bxyz(6, b, n_per, shuffle=False).numpy()
    k = flat_bxyz(7, b, m_per, shuffle=False).numpy()
    d, i = O.sp_three_nn(u, k)
    db, ib = O.three_nn(u[:, 1:].reshape(b, n_per, 3), k[:, 1:].reshape(b, m_per, 3))
    assert np.array_equal(d.reshape(b, n_per, 3), db)
    assert np.array_equal(i.reshape(b, n_per, 3), ib + (np.arange(b) * m_per)[:, None, None])


def test_sp_three_nn_missing_batch_gives_inf_and_zero():
    u = flat_bxyz(8, 2, 10, shuffle=False).numpy()
    k = flat_bxyz(9, 2, 2, shuffle=True).numpy()
    k = k[k[:, 0] == 0]
    d, i = O.sp_three_nn(u, k)
    assert np.isinf(d[10:]).all() and (i[10:] == 0).all()
    assert np.isinf(d[:10, 2]).all() and np.isfinite(d[:10, :2]).all()


def test_gather_group_interpolate_and_grads():
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(2, 4, 50, generator=g, dtype=torch.float64)
    idx = torch.randint(0, 50, (2, 7, 3), generator=g)
    w = torch.rand(2, 7, 3, generator=g, dtype=torch.float64)
    f32 = feats.float().numpy()
    out = O.grouping_operation(f32, idx.int().numpy())
    assert np.array_equal(out, np.take_along_axis(f32[:, :, None, :].repeat(7, 2), idx.numpy()[:, None].repeat(4, 1), 3))
    assert np.array_equal(O.gather_operation(f32, idx[:, :, 0].int().numpy()), out[..., 0])
    want = (torch.gather(feats[:, :, None, :].expand(2, 4, 7, 50), 3, idx[:, None].expand(2, 4, 7, 3)) * w[:, None]).sum(-1)
    got = O.three_interpolate(f32, idx.int().numpy(), w.float().numpy())
    assert np.allclose(got, want.numpy(), rtol=1e-5, atol=1e-6)
    go = torch.randn(2, 4, 7, generator=g, dtype=torch.float64)
    f = feats.clone().requires_grad_(True)
    (torch.gather(f[:, :, None, :].expand(2, 4, 7, 50), 3, idx[:, None].expand(2, 4, 7, 3)) * w[:, None]).sum(-1).backward(go)
    gi = O.three_interpolate_grad(go.float().numpy(), idx.int().numpy(), w.float().numpy(), 50)
    assert np.allclose(gi, f.grad.numpy(), rtol=1e-4, atol=1e-5)
    gg = O.grouping_operation_grad(np.ones((2, 4, 7, 3), np.float32), idx.int().numpy(), 50)
    counts = np.stack([np.bincount(idx[b].flatten().numpy(), minlength=50) for b in range(2)])
    assert np.array_equal(gg, counts[:, None, :].repeat(4, 1).astype(np.float32))


REF_FIXTURE = os.path.join(GOLDEN, "neighbour_ref.npz")


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="fixture is produced on the GPU box by tools/make_neighbour_golden.py")
def test_c_oracle_equals_reference_kernels_fixture():
    """Bit-exact agreement of the C restatement with outputs of the reference's own CUDA kernels."""
    gold = np.load(REF_FIXTURE)
    g = lambda s: torch.Generator().manual_seed(s)
    a, b = cad_like_cloud(1, 2, 1000).numpy(), uniform_cloud(2, 1, 4096).numpy()
    i1, t1 = O.furthest_point_sample(a, 64, return_temp=True)
    assert np.array_equal(i1, gold["fps_cad_idx"]) and np.array_equal(t1, gold["fps_cad_temp"])
    assert np.array_equal(O.furthest_point_sample(b, 128), gold["fps_uni_idx"])
    xyz, cad = uniform_cloud(3, 2, 2048).numpy(), cad_like_cloud(4, 2, 1500).numpy()
    assert np.array_equal(O.ball_query(0.1, 16, xyz, xyz[:, :64]), gold["bq_uni"])
    assert np.array_equal(O.ball_query(0.02, 8, cad, cad[:, :50]), gold["bq_cad"])
    u, k = cad_like_cloud(5, 2, 512).numpy(), cad_like_cloud(6, 2, 200).numpy()
    d3, i3 = O.three_nn(u, k)
    dk, ik = O.knn(8, u, k)
    assert np.array_equal(d3, gold["nn3_d2"]) and np.array_equal(i3, gold["nn3_idx"])
    assert np.array_equal(dk, gold["knn_d2"]) and np.array_equal(ik, gold["knn_idx"])
    feats = torch.randn(2, 6, 200, generator=g(7)).numpy()
    w = torch.rand(2, 512, 3, generator=g(8)).numpy()
    assert np.array_equal(O.three_interpolate(feats, i3, w), gold["interp"])
    feats = torch.randn(2, 5, 300, generator=g(9)).numpy()
    idx = torch.randint(0, 300, (2, 20, 4), generator=g(10), dtype=torch.int32).numpy()
    assert np.array_equal(O.grouping_operation(feats, idx), gold["group"])
    assert np.array_equal(O.gather_operation(feats, idx[:, :, 0]), gold["gather"])
    u, k = flat_bxyz(11, 3, 100, shuffle=False), flat_bxyz(12, 3, 40)
    u[:, 1:] = (u[:, 1:] * 64).round() / 64
    k[:, 1:] = (k[:, 1:] * 64).round() / 64
    k = k[k[:, 0] != 1].contiguous()
    d, i = O.sp_three_nn(u.numpy(), k.numpy())
    assert np.array_equal(d, gold["sp_d2"]) and np.array_equal(i, gold["sp_idx"])
    feats = torch.randn(k.shape[0], 8, generator=g(13)).numpy()
    w = torch.rand(300, 3, generator=g(14)).numpy()
    assert np.array_equal(O.sp_three_interpolate(feats, i, w), gold["sp_interp"])


@pytest.mark.parametrize("seed,b,side,m,n_per", [(1, 3, 32, 1500, 200), (2, 2, 16, 300, 150), (3, 4, 8, 60, 100),
                                                  (4, 1, 64, 40, 64), (5, 8, 64, 12, 32)])
def test_slab_walk_pruning_rule_is_exact(seed, b, side, m, n_per):
    """The search the product runs on the pyramid levels (slabs of equal first voxel index, walked outwards, pruned by
    the plane distance with a 1e-6 margin, candidates ranked by (d, index)) returns exactly what the reference-order
    scan returns — including queries on voxel corners, where up to eight centres tie, instances with fewer than three
    voxels, and voxels given in shuffled order — while visiting a fraction of the candidates."""
    rng = np.random.default_rng(seed)
    vox = np.concatenate([rng.integers(0, b, (m, 1)), rng.integers(0, side, (m, 3))], 1).astype(np.int32)
    vox = np.unique(vox, axis=0)
    vox = vox[rng.permutation(len(vox))]
    ext = np.full(3, 0.6 / side, np.float32)
    off = np.full(3, -0.3, np.float32)
    centres = ((vox[:, 1:].astype(np.float32) * ext) + off) + np.float32(0.5) * ext      # torch's evaluation order
    known = np.concatenate([vox[:, :1].astype(np.float32), centres.astype(np.float32)], 1)
    ids = np.repeat(np.arange(b), n_per).astype(np.float32)[:, None]
    pts = ((rng.random((b * n_per, 3)) - 0.5) * 0.7).astype(np.float32)
    corner = (rng.integers(0, side + 1, (b * n_per, 3)).astype(np.float32) * ext + off).astype(np.float32)
    pts[::3] = corner[::3]                                                                # distance ties
    unknown = np.concatenate([ids, pts], 1).astype(np.float32)
    d_ref, i_ref = O.sp_three_nn(unknown, known)
    d_got, i_got, visited = O.sp_three_nn_slab_model(unknown, vox, ext, off, side)
    assert np.array_equal(i_got, i_ref)
    assert np.array_equal(d_got, d_ref.astype(np.float32))
    full = sum(int((vox[:, 0] == k).sum()) for k in range(b)) * n_per
    if len(vox) > 200:
        assert visited < 0.6 * full, (visited, full)
