"""GPU parity of the input side: voxelisation, voxel sets / rulebooks, sparse convolutions
and average pools of the two towers, against oracle/backbone_oracle.py (a restatement of libs/pointgroup_ops and
libs/spconv — PARITY UNPINNED by the reference, which cannot be built here; the restatement is checked against a
dense formulation in tests/test_backbone_oracle.py).

Bars: voxel indices, maps and row orders bit-exact; voxel means and average pools bit-exact (fixed summation order);
convolutions fp32-faithful given their fp16-rounded operand rows (1e-5 of the output scale) and within 2e-3 of the
fp32 towers end to end (activations are rounded once to fp16 between layers, as everywhere on the inference path)."""
import numpy as np
import pytest
import torch

from oracle import backbone_oracle as BO
from dcl_testutil import rel_err

pytestmark = pytest.mark.gpu

from dcl_net_b200 import _lib as L                                   # noqa: E402
from dcl_net_b200 import backbone as BB                               # noqa: E402
from dcl_net_b200 import synthetic                                    # noqa: E402


def _clouds(seed, b, n=1024):
    pts_a = synthetic.object_clouds(seed, b, n, partial=True)
    pts_b = synthetic.object_clouds(seed + 1, b, n)
    g = torch.Generator().manual_seed(seed + 2)
    return pts_a, torch.rand(b * n, 3, generator=g), pts_b, torch.rand(b * n, 3, generator=g)


def _towers(seed, b, dev, n=1024):
    torch.manual_seed(seed)
    oracle = [BO.BackboneOracle().eval(), BO.BackboneOracle().eval()]
    for o in oracle:                      # non-trivial BatchNorm statistics
        for m in o.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.normal_(0, 0.1)
    mine = [BB.Backbone_SPCONV().eval(), BB.Backbone_SPCONV().eval()]
    for m, o in zip(mine, oracle):
        missing = m.load_state_dict(o.state_dict())
        assert not missing.missing_keys and not missing.unexpected_keys
    clouds = _clouds(seed, b, n)
    caps = BB.SparseTowers.plan_capacities([clouds[0], clouds[2]], b, n, dev)
    towers = BB.SparseTowers(mine[0].to(dev), mine[1].to(dev), dev, b, n, caps)
    return oracle, towers, clouds


def _oracle_chain(x):
    """Index sets of one tower: [(conv-out indices, pool-out indices)] per level, via the restated get_indice_pairs."""
    out = []
    ind, shape = x.indices, x.spatial_shape
    for _ in range(4):
        ci, cshape, _ = BO.get_indice_pairs(ind, shape, 3, 1, 1, False)
        pi, pshape, _ = BO.get_indice_pairs(ci, cshape, 3, 2, 1, False)
        out.append((ci, pi))
        ind, shape = pi, pshape
    return out


@pytest.mark.parametrize("b", [1, 3])
def test_voxel_sets_and_row_order(cuda_dev, b):
    """Every conv-out / pool-out set of both towers: same voxels in the same (batch, linear index) order."""
    oracle, towers, clouds = _towers(10 + b, b, cuda_dev)
    towers.run(*(c.to(cuda_dev) for c in clouds))
    towers.check_errors()
    for tower, (pts, rgb) in enumerate(((clouds[0], clouds[1]), (clouds[2], clouds[3]))):
        x, _ = BO.tower_input(pts, rgb, b)
        assert towers.set_rows(tower, 0) == x.indices.shape[0]
        for level, (ci, pi) in enumerate(_oracle_chain(x)):
            for s, want in ((2 * level + 1, ci), (2 * level + 2, pi)):
                n = towers.set_rows(tower, s)
                assert n == want.shape[0], (tower, s)
                got = towers.t[tower].indices[s]
                assert torch.equal(got[:n].cpu(), want.int()), (tower, s)
                assert bool((got[n:, 0] == b).all()), "unused rows must carry batch id == B"


def test_voxelization_idx_and_mean_match_reference_semantics(cuda_dev):
    b, n = 3, 512
    g = torch.Generator().manual_seed(4)
    pts = synthetic.object_clouds(9, b, n)
    pts[5] = pts[2]
    pts[n + 7] = pts[n + 1]                                           # several points per voxel
    rgb = torch.rand(b * n, 3, generator=g)
    ids = torch.arange(b).repeat_interleave(n).view(-1, 1)
    coords = torch.cat([ids, BO.voxel_indices_from_points(pts)], 1)
    want_c, want_in, want_out = BO.voxelization_idx(coords, b, 4)
    got_c, got_in, got_out = BB.voxelization_idx(coords.to(cuda_dev), b, 4)
    assert torch.equal(got_c.cpu(), want_c)
    assert torch.equal(got_in.cpu(), want_in)
    assert torch.equal(got_out.cpu(), want_out)
    feats = torch.cat([torch.ones(b * n, 1), rgb, pts], 1)
    got_f = BB.voxelization(feats.to(cuda_dev), got_out, 4)
    assert torch.equal(got_f.cpu(), BO.voxelization_mean(feats, want_out))   # bit-exact: same order, same roundings


def test_first_layer_operand_rows(cuda_dev):
    """feat16 = [mean | fp16 remainder] of the voxel means, voxels in sorted order: hi + lo carries ~21 bits."""
    b = 2
    oracle, towers, clouds = _towers(3, b, cuda_dev)
    towers.run(*(c.to(cuda_dev) for c in clouds))
    x, _ = BO.tower_input(clouds[0], clouds[1], b)
    keys = ((x.indices[:, 0].long() * 64 + x.indices[:, 1]) * 64 + x.indices[:, 2]) * 64 + x.indices[:, 3]
    order = torch.argsort(keys)
    want = x.features[order]                                         # sorted (batch, linear index) order
    f16 = towers.t[0].feat16.float().cpu().view(b, 1024, 16)
    cnt = torch.bincount(x.indices[:, 0].long(), minlength=b)
    got = torch.cat([f16[i, :cnt[i]] for i in range(b)])
    val = got[:, :7].clone()
    val[:, 1:7] += got[:, 7:13]
    assert (val - want).abs().max().item() <= 2.0 ** -20 * want.abs().max().item()
    assert bool((got[:, 13:] == 0).all())


def _oracle_pairs(ind_in, shape_in, subm, stride=1):
    return BO.get_indice_pairs(ind_in, shape_in, 3, stride, 1, subm)


@pytest.mark.parametrize("b", [2])
def test_each_layer_given_its_own_input(cuda_dev, b):
    """Layer by layer: every convolution against fp64 indice_conv over the SAME fp16 operand rows the kernel read and
    the folded weights (isolates the kernel's arithmetic: 2 MMAs with fp16 hi/lo weights, fp32 accumulation), and
    every average pool bit-exact against the restated avgpool over the same fp32 input."""
    oracle, towers, clouds = _towers(21, b, cuda_dev)
    towers.run(*(c.to(cuda_dev) for c in clouds))
    towers.check_errors()
    for tower, (pts, rgb) in enumerate(((clouds[0], clouds[1]), (clouds[2], clouds[3]))):
        t = towers.t[tower]
        x, _ = BO.tower_input(pts, rgb, b)
        keys = ((x.indices[:, 0].long() * 64 + x.indices[:, 1]) * 64 + x.indices[:, 2]) * 64 + x.indices[:, 3]
        order = torch.argsort(keys)
        ind, shape = x.indices[order], x.spatial_shape
        # operand rows of level 0 in dense sorted order (the kernel reads them from per-instance slots)
        cnt = torch.bincount(ind[:, 0].long(), minlength=b)
        f16 = t.feat16.float().cpu().view(b, 1024, 16)
        feat = torch.cat([f16[i, :cnt[i]] for i in range(b)]).double()
        blocks = [blk for mod in (oracle[tower].module1, oracle[tower].module2, oracle[tower].module3,
                                  oracle[tower].module4) for blk in mod]
        for level in range(4):
            for j, subm in enumerate((False, True)):
                blk = blocks[2 * level + j]
                conv, bn = blk.layers[0], blk.layers[1]
                scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).double()
                shift = (bn.bias - bn.running_mean * scale.float()).double()
                w = conv.weight.detach().double() * scale.view(1, 1, 1, 1, -1)
                if level == 0 and j == 0:
                    w16 = torch.zeros(3, 3, 3, 16, w.shape[4], dtype=torch.float64)
                    for c, src in enumerate(BB._CONV0_CHANNELS):
                        if src >= 0:
                            w16[:, :, :, c] = w[:, :, :, src]
                    w = w16
                out_ind, out_shape, pairs = _oracle_pairs(ind, shape, subm)
                want = torch.relu(BO.indice_conv(feat, w, pairs, out_ind.shape[0]) + shift)
                n = out_ind.shape[0]
                got = (t.conv16[level] if j == 0 else t.subm32[level])[:n].double().cpu()
                tol = 1e-3 if j == 0 else 1e-5          # conv-out rows are stored in fp16 (half an ulp = 4.9e-4)
                assert (got - want).abs().max().item() <= tol * want.abs().max().item(), (tower, level, j)
                ind, shape = out_ind, out_shape
                feat = got                              # the next layer reads exactly what this one stored
            p_ind, p_shape, p_pairs = _oracle_pairs(ind, shape, False, stride=2)
            n = p_ind.shape[0]
            src32 = t.subm32[level][:ind.shape[0]].cpu()
            want_pool = BO.indice_avgpool(src32, p_pairs, n)
            assert torch.equal(t.pool32[level][:n].cpu(), want_pool), (tower, level, "avgpool")
            if level < 3:
                assert torch.equal(t.pool16[level][:n].cpu(), want_pool.to(torch.float16))
            ind, shape = p_ind, p_shape
            feat = want_pool.to(torch.float16).double()


@pytest.mark.parametrize("b", [4])
def test_towers_end_to_end_vs_fp32_oracle(cuda_dev, b):
    """The four pyramid levels of both towers against the fp32 restated towers on the same clouds."""
    oracle, towers, clouds = _towers(33, b, cuda_dev)
    levels = towers.run(*(c.to(cuda_dev) for c in clouds))
    towers.check_errors()
    for tower, (pts, rgb) in enumerate(((clouds[0], clouds[1]), (clouds[2], clouds[3]))):
        x, _ = BO.tower_input(pts, rgb, b)
        with torch.no_grad():
            want = oracle[tower](x)
        for level in range(4):
            n = want[level].indices.shape[0]
            assert torch.equal(levels[tower][level].indices[:n].cpu(), want[level].indices.int())
            assert rel_err(levels[tower][level].features[:n], want[level].features) < 2e-3, (tower, level)
