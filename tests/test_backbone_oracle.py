"""CPU: the sparse-conv / voxelisation restatement (oracle/backbone_oracle.py; parity unpinned by the reference, which
cannot be built here) against an independent dense formulation of the same operators, and its ordering rules."""
import torch
import torch.nn.functional as F

from oracle import backbone_oracle as BO


def _random_sparse(seed, b, grid, m, c):
    g = torch.Generator().manual_seed(seed)
    ind = torch.cat([torch.randint(0, b, (m, 1), generator=g), torch.randint(0, grid, (m, 3), generator=g)], 1)
    ind = torch.unique(ind, dim=0)
    ind = ind[torch.randperm(ind.shape[0], generator=g)]          # arbitrary row order, as voxelisation hands it over
    return BO.sparse_tensor(torch.randn(ind.shape[0], c, generator=g), ind, [grid] * 3, b)


def _dense(x, c):
    d = torch.zeros(x.batch_size, c, *x.spatial_shape)
    i = x.indices.long()
    d[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = x.features
    return d


def test_sparse_conv_equals_dense_conv_on_its_output_set():
    x = _random_sparse(0, 2, 9, 60, 5)
    conv = BO.SparseConv3dO(5, 6, subm=False)
    y = conv(x)
    dense = F.conv3d(_dense(x, 5), conv.weight.permute(4, 3, 0, 1, 2), padding=1)
    i = y.indices.long()
    assert torch.allclose(y.features, dense[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]], atol=1e-5)
    # output set = dilation of the input set by the 3x3x3 stencil (in-grid), rows sorted by linear index
    ones = _dense(BO.sparse_tensor(torch.ones(x.indices.shape[0], 1), x.indices, x.spatial_shape, 2), 1)
    occ = F.max_pool3d(ones, 3, 1, 1) > 0
    assert occ.sum().item() == y.indices.shape[0]
    assert bool(occ[i[:, 0], 0, i[:, 1], i[:, 2], i[:, 3]].all())
    keys = ((i[:, 0] * 9 + i[:, 1]) * 9 + i[:, 2]) * 9 + i[:, 3]
    assert bool((keys[1:] > keys[:-1]).all())


def test_subm_conv_equals_dense_conv_on_the_input_set():
    x = _random_sparse(1, 3, 8, 80, 4)
    conv = BO.SparseConv3dO(4, 7, subm=True)
    y = conv(x)
    assert torch.equal(y.indices, x.indices)
    dense = F.conv3d(_dense(x, 4), conv.weight.permute(4, 3, 0, 1, 2), padding=1)
    i = x.indices.long()
    assert torch.allclose(y.features, dense[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]], atol=1e-5)


def test_avg_pool_averages_over_present_inputs():
    x = _random_sparse(2, 2, 8, 70, 3)
    y = BO.avg_pool(x)
    assert y.spatial_shape == [4, 4, 4]
    dx = _dense(x, 3)
    ones = _dense(BO.sparse_tensor(torch.ones(x.indices.shape[0], 1), x.indices, x.spatial_shape, 2), 1)
    s = F.avg_pool3d(dx, 3, 2, 1, count_include_pad=True) * 27
    n = F.avg_pool3d(ones, 3, 2, 1, count_include_pad=True) * 27
    i = y.indices.long()
    want = s[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] / n[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]].round()
    assert torch.allclose(y.features, want, atol=1e-5)
    assert int((n.round() > 0).sum()) == y.indices.shape[0]


def test_voxelisation_first_appearance_order_and_mean():
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(2 * 50, 3, generator=g) - 0.5) * 0.05
    pts[7] = pts[3]                       # two points in one voxel
    rgb = torch.rand(100, 3, generator=g)
    x, (coords, omap) = BO.tower_input(pts, rgb, 2)
    assert x.features.shape[1] == 7 and torch.equal(x.indices.long(), coords)
    assert torch.allclose(x.features[:, 0], torch.ones(x.features.shape[0]))       # mean of the constant-1 channel
    first = omap[:, 1].long()
    assert bool((first[1:] > first[:-1]).all())                                    # numbered by first appearance
    v = (omap[:, 0] > 1).nonzero()[0, 0]
    members = omap[v, 1:1 + int(omap[v, 0])].long()
    assert torch.allclose(x.features[v, 4:], pts[members].mean(0), atol=1e-7)


def test_backbone_shapes_and_parameter_names():
    torch.manual_seed(0)
    net = BO.BackboneOracle().eval()
    names = [n for n, _ in net.named_parameters()]
    assert "module1.0.layers.0.weight" in names and "module4.1.layers.1.bias" in names
    assert net.module1[0].layers[0].weight.shape == (3, 3, 3, 7, 16) and not net.module1[0].layers[0].subm
    assert net.module1[1].layers[0].subm and not net.module2[0].layers[0].subm
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(2 * 64, 3, generator=g)
    pts = 0.05 * pts / pts.norm(dim=1, keepdim=True)
    x, _ = BO.tower_input(pts, torch.rand(128, 3, generator=g), 2)
    with torch.no_grad():
        f1, f2, f3, f4 = net(x)
    assert [f.features.shape[1] for f in (f1, f2, f3, f4)] == [32, 64, 128, 256]
    assert [f.spatial_shape[0] for f in (f1, f2, f3, f4)] == [32, 16, 8, 4]
