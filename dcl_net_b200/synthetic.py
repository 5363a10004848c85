"""Synthetic inputs of the hot path (there are no datasets or checkpoints offline).

`backbone_levels` stands in for the two sparse-conv towers (out of this path's scope): for a batch of
point clouds it produces the four-level voxel pyramid they would output — occupied voxels at scales
[2,4,6,8] x unit voxel (models/DCL_Net.py:54), dilated by a 3x3x3 stencil as SparseConv3d/SparseAvgPool3d
do (models/Modules.py:151-157), random features of width (32,64,128,256), rows shuffled across batch items
(spconv does not keep them batch-sorted).  SURVEY.md §8d config 3.
"""
import types

import torch


def object_clouds(seed, b, n, partial=False):
    """(b*n, 3) points in metres sampled on closed object-sized SURFACES (a randomly oriented ellipsoid with
    semi-axes 3-8 cm per instance, 0.5 mm noise), like the CAD templates / depth crops the reference feeds
    (YCB objects: radius 0.046-0.159 m).  Surface clouds give the voxel-pyramid occupancy the
    survey measured on the real CAD clouds (~800 / 300 / 130 / 50 rows per instance at levels 1-4); points filling
    a cube would give ~4x more.  partial=True keeps only the half facing +z (a single-view observation)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(b, n, 3, generator=g)
    if partial:
        d[..., 2] = d[..., 2].abs()
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    axes = 0.03 + 0.05 * torch.rand(b, 1, 3, generator=g)
    q, _ = torch.linalg.qr(torch.randn(b, 3, 3, generator=g))
    pts = (d * axes) @ q.transpose(1, 2) + 0.0005 * torch.randn(b, n, 3, generator=g)
    return pts.reshape(b * n, 3).contiguous()


def backbone_levels(seed, points, b, unit=0.006, scales=(2, 4, 6, 8), channels=(32, 64, 128, 256), shuffle=True):
    g = torch.Generator().manual_seed(seed)
    n_per = points.shape[0] // b
    ids = torch.arange(b).repeat_interleave(n_per)
    offset = -0.5 * unit * 64  # models/Modules.py:234
    stencil = torch.stack(torch.meshgrid(*([torch.arange(-1, 2)] * 3), indexing="ij"), -1).reshape(-1, 3)
    levels = []
    for scale, ch in zip(scales, channels):
        ext = unit * scale
        lim = 64 // scale + (1 if 64 % scale else 0)
        vox = torch.floor((points - offset) / ext).long().clamp(0, lim - 1)
        vox = (vox[:, None, :] + stencil[None]).reshape(-1, 3)
        bid = ids[:, None].expand(-1, 27).reshape(-1, 1)
        keep = ((vox >= 0) & (vox < lim)).all(1)
        ind = torch.unique(torch.cat([bid, vox], 1)[keep], dim=0).int()
        if shuffle:
            ind = ind[torch.randperm(ind.shape[0], generator=g)]
        feats = torch.randn(ind.shape[0], ch, generator=g)
        levels.append(types.SimpleNamespace(features=feats, indices=ind.contiguous()))
    return levels


def levels_to(levels, device, non_blocking=False):
    return [types.SimpleNamespace(features=l.features.to(device, non_blocking=non_blocking),
                                  indices=l.indices.to(device, non_blocking=non_blocking)) for l in levels]


def point_colours(seed, n_points):
    """(n,3) RGB in [0,1): the per-point colour channels of the dataloader's [1, rgb, xyz] features."""
    return torch.rand(n_points, 3, generator=torch.Generator().manual_seed(seed))
