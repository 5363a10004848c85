"""TEST INFRASTRUCTURE — PyTorch restatement of the input side of Network.forward:
voxelisation (libs/pointgroup_ops) and the two sparse-conv towers (libs/spconv, models/Modules.py:100-159).

PARITY UNPINNED by the reference: spconv / pointgroup_ops cannot be built here (THC, google dense_hash_map, cmake),
so no output of the reference itself pins these restatements.  What they follow, line by line:

  voxelization_idx      libs/pointgroup_ops/src/voxelize/voxelize.cpp:57-163   voxels numbered in order of first
                        appearance; output_coords = coords of a voxel's first point; rule rows [n, i_1 .. i_n, 0 ..]
  voxelization (mean)   voxelize.cu:10-23   out[m] = sum over the rule's points, in rule order, of (1/n) * feats[i]
  get_indice_pairs      libs/spconv/include/spconv/spconv_ops.h:27-136, indice.cu.h:24-220, geometry.h:24-86
                        SparseConv3d: output set = every in-grid position some input reaches through the kernel,
                        rows ordered by linear index ((b*X + x)*Y + y)*Z + z  (torch::_unique sorts, spconv_ops.h:122);
                        SubMConv3d: output set and order = the input's; kernel offset k = (k0*3 + k1)*3 + k2 pairs
                        output o with input o*stride - padding + k.
  indice_conv           spconv_ops.h:253-349   out[o] += in[i] @ W[k] over the pairs of offset k, offsets ascending
  SparseAvgPool3d       spconv/pool.py:225-247, functional.py:135-162 (use_gs=False), src/spconv/avgpool.cu:27-53,
                        summaryRF.cu:27-41   out[o] = sum over offsets ascending of in[i] / rf[o], rf[o] = number
                        of inputs in o's window
  Backbone_SPCONV       models/Modules.py:100-159 with dims [7,16,32,32,64,64,128,128,256], stride layers [1,3,5]
                        (models/DCL_Net.py:47-52): module_i = [SparseConv3d k3 s1 p1 -> BN1d -> ReLU,
                        SubMConv3d k3 -> BN1d -> ReLU], each followed by SparseAvgPool3d(k3, s2, p1).

tests/test_cpu_oracle.py checks this restatement against an independent dense formulation (F.conv3d / F.avg_pool3d
on the zero-filled grid) so that the pairing and ordering rules above are at least self-consistent.
"""
import types

import torch
import torch.nn as nn


# ------------------------------------------------------------------------------------ voxelisation
def voxel_indices_from_points(points, unit=0.006, limit=64):
    """YCBV/dataloader_test_YCBV.py:177,186: (p + 0.5*total_extent) / unit in fp32, truncated (.long())."""
    total = torch.tensor(unit * limit, dtype=torch.float32)
    return ((points + total * 0.5) / torch.tensor(unit, dtype=torch.float32)).long()


def voxelization_idx(coords, batch_size, mode=4):
    """coords (N,4) long bxyz -> (output_coords (M,4) long, input_map (N,) int32, output_map (M, maxActive+1) int32)."""
    assert mode == 4
    seen, rows = {}, []
    input_map = torch.empty(coords.shape[0], dtype=torch.int32)
    for i, c in enumerate(coords.tolist()):
        key = tuple(c)
        v = seen.get(key)
        if v is None:
            v = seen[key] = len(rows)
            rows.append([])
        rows[v].append(i)
        input_map[i] = v
    max_active = max(len(r) for r in rows)
    output_map = torch.zeros(len(rows), max_active + 1, dtype=torch.int32)
    output_coords = torch.zeros(len(rows), 4, dtype=torch.long)
    for v, r in enumerate(rows):
        output_map[v, 0] = len(r)
        output_map[v, 1:1 + len(r)] = torch.tensor(r, dtype=torch.int32)
        output_coords[v] = coords[r[0]]
    return output_coords, input_map, output_map


def voxelization_mean(feats, output_map):
    """voxelize.cu:10-23 with average=True: one thread per plane adds multiplier*inp in rule order (fp32)."""
    m, c = output_map.shape[0], feats.shape[1]
    out = torch.zeros(m, c, dtype=feats.dtype)
    n = output_map[:, 0].long()
    mult = (1.0 / n.to(feats.dtype)).unsqueeze(1)
    for j in range(1, output_map.shape[1]):
        live = (n >= j)
        idx = output_map[:, j].long()
        out[live] = out[live] + mult[live] * feats[idx[live]]
    return out


# ------------------------------------------------------------------------------------ sparse tensors
def sparse_tensor(features, indices, spatial_shape, batch_size):
    return types.SimpleNamespace(features=features, indices=indices.int(), spatial_shape=list(spatial_shape),
                                 batch_size=batch_size)


def _linear(indices, shape):
    i = indices.long()
    return ((i[:, 0] * shape[0] + i[:, 1]) * shape[1] + i[:, 2]) * shape[2] + i[:, 3]


def _unlinear(keys, shape):
    z = keys % shape[2]
    y = (keys // shape[2]) % shape[1]
    x = (keys // (shape[2] * shape[1])) % shape[0]
    b = keys // (shape[2] * shape[1] * shape[0])
    return torch.stack([b, x, y, z], 1).int()


def get_indice_pairs(indices, spatial_shape, ksize=3, stride=1, padding=1, subm=False):
    """-> (out_indices (Mo,4) int32, out_shape, pairs: list over the k^3 offsets of (in_rows, out_rows) long tensors)."""
    dev = indices.device
    if subm:
        out_shape = list(spatial_shape)
        stride, padding = 1, ksize // 2
    else:
        out_shape = [(s + 2 * padding - (ksize - 1) - 1) // stride + 1 for s in spatial_shape]
    ind = indices.long()
    cand = []   # per offset: (input row, output linear key)
    for k0 in range(ksize):
        for k1 in range(ksize):
            for k2 in range(ksize):
                k = torch.tensor([k0, k1, k2], device=dev)
                num = ind[:, 1:] + padding - k            # = out * stride
                ok = (num % stride == 0).all(1)
                o = torch.div(num, stride, rounding_mode="floor")
                ok &= ((o >= 0) & (o < torch.tensor(out_shape, device=dev))).all(1)
                rows = ok.nonzero().squeeze(1)
                key = _linear(torch.cat([ind[rows, :1], o[rows]], 1), out_shape)
                cand.append((rows, key))
    if subm:
        in_keys = _linear(ind, out_shape)
        order = torch.argsort(in_keys)
        sorted_keys = in_keys[order]
        pairs = []
        for rows, key in cand:
            pos = torch.searchsorted(sorted_keys, key).clamp_max(sorted_keys.numel() - 1)
            hit = sorted_keys[pos] == key
            pairs.append((rows[hit], order[pos[hit]]))
        return indices.int(), out_shape, pairs
    all_keys = torch.unique(torch.cat([key for _, key in cand]))      # sorted, as torch::_unique gives the reference
    pairs = [(rows, torch.searchsorted(all_keys, key)) for rows, key in cand]
    return _unlinear(all_keys, out_shape), out_shape, pairs


def indice_conv(features, weight, pairs, num_out):
    """weight (k,k,k,Cin,Cout); out[o] += in[i] @ W[k], offsets ascending (spconv_ops.h:293-340)."""
    w = weight.reshape(-1, weight.shape[-2], weight.shape[-1])
    out = torch.zeros(num_out, w.shape[2], dtype=features.dtype, device=features.device)
    for k, (rows_in, rows_out) in enumerate(pairs):
        if rows_in.numel():
            out.index_add_(0, rows_out, features[rows_in] @ w[k])
    return out


def indice_avgpool(features, pairs, num_out):
    rf = torch.zeros(num_out, dtype=torch.int32, device=features.device)
    for _, rows_out in pairs:
        rf.index_add_(0, rows_out, torch.ones_like(rows_out, dtype=torch.int32))
    out = torch.zeros(num_out, features.shape[1], dtype=features.dtype, device=features.device)
    div = rf.to(features.dtype).unsqueeze(1)
    for rows_in, rows_out in pairs:     # within an offset every output row occurs at most once
        if rows_in.numel():
            out[rows_out] = out[rows_out] + features[rows_in] / div[rows_out]
    return out


class SparseConv3dO(nn.Module):
    """Parameter layout of spconv.SparseConv3d / SubMConv3d (conv.py:98-111): weight (k,k,k,Cin,Cout), no bias."""

    def __init__(self, cin, cout, subm):
        super().__init__()
        self.subm = subm
        self.weight = nn.Parameter(torch.empty(3, 3, 3, cin, cout))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        out_ind, out_shape, pairs = get_indice_pairs(x.indices, x.spatial_shape, 3, 1, 1, self.subm)
        feats = indice_conv(x.features, self.weight, pairs, out_ind.shape[0])
        return sparse_tensor(feats, out_ind, out_shape, x.batch_size)


class BlockO(nn.Module):
    """BasicBlock_SPCONV (models/Modules.py:11-56): conv -> BatchNorm1d -> ReLU, parameters under `layers`."""

    def __init__(self, cin, cout, subm):
        super().__init__()
        self.layers = nn.Sequential(SparseConv3dO(cin, cout, subm), nn.BatchNorm1d(cout), nn.ReLU())

    def forward(self, x):
        y = self.layers[0](x)
        y.features = self.layers[2](self.layers[1](y.features))
        return y


def avg_pool(x):
    out_ind, out_shape, pairs = get_indice_pairs(x.indices, x.spatial_shape, 3, 2, 1, False)
    return sparse_tensor(indice_avgpool(x.features, pairs, out_ind.shape[0]), out_ind, out_shape, x.batch_size)


class BackboneOracle(nn.Module):
    """Backbone_SPCONV (models/Modules.py:100-159): same module / parameter names (module1..4.{0,1}.layers.*)."""

    def __init__(self, dims=(7, 16, 32, 32, 64, 64, 128, 128, 256), stride_layers=(1, 3, 5)):
        super().__init__()
        modules = [[] for _ in range(len(stride_layers) + 1)]
        mi = 0
        for i in range(len(dims) - 1):
            subm = not ((i - 1) in stride_layers or i == 0)
            modules[mi].append(BlockO(dims[i], dims[i + 1], subm))
            if i in stride_layers:
                mi += 1
        self.module1, self.module2, self.module3, self.module4 = (nn.Sequential(*m) for m in modules)

    def forward(self, x):
        feats = []
        for mod in (self.module1, self.module2, self.module3, self.module4):
            x = avg_pool(mod(x))
            feats.append(x)
        return feats


def tower_input(points, rgb, b, unit=0.006, limit=64):
    """What Network.forward builds for one tower from the dataloader's tensors (models/DCL_Net.py:157-176):
    per-point features [1, rgb, xyz] -> mean-voxelised (M,7) features + (M,4) int indices."""
    n = points.shape[0] // b
    feats = torch.cat([torch.ones(points.shape[0], 1), rgb, points], 1)
    ids = torch.arange(b).repeat_interleave(n).view(-1, 1)
    coords = torch.cat([ids, voxel_indices_from_points(points, unit, limit)], 1)
    out_coords, _, out_map = voxelization_idx(coords, b, 4)
    return sparse_tensor(voxelization_mean(feats, out_map), out_coords, [limit] * 3, b), (out_coords, out_map)
