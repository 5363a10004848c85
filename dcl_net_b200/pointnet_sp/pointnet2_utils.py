"""Drop-in for the reference's libs/pointnet_sp/pointnet2_utils.py (flat, batch id in column 0).

    three_nn(unknown (N,4) bxyz, known (M,4) bxyz) -> (dist (N,3), idx (N,3) int32)      :9-38
    three_interpolate(features (M,C), idx (n,3), weight (n,3)) -> (n,C)                 :41-86
Like the reference these assert contiguity instead of copying.  three_nn runs the segmented
search (per-batch buckets built on the device); results are bit-identical to the reference's
full scan.  `nn_interpolate` is the fused form of models/Modules.py:213-226.
"""
from typing import Tuple

import torch
from torch.autograd import Function

from .. import _lib as L

_ws_cache = {}


def _workspace(nbytes, device):
    """Per-(device, stream) scratch for the bucket build; grown on demand, never shared across
    streams (the kernels of one call are ordered on their stream, so reuse on it is safe)."""
    if torch.cuda.is_current_stream_capturing():
        # a captured graph gets scratch of its own (from its private pool): graphs captured on the same capture
        # stream would otherwise share one buffer and race when they are replayed on different streams
        return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        L.require(unknown, torch.float32, "unknown")
        L.require(known, torch.float32, "known")
        N, m = unknown.size(0), known.size(0)
        dist2 = torch.empty(N, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(N, 3, dtype=torch.int32, device=unknown.device)
        lib = L.load()
        nbytes = lib.dcl_sp_three_nn_workspace_bytes(N, m)
        ws = _workspace(nbytes, unknown.device)
        L.check(lib.dcl_sp_three_nn_segmented(N, m, L.ptr(unknown), L.ptr(known), L.ptr(dist2), L.ptr(idx),
                                              L.ptr(ws), ws.numel(), L.stream_ptr()), "pointnet_sp.three_nn")
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


def three_nn_full_scan(unknown, known):
    """The reference-signature entry point (no scratch, O(N*M) scan); same results as three_nn."""
    assert unknown.is_contiguous() and known.is_contiguous()
    N, m = unknown.size(0), known.size(0)
    dist2 = torch.empty(N, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(N, 3, dtype=torch.int32, device=unknown.device)
    L.check(L.load().dcl_sp_three_nn_kernel_launcher_fast(N, m, L.ptr(unknown), L.ptr(known), L.ptr(dist2),
                                                          L.ptr(idx), L.stream_ptr()), "pointnet_sp.three_nn")
    return torch.sqrt(dist2), idx


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        L.require(features, torch.float32, "features")
        L.require(idx, torch.int32, "idx")
        L.require(weight, torch.float32, "weight")
        m, c = features.size()
        n = idx.size(0)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty(n, c, dtype=torch.float32, device=features.device)
        L.check(L.load().dcl_sp_three_interpolate_kernel_launcher_fast(
            c, m, n, L.ptr(features), L.ptr(idx), L.ptr(weight), L.ptr(output), L.stream_ptr()),
            "pointnet_sp.three_interpolate")
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        n, c = grad_out.size()
        grad_features = torch.zeros(m, c, dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.data.contiguous()
        L.check(L.load().dcl_sp_three_interpolate_grad_kernel_launcher_fast(
            c, n, m, L.ptr(grad_out_data), L.ptr(idx), L.ptr(weight), L.ptr(grad_features), L.stream_ptr()),
            "pointnet_sp.three_interpolate backward")
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


def nn_interpolate(target_points, query_points, query_feats, out=None, out_col0=0):
    """Fused three_nn -> inverse-distance weights -> three_interpolate (inference only, no autograd).
    Writes columns [out_col0, out_col0+C) of `out` (n, >=C) when given, so the four pyramid levels
    land directly in the concatenated (n, 480) buffer of models/Modules.py:250."""
    assert target_points.is_contiguous() and query_points.is_contiguous() and query_feats.is_contiguous()
    n, m, c = target_points.size(0), query_points.size(0), query_feats.size(1)
    if out is None:
        out = torch.empty(n, c, dtype=torch.float32, device=query_feats.device)
        out_col0 = 0
    assert out.is_contiguous() and out.size(0) == n and out.size(1) >= out_col0 + c
    lib = L.load()
    ws = _workspace(lib.dcl_sp_three_nn_workspace_bytes(n, m), target_points.device)
    L.check(lib.dcl_sp_nn_interpolate_fused(n, m, c, L.ptr(target_points), L.ptr(query_points), L.ptr(query_feats),
                                            L.ptr(out), out.size(1), out_col0, L.ptr(ws), ws.numel(),
                                            L.stream_ptr()), "pointnet_sp.nn_interpolate")
    return out


def nn_interpolate_pm(target_points, query_points, query_feats, out_pm, c_total, out_col0):
    """nn_interpolate writing columns [out_col0, out_col0+C) of a PM image (operand format of the tensor-core
    MLP kernels, include/dcl_b200.h) with c_total channels; n must be a multiple of 128."""
    assert target_points.is_contiguous() and query_points.is_contiguous() and query_feats.is_contiguous()
    n, m, c = target_points.size(0), query_points.size(0), query_feats.size(1)
    lib = L.load()
    ws = _workspace(lib.dcl_sp_three_nn_workspace_bytes(n, m), target_points.device)
    L.check(lib.dcl_sp_nn_interpolate_fused_pm(n, m, c, L.ptr(target_points), L.ptr(query_points), L.ptr(query_feats),
                                               L.ptr(out_pm), c_total, out_col0, L.ptr(ws), ws.numel(),
                                               L.stream_ptr()), "pointnet_sp.nn_interpolate_pm")
    return out_pm


def nn_interpolate_vox_pm(target_points, vox_indices, voxel_extent, offset, query_feats, out_pm, c_total, out_col0):
    """nn_interpolate_pm taking the sparse tensor's int32 (m,4) voxel indices (b,ix,iy,iz) plus the level's voxel
    extent / offset (3 floats each) instead of precomputed centres (Ops_tensor2points fused into the kernels)."""
    import ctypes
    assert target_points.is_contiguous() and vox_indices.is_contiguous() and query_feats.is_contiguous()
    L.require(vox_indices, torch.int32, "vox_indices")
    n, m, c = target_points.size(0), vox_indices.size(0), query_feats.size(1)
    lib = L.load()
    ws = _workspace(lib.dcl_sp_three_nn_workspace_bytes(n, m), target_points.device)
    ext = (ctypes.c_float * 3)(*[float(v) for v in voxel_extent])
    off = (ctypes.c_float * 3)(*[float(v) for v in offset])
    L.check(lib.dcl_sp_nn_interpolate_vox_pm(n, m, c, L.ptr(target_points), L.ptr(vox_indices),
                                             ctypes.cast(ext, ctypes.c_void_p), ctypes.cast(off, ctypes.c_void_p),
                                             L.ptr(query_feats), L.ptr(out_pm), c_total, out_col0, L.ptr(ws),
                                             ws.numel(), L.stream_ptr()), "pointnet_sp.nn_interpolate_vox_pm")
    return out_pm


def nn_interpolate_vox_levels_pm(target_points, levels, out_pm, c_total):
    """All pyramid levels of one tower in two launches (one bucket build, one search + interpolation), bit-identical
    to calling nn_interpolate_vox_pm per level.  `levels`: list of (vox_indices (m,4) int32, voxel_extent[3],
    offset[3], feats (m,c) fp32, out_col0[, grid_x]); grid_x = size of the level's voxel grid along the first
    axis (enables the slab walk of the search; 0 / omitted = plain per-instance scan)."""
    nn_interpolate_vox_towers_pm([(target_points, levels, out_pm, c_total)])
    return out_pm


def nn_interpolate_vox_towers_pm(towers, fmt=0):
    """Several towers — [(target_points, levels, out_pm, c_total)], levels as in nn_interpolate_vox_levels_pm, at most 8
    levels in total — in ONE pair of launches: the observed and the template cloud of the network share the bucket
    build launch and the search + interpolation launch.  fmt: 0 = PM image (bf16 hi/lo), 1 = PM16 image (fp16)."""
    import ctypes
    lib = L.load()
    tw = (L.SpTower * len(towers))()
    keep, nbytes = [], 0
    for tslot, (target_points, levels, out_pm, c_total) in zip(tw, towers):
        assert target_points.is_contiguous()
        arr = (L.SpLevel * len(levels))()
        for slot, spec in zip(arr, levels):
            vox, ext, off, feats, col0 = spec[:5]
            slot.grid_x = int(spec[5]) if len(spec) > 5 else 0
            assert vox.is_contiguous() and feats.is_contiguous()
            L.require(vox, torch.int32, "vox_indices")
            L.require(feats, torch.float32, "feats")
            slot.m, slot.c, slot.out_col0 = vox.size(0), feats.size(1), col0
            slot.vox_indices, slot.feats = L.ptr(vox), L.ptr(feats)
            slot.voxel_extent = (ctypes.c_float * 3)(*[float(v) for v in ext])
            slot.offset = (ctypes.c_float * 3)(*[float(v) for v in off])
        arr_p = ctypes.cast(arr, ctypes.c_void_p)
        nbytes += lib.dcl_sp_levels_workspace_bytes(len(levels), arr_p)
        tslot.n, tslot.c_total, tslot.nlevels = target_points.size(0), c_total, len(levels)
        tslot.unknown, tslot.out_pm, tslot.levels = L.ptr(target_points), L.ptr(out_pm), arr_p
        tslot.out_fmt = fmt
        keep.append(arr)
    ws = _workspace(nbytes, towers[0][0].device)
    L.check(lib.dcl_sp_nn_interpolate_towers_pm(len(towers), ctypes.cast(tw, ctypes.c_void_p), L.ptr(ws), ws.numel(),
                                                L.stream_ptr()), "pointnet_sp.nn_interpolate_vox_towers_pm")
