"""Drop-ins for the hot-path pieces of the reference's models/Modules.py.

    Aligner                          :162-169   fused tcgen05 kernel (dcl_fda_align_fwd)
    Head_MultiLayerPerceptron        :173-201   same layers / parameter names (library GEMMs)
    BasicBlock_3DCONV                :58-97     same layers / parameter names
    Ops_tensor2points                :204-211
    Ops_nearest_neighbor_interpolate :213-226   pointnet_sp kernels (fused in inference)
    Ops_GetPointFeat_spconv          :227-251
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .pointnet_sp import pointnet2_utils as pointnet2_utils_sp

_fda_ws = {}
# bench.py sets this to a list to collect (start, end) CUDA events around every launch of the fused kernel.
FDA_KERNEL_EVENTS = None


def _fda_workspace(nbytes, device):
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=device)   # private to the graph being captured
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _fda_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _fda_ws[key] = ws
    return ws


class FdaAlignFunction(torch.autograd.Function):
    """Autograd wrapper of the fused kernels (training path).  Forward = the fused tcgen05 kernel, saving the inputs,
    the outputs and the per-query log-sum-exp; backward = dcl_fda_bwd (csrc/fda_bwd.cu), which rebuilds
    A = exp(S - lse) tile by tile on chip (the reference keeps A (B,M,N) alive instead, models/Modules.py:167-169)
    and forms the five gradient products on tensor cores:
            dA = RE_2^T gE + RI_2^T gI;  dS = A o (dA - sum_m A o dA);
            dRE_2 = gE A^T;  dRI_2 = gI A^T + RI_1 dS^T;  dRI_1 = RI_2 dS."""

    @staticmethod
    def forward(ctx, RI_1, RI_2, RE_2):
        RE_embed, RI_embed, lse = _fda_align_kernel(RI_1, RI_2, RE_2, True)
        ctx.save_for_backward(RI_1, RI_2, RE_2, lse, RE_embed, RI_embed)
        ctx.mark_non_differentiable(lse)
        return RE_embed, RI_embed, lse

    @staticmethod
    def backward(ctx, gE, gI, _g_lse):
        RI_1, RI_2, RE_2, lse, RE_embed, RI_embed = ctx.saved_tensors
        gE = torch.zeros_like(RE_embed) if gE is None else gE.contiguous()
        gI = torch.zeros_like(RI_embed) if gI is None else gI.contiguous()
        if RI_2.shape[2] % 128 == 0 and USE_FUSED_FDA_BACKWARD:
            return fda_backward(RI_1, RI_2, RE_2, lse, RE_embed, RI_embed, gE, gI)
        L.warn_once("fda_bwd", "dcl_net_b200: FDA backward on library GEMMs (needs m % 128 == 0 for csrc/fda_bwd.cu)")
        return _fda_backward_unfused(RI_1, RI_2, RE_2, lse, gE, gI)


USE_FUSED_FDA_BACKWARD = True     # False: the unfused backward below (A/B runs, tests)


def _fda_backward_unfused(RI_1, RI_2, RE_2, lse, gE, gI):
    """The same gradients with A (B,M,N) materialised in fp32 and library GEMMs."""
    A = fda_attention_map(RI_1, RI_2, lse)                       # (B, M, N)
    # The forward's lse comes out of the split-bf16 logits; the fp32 logits recomputed here differ from those by
    # ~|S| 2^-17, a common factor per query column that the renormalisation removes exactly.
    A = A / A.sum(dim=1, keepdim=True)
    dA = torch.bmm(RE_2.transpose(1, 2), gE) + torch.bmm(RI_2.transpose(1, 2), gI)
    dS = A * (dA - (A * dA).sum(dim=1, keepdim=True))
    d_RE_2 = torch.bmm(gE, A.transpose(1, 2))
    d_RI_2 = torch.bmm(gI, A.transpose(1, 2)) + torch.bmm(RI_1, dS.transpose(1, 2))
    d_RI_1 = torch.bmm(RI_2, dS)
    return d_RI_1, d_RI_2, d_RE_2


def fda_backward(RI_1, RI_2, RE_2, lse, RE_embed, RI_embed, gE, gI):
    """(d RI_1, d RI_2, d RE_2) through dcl_fda_bwd.  Operand images are written by one dcl_tr_tile_pass launch."""
    import ctypes
    from .train_tail import TR_COPY, tile_pass
    B, C, N = RI_1.shape
    M, P = RI_2.shape[2], RE_2.shape[1]
    VC = P + C
    dev = RI_1.device
    u8 = lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=dev)
    cpad, vpad = (C + 127) // 128 * 128, (VC + 127) // 128 * 128
    q_k, k_k = u8(B * N * C * 4), u8(B * M * C * 4)
    v_k, g_k = u8(B * M * VC * 4), u8(B * N * VC * 4)
    q_t, k_t, g_t = u8(B * cpad * N * 4), u8(B * cpad * M * 4), u8(B * vpad * N * 4)
    dsum = (gE * RE_embed).sum(dim=1) + (gI * RI_embed).sum(dim=1)          # D (B, N)
    item = lambda x, n, **kw: dict({"x": x, "fmt": "cm", "b": B, "c": x.shape[1], "n": n, "mode": TR_COPY}, **kw)
    tile_pass([item(RI_1, N, out_k=q_k, out_t=q_t, t_rows=cpad),
               item(RI_2, M, out_k=k_k, out_t=k_t, t_rows=cpad),
               item(RI_2, M, out_k=v_k, k_col0=P, k_cols=VC),
               item(RE_2, M, out_k=v_k, k_col0=0, k_cols=VC),
               item(gE, N, out_k=g_k, k_col0=0, k_cols=VC, out_t=g_t, t_row0=0, t_rows=vpad),
               item(gI, N, out_k=g_k, k_col0=P, k_cols=VC, out_t=g_t, t_row0=P, t_rows=vpad)])
    d_q = torch.empty(B, C, N, dtype=torch.float32, device=dev)
    d_k = torch.empty(B, C, M, dtype=torch.float32, device=dev)
    d_v = torch.empty(B, VC, M, dtype=torch.float32, device=dev)
    job = (L.FdaBwdJob * 1)()
    for name, t in (("q_k", q_k), ("k_k", k_k), ("g_k", g_k), ("v_k", v_k), ("q_t", q_t), ("k_t", k_t), ("g_t", g_t),
                    ("lse", lse.contiguous()), ("dsum", dsum.contiguous()), ("d_q", d_q), ("d_k", d_k), ("d_v", d_v)):
        setattr(job[0], name, L.ptr(t))
    L.check(L.load().dcl_fda_bwd(1, ctypes.cast(job, ctypes.c_void_p), B, C, P, N, M, L.stream_ptr()), "fda_bwd")
    return d_q, d_k + d_v[:, P:], d_v[:, :P]


def fda_align(RI_1, RI_2, RE_2, return_lse=False):
    """One FDA direction, fused:  A = softmax_m(RI_2^T RI_1) is never materialised.

    RI_1 (B,C,N) queries, RI_2 (B,C,M) keys, RE_2 (B,P,M) values, fp32, C in {64,128}, P=256,
    N % 128 == 0, M % 64 == 0.  Returns RE_embed = RE_2 A (B,P,N) and RI_embed = RI_2 A (B,C,N)
    (models/Modules.py:167-168 and models/DCL_Net.py:213/215), optionally the per-query
    log-sum-exp (B,N).  Differentiable (FdaAlignFunction) when autograd is recording.
    """
    if torch.is_grad_enabled() and (RI_1.requires_grad or RI_2.requires_grad or RE_2.requires_grad):
        out = FdaAlignFunction.apply(RI_1.contiguous(), RI_2.contiguous(), RE_2.contiguous())
        return out if return_lse else out[:2]
    return _fda_align_kernel(RI_1, RI_2, RE_2, return_lse)


def _fda_align_kernel(RI_1, RI_2, RE_2, return_lse=False):
    RE_embed, RI_embed, _, _, lse = fda_align_formats(RI_1, RI_2, RE_2, return_lse=return_lse)
    return (RE_embed, RI_embed, lse) if return_lse else (RE_embed, RI_embed)


def fda_align_formats(RI_1, RI_2, RE_2, re_cm=True, ri_cm=True, re_pm=False, ri_pm=False, return_lse=False, pv_fmt=0):
    """fda_align (inference only) with a choice of output formats: `*_cm` = the reference's fp32 (B,ch,N) tensors,
    `*_pm` = point-major bf16 hi/lo images over the B*N query rows (fused_tail.pm_unpack restores (B*N, ch)), which
    the tensor-core MLPs that consume the aligned features (models/DCL_Net.py:216-228) read without a repack.
    pv_fmt 1: the P V products on once-rounded fp16 operands (1 MMA instead of 3; needs (N/128) even), `*_pm` outputs
    as PM16 images.  Returns (RE_cm, RI_cm, RE_pm, RI_pm, lse) with None for the formats not requested."""
    RI_1 = L.require(RI_1.contiguous(), torch.float32, "RI_1")
    RI_2 = L.require(RI_2.contiguous(), torch.float32, "RI_2")
    RE_2 = L.require(RE_2.contiguous(), torch.float32, "RE_2")
    B, C, N = RI_1.shape
    M, P = RI_2.shape[2], RE_2.shape[1]
    if RI_2.shape[:2] != (B, C) or RE_2.shape[0] != B or RE_2.shape[2] != M:
        raise ValueError("fda_align: inconsistent shapes")
    lib = L.load()
    nbytes = lib.dcl_fda_workspace_bytes(B, C, P, N, M)
    if nbytes == 0 or N % 128 or M % 64:
        raise ValueError(f"fda_align: unsupported shape C={C} P={P} N={N} M={M} "
                         "(need C in {64,128}, P=256, N%128==0, M%64==0)")
    ws = _fda_workspace(nbytes, RI_1.device)
    st = L.stream_ptr()
    L.check(lib.dcl_fda_pack_fmt(B, C, P, N, M, L.ptr(RI_1), L.ptr(RI_2), L.ptr(RE_2), L.ptr(ws), ws.numel(), pv_fmt, st),
            "fda_align (pack)")
    return fda_from_workspace(ws, B, C, N, M, re_cm, ri_cm, re_pm, ri_pm, return_lse, pv_fmt)


def fda_from_workspace(ws, B, C, N, M, re_cm=True, ri_cm=True, re_pm=False, ri_pm=False, return_lse=False, pv_fmt=0):
    """The fused kernel on operand images that already sit in `ws` (written by dcl_fda_pack, or directly by the
    disengage GEMMs' epilogue: fused_tail.py).  Same outputs as fda_align_formats."""
    return fda_from_workspaces([(ws, re_cm, ri_cm, re_pm, ri_pm, return_lse)], B, C, N, M, pv_fmt)[0]


def fda_from_workspaces(jobs, B, C, N, M, pv_fmt=0):
    """Up to two independent FDA problems of equal shape in ONE launch (the two directions of the dual FDA share
    their partial last waves).  jobs: [(workspace, re_cm, ri_cm, re_pm, ri_pm, return_lse)]; returns one
    (RE_cm, RI_cm, RE_pm, RI_pm, lse) tuple per job."""
    import ctypes
    P = 256
    lib = L.load()
    dev = jobs[0][0].device
    f32, u8 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.uint8, device=dev)
    arr = (L.FdaJob * len(jobs))()
    results = []
    esz = 2 if pv_fmt == 1 else 4     # bytes per element of a point-major output image
    for slot, (ws, re_cm, ri_cm, re_pm, ri_pm, want_lse) in zip(arr, jobs):
        out = (torch.empty(B, P, N, **f32) if re_cm else None, torch.empty(B, C, N, **f32) if ri_cm else None,
               torch.empty(B * N * P * esz, **u8) if re_pm else None, torch.empty(B * N * C * esz, **u8) if ri_pm else None,
               torch.empty(B, N, **f32) if want_lse else None)
        slot.workspace = L.ptr(ws)
        slot.RE_embed, slot.RI_embed, slot.RE_pm, slot.RI_pm, slot.lse = (L.ptr(x) for x in out)
        results.append(out)
    nbytes = min(j[0].numel() for j in jobs)
    if FDA_KERNEL_EVENTS is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(lib.dcl_fda_fwd_packed_jobs_fmt(len(jobs), ctypes.cast(arr, ctypes.c_void_p), B, C, P, N, M, nbytes, pv_fmt,
                                            L.stream_ptr()), "fda_align")
    if FDA_KERNEL_EVENTS is not None:
        ev1.record()
        FDA_KERNEL_EVENTS.append((ev0, ev1, len(jobs)))
    return results


def fda_attention_map(RI_1, RI_2, lse):
    """A (B,M,N) = exp(RI_2^T RI_1 - lse): the reference's `attention_map` (Modules.py:167)."""
    RI_1, RI_2, lse = RI_1.contiguous(), RI_2.contiguous(), lse.contiguous()
    B, C, N = RI_1.shape
    M = RI_2.shape[2]
    A = torch.empty(B, M, N, dtype=torch.float32, device=RI_1.device)
    L.check(L.load().dcl_fda_attention_map(B, C, N, M, L.ptr(RI_1), L.ptr(RI_2), L.ptr(lse), L.ptr(A),
                                           L.stream_ptr()), "fda_attention_map")
    return A


class Aligner(nn.Module):
    """forward(RI_1, RI_2, RE_2) -> (RE_embed, attention_map), as models/Modules.py:162-169.

    RE_embed comes from the fused kernel.  attention_map is what the reference returns second;
    it is rebuilt from the kernel's log-sum-exp only because this signature demands it —
    Network.forward uses `fda_align` directly and never materialises it.
    """

    def forward(self, RI_1, RI_2, RE_2):
        RE_embed, _, lse = fda_align(RI_1, RI_2, RE_2, return_lse=True)
        return RE_embed, fda_attention_map(RI_1, RI_2, lse)


class BasicBlock_3DCONV(nn.Module):
    def __init__(self, dim_in, dim_out, bias, size, stride, padding, norm, act, drop):
        super().__init__()
        layers = [nn.Conv3d(dim_in, dim_out, size, stride, padding, bias=bias)]
        if norm:
            layers.append(nn.BatchNorm3d(dim_out))
        layers += _activation(act)
        if drop > 0:
            layers.append(nn.Dropout(drop))
        self.layers = nn.Sequential(*layers)

    def forward(self, input):
        return self.layers(input)


def _activation(act):
    if act == "relu":
        return [nn.ReLU()]
    if act == "sigmoid":
        return [nn.Sigmoid()]
    if act == "tanh":
        return [nn.Tanh()]
    if act == "none":
        return []
    raise NotImplementedError(act)


class Head_MultiLayerPerceptron(nn.Module):
    """Conv1d(k=1) -> activation -> [BatchNorm1d] -> [Dropout] per layer (BN after the activation)."""

    def __init__(self, list_dim, list_act, list_bn, list_drop):
        super().__init__()
        layers, dim_inp = [], list_dim[0]
        for dim, act, bn, drop in zip(list_dim[1:], list_act, list_bn, list_drop):
            layers.append(nn.Conv1d(dim_inp, dim, 1))
            layers += _activation(act)
            if bn:
                layers.append(nn.BatchNorm1d(dim))
            if drop > 0.0:
                layers.append(nn.Dropout(drop))
            dim_inp = dim
        self.layers = nn.Sequential(*layers)

    def forward(self, input):
        return self.layers(input)


_const_cache = {}


def _const_tensor(values, device):
    """fp32 device tensor of a small host constant, created once per (values, device): a host-to-device copy per
    call would also keep a training step from being captured in a CUDA graph."""
    key = (tuple(float(v) for v in np.asarray(values).reshape(-1)), str(device))
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.tensor(key[0], dtype=torch.float32, device=device)
    return t


def Ops_tensor2points(tensor, offset=(0., -40., -3.), voxel_extent=(.1, .1, .2)):
    """Sparse tensor (.features (Mv,C), .indices (Mv,4) int bxyz) -> (features, voxel centres bxyz)."""
    indices = tensor.indices.float()
    offset = _const_tensor(offset, indices.device)
    voxel_extent = _const_tensor(voxel_extent, indices.device)
    indices[:, 1:] = indices[:, 1:] * voxel_extent + offset + .5 * voxel_extent
    return tensor.features, indices


def Ops_nearest_neighbor_interpolate(target_points, query_points, query_feats):
    """(n,4) bxyz targets, (m,4) bxyz sources, (m,C) features -> (n,C)."""
    if not (torch.is_grad_enabled() and query_feats.requires_grad):
        return pointnet2_utils_sp.nn_interpolate(target_points.contiguous(), query_points.contiguous(),
                                                 query_feats.contiguous())
    dist, idx = pointnet2_utils_sp.three_nn(target_points, query_points)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=1, keepdim=True)
    weight = dist_recip / norm
    return pointnet2_utils_sp.three_interpolate(query_feats, idx, weight)


class Ops_GetPointFeat_spconv(nn.Module):
    def __init__(self, scale_lists=[2, 4, 8, 16], unit_voxel_extent=np.array([0.015, 0.015, 0.015]),
                 voxel_num_limit=np.array([64, 64, 64])):
        super().__init__()
        self.scale_lists = scale_lists
        self.unit_voxel_extent = np.asarray(unit_voxel_extent)
        self.voxel_num_limit = np.asarray(voxel_num_limit)
        self.offset = -0.5 * self.unit_voxel_extent * self.voxel_num_limit

    def _pm_tower(self, points, batch_ids, levels, fmt=0):
        points = torch.cat([batch_ids.view(-1, 1).float(), points], 1).contiguous()
        width = sum(f.features.shape[1] for f in levels)
        out = torch.empty(points.shape[0] * width * (2 if fmt == 1 else 4), dtype=torch.uint8, device=points.device)
        # Ops_tensor2points is fused into the kernels: they take the int voxel indices and form the centres
        # ((i * ext) + offset) + 0.5 * ext in fp32, in torch's evaluation order.
        off = torch.as_tensor(np.asarray(self.offset), dtype=torch.float32).tolist()
        specs, col = [], 0
        for scale, feats in zip(self.scale_lists, levels):
            ext = torch.as_tensor(np.asarray(self.unit_voxel_extent * scale), dtype=torch.float32).tolist()
            grid_x = int(np.ceil(self.voxel_num_limit[0] / scale))   # voxel indices of this level lie in [0, grid_x)
            specs.append((feats.indices.contiguous(), ext, off, feats.features.contiguous(), col, grid_x))
            col += feats.features.shape[1]
        return points, specs, out, width

    def forward_pm(self, points, batch_ids, feats1, feats2, feats3, feats4):
        """Inference only: the concatenated (n, 480) point features as a PM image (operand of the tensor-core
        disengage GEMMs) instead of an fp32 matrix.  All four levels in one call (one bucket-build launch, one
        search + interpolation launch)."""
        tower = self._pm_tower(points, batch_ids, [feats1, feats2, feats3, feats4])
        pointnet2_utils_sp.nn_interpolate_vox_towers_pm([tower])
        return tower[2]

    def forward_pm_pair(self, points_a, ids_a, levels_a, points_b, ids_b, levels_b, fmt=0):
        """forward_pm for the observed and the template cloud together: the two towers share both launches.
        fmt 1 writes PM16 images (fp16), the activation format of the fp16 inference path."""
        ta = self._pm_tower(points_a, ids_a, list(levels_a), fmt)
        tb = self._pm_tower(points_b, ids_b, list(levels_b), fmt)
        pointnet2_utils_sp.nn_interpolate_vox_towers_pm([ta, tb], fmt)
        return ta[2], tb[2]

    def forward(self, points, batch_ids, feats1, feats2, feats3, feats4):
        points = torch.cat([batch_ids.view(-1, 1).float(), points], 1).contiguous()
        levels = [feats1, feats2, feats3, feats4]
        fused = not (torch.is_grad_enabled() and any(f.features.requires_grad for f in levels))
        if fused:
            width = sum(f.features.shape[1] for f in levels)
            out = torch.empty(points.shape[0], width, dtype=torch.float32, device=points.device)
            col = 0
            for scale, feats in zip(self.scale_lists, levels):
                vx_feats, vx_points = Ops_tensor2points(feats, self.offset, self.unit_voxel_extent * scale)
                pointnet2_utils_sp.nn_interpolate(points, vx_points.contiguous(), vx_feats.contiguous(), out, col)
                col += vx_feats.shape[1]
            return out
        outs = []
        for scale, feats in zip(self.scale_lists, levels):
            vx_feats, vx_points = Ops_tensor2points(feats, self.offset, self.unit_voxel_extent * scale)
            outs.append(Ops_nearest_neighbor_interpolate(points, vx_points, vx_feats))
        return torch.cat(outs, dim=1)
