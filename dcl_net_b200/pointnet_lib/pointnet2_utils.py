"""Drop-in for the reference's libs/pointnet_lib/pointnet2_utils.py.

Same names, argument order, return dtypes and backward contracts (index ops return None
gradients; gather / grouping / three_interpolate return the gradient w.r.t. `features` only):
    furthest_point_sample  :10-37     gather_operation    :40-76     knn               :78-108
    three_nn               :110-141   three_interpolate   :144-192   grouping_operation :195-238
    ball_query             :241-271   QueryAndGroup :274-307   GroupAll :310-333   KNNAndGroup :335-383
Every op runs the sm_100a kernels of libdcl_b200.so on the current CUDA stream; outputs and
scratch are allocated here, by the caller, exactly as the reference does.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib as L


def _f32(t, what):
    return L.require(t.contiguous(), torch.float32, what)


def _i32(t, what):
    return L.require(t.contiguous(), torch.int32, what)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B,N,3) -> (B,npoint) int32 indices; first index is always 0."""
        xyz = _f32(xyz, "xyz")
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        L.check(L.load().dcl_lib_furthest_point_sampling_kernel_launcher(
            B, N, npoint, L.ptr(xyz), L.ptr(temp), L.ptr(output), L.stream_ptr()), "furthest_point_sample")
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""
        features, idx = _f32(features, "features"), _i32(idx, "idx")
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, dtype=torch.float32, device=features.device)
        L.check(L.load().dcl_lib_gather_points_kernel_launcher_fast(
            B, C, N, npoint, L.ptr(features), L.ptr(idx), L.ptr(output), L.stream_ptr()), "gather_operation")
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        grad_out_data = _f32(grad_out.data, "grad_out")
        L.check(L.load().dcl_lib_gather_points_grad_kernel_launcher_fast(
            B, C, N, npoint, L.ptr(grad_out_data), L.ptr(idx), L.ptr(grad_features), L.stream_ptr()),
            "gather_operation backward")
        return grad_features, None


gather_operation = GatherOperation.apply


class KNN(Function):
    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B,N,3), known (B,M,3) -> dist (B,N,k) L2 distances ascending, idx (B,N,k) int32."""
        if not 1 <= k <= 200:
            raise ValueError("knn: k must be in [1, 200] (the reference's fixed per-thread list size)")
        unknown, known = _f32(unknown, "unknown"), _f32(known, "known")
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, k, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, k, dtype=torch.int32, device=unknown.device)
        L.check(L.load().dcl_lib_knn_kernel_launcher_fast(
            B, N, m, k, L.ptr(unknown), L.ptr(known), L.ptr(dist2), L.ptr(idx), L.stream_ptr()), "knn")
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B,N,3), known (B,M,3) -> dist (B,N,3), idx (B,N,3) int32."""
        unknown, known = _f32(unknown, "unknown"), _f32(known, "known")
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
        L.check(L.load().dcl_lib_three_nn_kernel_launcher_fast(
            B, N, m, L.ptr(unknown), L.ptr(known), L.ptr(dist2), L.ptr(idx), L.stream_ptr()), "three_nn")
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B,C,M), idx (B,n,3), weight (B,n,3) -> (B,C,n)."""
        features, idx, weight = _f32(features, "features"), _i32(idx, "idx"), _f32(weight, "weight")
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        L.check(L.load().dcl_lib_three_interpolate_kernel_launcher_fast(
            B, c, m, n, L.ptr(features), L.ptr(idx), L.ptr(weight), L.ptr(output), L.stream_ptr()),
            "three_interpolate")
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        grad_out_data = _f32(grad_out.data, "grad_out")
        L.check(L.load().dcl_lib_three_interpolate_grad_kernel_launcher_fast(
            B, c, n, m, L.ptr(grad_out_data), L.ptr(idx), L.ptr(weight), L.ptr(grad_features), L.stream_ptr()),
            "three_interpolate backward")
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""
        features = _f32(features, "features")
        idx = idx.contiguous().int()  # the reference casts here too (:210)
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, nfeatures, nsample, dtype=torch.float32, device=features.device)
        L.check(L.load().dcl_lib_group_points_kernel_launcher_fast(
            B, C, N, nfeatures, nsample, L.ptr(features), L.ptr(idx), L.ptr(output), L.stream_ptr()),
            "grouping_operation")
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        grad_out_data = _f32(grad_out.data, "grad_out")
        L.check(L.load().dcl_lib_group_points_grad_kernel_launcher_fast(
            B, C, N, npoint, nsample, L.ptr(grad_out_data), L.ptr(idx), L.ptr(grad_features), L.stream_ptr()),
            "grouping_operation backward")
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3) -> idx (B,npoint,nsample) int32, zero rows where the ball is empty."""
        new_xyz, xyz = _f32(new_xyz, "new_xyz"), _f32(xyz, "xyz")
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        L.check(L.load().dcl_lib_ball_query_kernel_launcher_fast(
            B, N, npoint, float(radius), nsample, L.ptr(new_xyz), L.ptr(xyz), L.ptr(idx), L.stream_ptr()),
            "ball_query")
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping; (B, C+3, npoint, nsample) with features first, as the reference (:274-307)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_features, grouped_xyz], dim=1) if self.use_xyz else grouped_features


class GroupAll(nn.Module):
    """(B, 3+C, 1, N) with xyz first, as the reference (:310-333)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class KNNAndGroup(nn.Module):
    """k-NN + grouping (:335-383).  The reference calls knn(xyz, new_xyz, radius, nsample), which does
    not match knn's own signature (k, unknown, known) and cannot run; the evident intent
    — the nsample nearest points of xyz around every new_xyz centre — is what this computes when no
    precomputed idx is passed."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz=None, idx=None, features=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            _, idx = knn(self.nsample, new_xyz, xyz)
        idx = idx.detach()
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz -= new_xyz.transpose(1, 2).contiguous().unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
