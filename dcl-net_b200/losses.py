"""Point-set distances of the reference's loss / metric code without the B x N x M x 3 broadcast.

    L2_Dis, CD_Dis     models/DCL_Net.py:304-311, models/refiner.py:126-133
    ADD-S metric       tools/test_YCBV_stage1.py:186-188   (`cd_dis` there)

`cd_dis` / `nearest_dist` are differentiable: the forward is the `dcl_nearest_dist` kernel (min distance + argmin per
point), the backward is the gradient of ||p_i - q_j*|| at the recorded nearest neighbour — what autograd gives the
reference through `torch.min(...)[0]` of the broadcast norm.
"""
import torch

from . import _lib as L


def _nearest(a, b, want_idx):
    a = L.require(a.contiguous(), torch.float32, "points a")
    b = L.require(b.contiguous(), torch.float32, "points b")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 3 or b.shape[2] != 3 or a.shape[0] != b.shape[0]:
        raise ValueError("nearest_dist: expected (B,N,3) and (B,M,3)")
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    if M == 0:
        raise ValueError("nearest_dist: empty target cloud")
    dist = torch.empty(B, N, dtype=torch.float32, device=a.device)
    idx = torch.empty(B, N, dtype=torch.int32, device=a.device) if want_idx else None
    L.check(L.load().dcl_nearest_dist(B, N, M, L.ptr(a), L.ptr(b), L.ptr(dist), L.ptr(idx), L.stream_ptr()),
            "nearest_dist")
    return dist, idx


class NearestDistFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        dist, idx = _nearest(a, b, True)
        ctx.save_for_backward(a, b, dist, idx)
        return dist

    @staticmethod
    def backward(ctx, g):
        a, b, dist, idx = ctx.saved_tensors
        j = idx.long().unsqueeze(-1).expand(-1, -1, 3)
        diff = a - torch.gather(b, 1, j)
        # d||x||/dx = x / ||x||; the reference's torch.norm has a zero subgradient at zero distance
        unit = torch.where(dist.unsqueeze(-1) > 0, diff / dist.unsqueeze(-1).clamp_min(1e-38), torch.zeros_like(diff))
        ga = g.unsqueeze(-1) * unit
        gb = torch.zeros_like(b).scatter_add_(1, j, -ga)
        return ga, gb


def nearest_dist(a, b):
    """(B,N,3), (B,M,3) -> (B,N): min_j ||a_i - b_j||  ==  torch.min(torch.norm(a[:,:,None] - b[:,None], dim=3), 2)[0]."""
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return NearestDistFunction.apply(a.contiguous(), b.contiguous())
    return _nearest(a, b, False)[0]


def L2_Dis(pred, target):
    return torch.norm(pred - target, dim=2)


def CD_Dis(pred, target):
    """0.5 * (min over targets + min over predictions), as the reference (which needs N == M for the sum)."""
    return 0.5 * (nearest_dist(pred, target) + nearest_dist(target, pred))


def adds_metric(points_posed_pred, points_posed_gt):
    """ADD-S of tools/test_YCBV_stage1.py:188: mean over the model points of the distance to the closest GT-posed point."""
    return nearest_dist(points_posed_pred, points_posed_gt).mean(dim=1)
