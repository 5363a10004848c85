"""The reference's loss modules and metric without the B x N x M x 3 broadcast.

    losses             models/DCL_Net.py:261-311      (same class name, forward signature and returned keys)
    losses_refiner     models/refiner.py:97-133
    L2_Dis, CD_Dis     models/DCL_Net.py:304-311, models/refiner.py:126-133
    ADD-S metric       tools/test_YCBV_stage1.py:186-188   (`cd_dis` there)

`cd_dis` / `nearest_dist` are differentiable: the forward is the `dcl_nearest_dist` kernel (min distance + argmin per
point), the backward is the gradient of ||p_i - q_j*|| at the recorded nearest neighbour — what autograd gives the
reference through `torch.min(...)[0]` of the broadcast norm.
"""
import torch
import torch.nn as nn

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


def get_cano_label(points_tmp, points_inp, rot_pred, trans_gt):
    """models/DCL_Net.py:312-317: for every canonicalised input point its nearest template point (the reference
    runs knn(1, ...) and a gather; nearest_dist's argmin is the same lowest-index nearest neighbour)."""
    points_inp_cano = torch.bmm((points_inp - trans_gt), rot_pred)
    _, idx = _nearest(points_inp_cano, points_tmp, True)
    return torch.gather(points_tmp, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3))


def adds_metric(points_posed_pred, points_posed_gt):
    """ADD-S of tools/test_YCBV_stage1.py:188: mean over the model points of the distance to the closest GT-posed point."""
    return nearest_dist(points_posed_pred, points_posed_gt).mean(dim=1)


class losses(nn.Module):
    """Stage-1 training loss, models/DCL_Net.py:261-303 line by line; only CD_Dis runs on the kernel."""

    def __init__(self, cfg=None) -> None:
        super().__init__()

    L2_Dis = staticmethod(L2_Dis)
    CD_Dis = staticmethod(CD_Dis)
    get_cano_label = staticmethod(get_cano_label)

    def forward(self, loss_inp_pred, loss_inp_gt):
        rot_pred, trans_pred = loss_inp_pred["rot_pred"], loss_inp_pred["trans_pred"]
        sym_flag = loss_inp_pred["sym_flag"]
        dev = rot_pred.device
        rot_gt, trans_gt = loss_inp_gt["rot_gt"].to(dev), loss_inp_gt["trans_gt"].to(dev)
        points_tmp, points_inp = loss_inp_gt["points_tmp"], loss_inp_gt["points_inp"]
        conf = loss_inp_pred["conf"]

        points_tmp_posed_pred = torch.bmm(points_tmp, rot_pred.transpose(1, 2)) + trans_pred.unsqueeze(1)
        points_tmp_posed_gt = torch.bmm(points_tmp, rot_gt.transpose(1, 2)) + trans_gt.unsqueeze(1)
        asym, sym = (1 - sym_flag).unsqueeze(1), sym_flag.unsqueeze(1)
        loss_pose = (asym * self.L2_Dis(points_tmp_posed_pred, points_tmp_posed_gt)
                     + sym * self.CD_Dis(points_tmp_posed_pred, points_tmp_posed_gt)).mean(dim=1).mean()

        Xo_pred, Yc_pred = loss_inp_pred["Xo_pred"], loss_inp_pred["Yc_pred"]
        points_inp_posed_pred = torch.bmm(points_inp - trans_pred.unsqueeze(1), rot_pred).detach()
        points_inp_posed_gt = torch.bmm(points_inp - trans_gt.unsqueeze(1), rot_gt).detach()
        loss_Xo = asym * self.L2_Dis(Xo_pred, points_inp_posed_gt) + 0.5 * sym * (
            self.CD_Dis(Xo_pred, points_tmp) + self.L2_Dis(Xo_pred, points_inp_posed_pred))
        loss_Xo_ = loss_Xo.mean()
        loss_Yc = asym * self.L2_Dis(Yc_pred, points_tmp_posed_gt) + 0.5 * sym * (
            self.CD_Dis(Yc_pred, points_tmp_posed_gt) + self.L2_Dis(Yc_pred, points_tmp_posed_pred.detach()))
        loss_Yc_ = loss_Yc.mean()
        loss_conf = torch.mean(torch.cat([loss_Xo, loss_Yc], dim=1).detach() * conf - 0.01 * torch.log(conf))
        loss_all = loss_pose + 5 * loss_Xo_ + 1 * loss_Yc_ + 1 * loss_conf
        return {"loss_pose": loss_pose, "loss_Xo": loss_Xo_, "loss_Yc": loss_Yc_, "loss_conf": loss_conf,
                "loss_all": loss_all}


class losses_refiner(nn.Module):
    """Stage-2 training loss, models/refiner.py:97-125."""

    def __init__(self, cfg=None) -> None:
        super().__init__()

    L2_Dis = staticmethod(L2_Dis)
    CD_Dis = staticmethod(CD_Dis)

    def forward(self, loss_inp_pred_refiner, trans_cur, rot_cur, points_tmp, sym_flag, loss_inp_gt):
        delta_rot_pred, delta_trans_pred = loss_inp_pred_refiner["rot_pred"], loss_inp_pred_refiner["trans_pred"]
        dev = delta_rot_pred.device
        rot_gt, trans_gt = loss_inp_gt["rot_gt"].to(dev), loss_inp_gt["trans_gt"].to(dev)
        points_tmp_posed_pred = torch.bmm(points_tmp, delta_rot_pred.transpose(1, 2)) + delta_trans_pred.unsqueeze(1)
        points_tmp_posed_gt = torch.bmm(points_tmp, rot_gt.transpose(1, 2)) + trans_gt.unsqueeze(1)
        points_tmp_posed_refined = torch.bmm(points_tmp_posed_pred, rot_cur.transpose(1, 2)) + trans_cur.unsqueeze(1)
        loss_pose = ((1 - sym_flag).unsqueeze(1) * self.L2_Dis(points_tmp_posed_refined, points_tmp_posed_gt)
                     + sym_flag.unsqueeze(1) * self.CD_Dis(points_tmp_posed_refined, points_tmp_posed_gt)).mean(dim=1).mean()
        return {"loss_pose": loss_pose, "loss_all": loss_pose}
