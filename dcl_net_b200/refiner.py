"""Drop-in for the reference's models/refiner.py and the stage-2 loop that drives it.

    Refiner(cfg).forward({"input_features","conf","obj_idx"}) -> {"trans_pred","rot_pred"}   refiner.py:57-95
    refine_poses(...)   the iteration of tools/test_YCBV_stage2.py:204-225 with the pose composition
                        and re-canonicalisation fused into one kernel (dcl_pose_compose)
"""
import torch
import torch.nn as nn

from . import _lib as L
from .dcl_net import ortho9d2matrix  # noqa: F401  (models/refiner.py:35-56 duplicates models/DCL_Net.py:15-36)
from .modules import Head_MultiLayerPerceptron


class Refiner(nn.Module):
    def __init__(self, cfg=None) -> None:
        super().__init__()
        plain = ([False] * 3, [0.0] * 3)
        self.MLP_share = Head_MultiLayerPerceptron([256 + 3, 512, 512, 1024], ["relu"] * 3, *plain)
        self.regressor_rot2 = Head_MultiLayerPerceptron([1024, 512, 128, 9], ["relu", "relu", "none"], *plain)
        self.regressor_trans2 = Head_MultiLayerPerceptron([1024, 512, 128, 3], ["relu", "relu", "none"], *plain)
        self.use_fused = True          # inference: shared MLP on tensor cores inside refine_poses (fused_tail.FusedRefiner)
        self.precision = "fp16"        # operand format when the caller brings no packed feature image (see dcl_net.Network)
        self._fused_refiner = None

    def _apply(self, fn, *args, **kwargs):
        self._fused_refiner = None     # parameters moved / cast: packed copies are stale
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._fused_refiner = None
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode=True):
        if mode != self.training:
            self._fused_refiner = None
        return super().train(mode)

    def _mlp_share(self, x):
        """MLP_share (Conv1d 259 -> 512 -> 512 -> 1024, ReLU after each; models/refiner.py:61-63).  In training on CUDA it
        runs — forward and backward — on the tensor-core training kernels (train_tail.mlp_stacks), like the stage-1
        stacks: the 3 coordinate channels are zero-padded to a 32-channel block (input and the matching weight
        columns, by differentiable torch ops), the 256 feature channels follow as a second input."""
        b, c, n = x.shape
        if not (self.training and torch.is_grad_enabled() and getattr(self, "use_train_kernels", True) and x.is_cuda
                and x.dtype == torch.float32 and c == 259 and n % 128 == 0):
            return self.MLP_share(x)
        from .train_tail import StackSpec, head_layers, mlp_stacks
        lays, _ = head_layers(self.MLP_share)
        kind, w, bias, bn, relu_mod = lays[0]
        w = torch.cat([w[:, :3], w.new_zeros(w.shape[0], 29), w[:, 3:]], dim=1)          # (512, 288)
        pts = torch.nn.functional.pad(x[:, :3], (0, 0, 0, 29))                            # (b, 32, n)
        out, = mlp_stacks([StackSpec([(pts, "cm"), (x[:, 3:].contiguous(), "cm")],
                                     [(kind, w, bias, bn, relu_mod)] + lays[1:])], b, n)
        return out

    def forward(self, input_dict):
        input_features = input_dict["input_features"]
        conf = input_dict["conf"]
        conf_softmax = torch.softmax(conf.unsqueeze(1), dim=2)[:, :, :1024]
        shared_feature = self._mlp_share(input_features)
        shared_feature = (shared_feature * conf_softmax).sum(dim=2, keepdim=True)
        if torch.is_grad_enabled():
            ortho9d_pred2 = self.regressor_rot2(shared_feature).squeeze(-1)
            delta_t = self.regressor_trans2(shared_feature).squeeze(-1)
        else:
            from .dcl_net import pose_heads
            ortho9d_pred2, delta_t = pose_heads(shared_feature.squeeze(-1), self.regressor_rot2, self.regressor_trans2)
        # differentiable (torch.svd formula) when autograd records a gradient through it, as the reference's
        # losses_refiner needs (models/refiner.py:35-56,97-125); the dcl_svd3_project kernel otherwise
        delta_R = ortho9d2matrix(ortho9d_pred2[:, :3], ortho9d_pred2[:, 3:6], ortho9d_pred2[:, 6:])
        return {"trans_pred": delta_t, "rot_pred": delta_R}


def pose_compose_(R, t, dR, dt, points_in, out_cm):
    """In place: t <- R dt + t, R <- R dR (skipped when dR is None); then writes the canonicalised cloud
    (points_in - t) R channel-major into out_cm[:, :3, :] (out_cm is (B, >=3, N), contiguous)."""
    B, N = points_in.shape[0], points_in.shape[1]
    assert out_cm.is_contiguous() and out_cm.shape[0] == B and out_cm.shape[2] == N and out_cm.shape[1] >= 3
    L.check(L.load().dcl_pose_compose(B, N, L.ptr(R), L.ptr(t), L.ptr(dR), L.ptr(dt), L.ptr(points_in),
                                      L.ptr(out_cm), out_cm.shape[1] * N, L.stream_ptr()), "pose_compose")


def refine_poses(refiner, points_inp, rot_pred, trans_pred, F_Xo_p, conf, iteration=2, F_Xo_p_pm=None, pm_fmt=0):
    """Stage-2 iterative refinement.  points_inp (B,N,3), rot_pred (B,3,3), trans_pred (B,3),
    F_Xo_p (B,256,N), conf (B,2N) -> refined (rot, trans).  Inference runs the refiner's shared MLP on tensor
    cores (fused_tail.FusedRefiner; F_Xo_p_pm = the point-major image of F_Xo_p if the caller already has it, in
    format pm_fmt — stage 1 returns both; without an image the refiner's own `precision` picks the format);
    otherwise the refiner input buffer (B,259,N) is built once and each iteration only rewrites its first three
    channels."""
    B, N, _ = points_inp.shape
    from .fused_tail import FusedRefiner
    if getattr(refiner, "use_fused", True) and FusedRefiner.supported(refiner, B, N, F_Xo_p, conf, F_Xo_p_pm):
        if F_Xo_p_pm is None:
            pm_fmt = L.FMT_F16 if getattr(refiner, "precision", "fp16") == "fp16" else L.FMT_BF16X2
        fused = getattr(refiner, "_fused_refiner", None)
        if fused is None or fused.fmt != pm_fmt:
            fused = refiner._fused_refiner = FusedRefiner(refiner, pm_fmt)
        return fused.refine(points_inp, rot_pred, trans_pred, F_Xo_p, conf, iteration, F_Xo_p_pm)
    points_inp = points_inp.contiguous()
    rot_cur, trans_cur = rot_pred.clone().contiguous(), trans_pred.clone().contiguous()
    inp_refiner = torch.empty(B, 3 + F_Xo_p.shape[1], N, dtype=torch.float32, device=points_inp.device)
    inp_refiner[:, 3:, :] = F_Xo_p
    pose_compose_(rot_cur, trans_cur, None, None, points_inp, inp_refiner)
    for _ in range(iteration):
        out = refiner({"input_features": inp_refiner, "conf": conf, "obj_idx": None})
        pose_compose_(rot_cur, trans_cur, out["rot_pred"].contiguous(), out["trans_pred"].contiguous(),
                      points_inp, inp_refiner)
    return rot_cur, trans_cur
