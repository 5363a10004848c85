"""Drop-in for the hot-path part of the reference's models/DCL_Net.py.

    ortho9d2matrix(x_raw, y_raw, z_raw) -> (B,3,3)         :15-36   dcl_svd3_project kernel
    Network(cfg, mode).forward(data) -> dict                :38-259  same modules / parameter names

Entry points: `forward(data)` (the reference's interface, from the dataloader's per-point tensors: voxelisation and
the two sparse-conv towers run on the device, backbone.py — inference only; needs with_backbone=True or an injected
`backbone` callable), `forward_from_points` (raw clouds + colours), `forward_from_backbone` (pyramid levels -> pose:
point-feature interpolation, FDA, pose) and `forward_from_point_feats` (FDA + pose).
"""
import numpy as np
import torch
import torch.nn as nn
from functools import partial

from . import _lib as L
from .modules import (BasicBlock_3DCONV, Head_MultiLayerPerceptron, Ops_GetPointFeat_spconv, fda_align,
                      fda_attention_map)


def svd3_project(m9, normalize_columns):
    """(B,9) fp32 -> (B,3,3) rotation; see include/dcl_b200.h:dcl_svd3_project."""
    m9 = L.require(m9.contiguous(), torch.float32, "svd3_project input")
    B = m9.shape[0]
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=m9.device)
    L.check(L.load().dcl_svd3_project(B, L.ptr(m9), int(normalize_columns), L.ptr(R), L.stream_ptr()), "svd3_project")
    return R


def _head_struct(head, keep):
    """dcl_pose_head_mlp for a Head_MultiLayerPerceptron [d_in -> d_h1 -> d_h2 -> d_out] (Conv1d, ReLU, Conv1d,
    ReLU, Conv1d); returns None when the module does not have that shape."""
    convs = [m for m in head.layers if isinstance(m, nn.Conv1d)]
    others = [m for m in head.layers if not isinstance(m, (nn.Conv1d, nn.ReLU))]
    if len(convs) != 3 or others or len(list(head.layers)) != 5:
        return None
    s = L.PoseHeadMlp()
    tensors = []
    for conv in convs:
        tensors += [conv.weight.detach().reshape(conv.out_channels, conv.in_channels).contiguous(),
                    conv.bias.detach().contiguous()]
    if any(t.dtype != torch.float32 for t in tensors) or max(c.in_channels for c in convs) > 1024:
        return None
    keep += tensors
    s.w1, s.b1, s.w2, s.b2, s.w3, s.b3 = (L.ptr(t) for t in tensors)
    s.d_in, s.d_h1, s.d_h2, s.d_out = convs[0].in_channels, convs[1].in_channels, convs[2].in_channels, convs[2].out_channels
    return s


def pose_heads(pooled, rot_head, trans_head):
    """regressor_rot / regressor_trans (or the refiner's *_2 pair) on the pooled (B, d_in) feature in ONE kernel
    (csrc/pose_head.cu) -> (ortho9d (B,9), trans (B,3)).  Inference only."""
    import ctypes
    keep = []
    hs, ht = _head_struct(rot_head, keep), _head_struct(trans_head, keep)
    pooled = pooled.contiguous()
    if hs is None or ht is None or pooled.dtype != torch.float32:
        L.warn_once("pose_heads", "dcl_net_b200.pose_heads: head is not a fp32 [d_in -> h1 -> h2 -> out] Conv1d/ReLU "
                                  "stack of width <= 1024; running it as PyTorch layers instead of csrc/pose_head.cu")
        x = pooled.unsqueeze(-1)
        return rot_head(x).squeeze(-1), trans_head(x).squeeze(-1)
    B = pooled.shape[0]
    o9 = torch.empty(B, hs.d_out, dtype=torch.float32, device=pooled.device)
    t3 = torch.empty(B, ht.d_out, dtype=torch.float32, device=pooled.device)
    lib = L.load()
    ps, pt = ctypes.cast(ctypes.pointer(hs), ctypes.c_void_p), ctypes.cast(ctypes.pointer(ht), ctypes.c_void_p)
    ws = torch.empty(lib.dcl_pose_head_workspace_bytes(B, ps, pt), dtype=torch.uint8, device=pooled.device)
    L.check(lib.dcl_pose_head(B, L.ptr(pooled), ps, pt, L.ptr(o9), L.ptr(t3), L.ptr(ws), ws.numel(), L.stream_ptr()),
            "pose_head")
    return o9, t3


class ProjectSO3Function(torch.autograd.Function):
    """R = U diag(1,1,det(UV^T)) V^T of a (B,3,3) matrix M, forward on the dcl_svd3_project kernel, backward in closed
    form without an SVD: R^T M = P is symmetric at the optimum, so a perturbation dM turns R by the skew matrix Omega
    solving Omega P + P Omega = R^T dM - dM^T R, i.e. (tr(P) I - P) omega = axial(R^T dM - dM^T R); transposing that
    linear map gives  dL/dM = R [gamma]x^T-type skew matrix  with  gamma = (tr(P) I - P)^-1 alpha,
    alpha_k = eps_ijk (R^T G)_ij.  Agrees with autograd through the reference's torch.svd formula
    (models/DCL_Net.py:22-35) to rounding, reflections included (tests/test_oracle_golden.py), and — unlike torch.svd —
    is a handful of elementwise / bmm launches that a CUDA graph can capture."""

    @staticmethod
    def forward(ctx, m):
        r = svd3_project(m.reshape(m.shape[0], 9), False)
        ctx.save_for_backward(m, r)
        return r

    @staticmethod
    def backward(ctx, g):
        m, r = ctx.saved_tensors
        return so3_projection_backward(m, r, g)


def so3_projection_backward(m, r, g):
    """dL/dM of R = project_SO3(M) given G = dL/dR; plain torch ops on (B,3,3) tensors (any device / dtype)."""
    rt = r.transpose(1, 2)
    a = rt @ g
    alpha = torch.stack((a[:, 1, 2] - a[:, 2, 1], a[:, 2, 0] - a[:, 0, 2], a[:, 0, 1] - a[:, 1, 0]), 1)
    p = rt @ m
    k = p.diagonal(dim1=1, dim2=2).sum(1).view(-1, 1, 1) * torch.eye(3, dtype=m.dtype, device=m.device) - p
    r0, r1, r2 = k[:, 0], k[:, 1], k[:, 2]
    c0, c1, c2 = torch.cross(r1, r2, dim=1), torch.cross(r2, r0, dim=1), torch.cross(r0, r1, dim=1)
    det = (r0 * c0).sum(1, keepdim=True)
    gamma = (c0 * alpha[:, 0:1] + c1 * alpha[:, 1:2] + c2 * alpha[:, 2:3]) / det      # K^-1 alpha (K symmetric)
    z = torch.zeros_like(gamma[:, 0])
    skew = torch.stack((torch.stack((z, gamma[:, 2], -gamma[:, 1]), 1), torch.stack((-gamma[:, 2], z, gamma[:, 0]), 1),
                        torch.stack((gamma[:, 1], -gamma[:, 0], z), 1)), 1)
    return r @ skew


USE_TORCH_SVD_AUTOGRAD = False   # True: differentiate through torch.svd as the reference does (A/B, tests)


def ortho9d2matrix(x_raw, y_raw, z_raw):
    """Rotation from three raw 3-vectors: columns normalised by (|v| + 1e-8), then the SO(3) projection
    U diag(1,1,det(UV^T)) V^T (models/DCL_Net.py:22-35).  Runs the dcl_svd3_project kernel; when autograd is recording
    a gradient through it (training), the column normalisation is differentiated by autograd and the projection by
    ProjectSO3Function (same kernel forward, closed-form backward) — or, with USE_TORCH_SVD_AUTOGRAD, through
    torch.svd exactly as the reference does."""
    if torch.is_grad_enabled() and (x_raw.requires_grad or y_raw.requires_grad or z_raw.requires_grad):
        def unit(v):
            return v / (torch.sqrt(v.pow(2).sum(1, keepdim=True)) + 1e-8)
        m = torch.stack((unit(x_raw), unit(y_raw), unit(z_raw)), dim=2)
        if not USE_TORCH_SVD_AUTOGRAD and m.is_cuda and m.dtype == torch.float32:
            return ProjectSO3Function.apply(m.contiguous())
        u, _, v = torch.svd(m)
        sigma = torch.ones(m.shape[0], 3, dtype=m.dtype, device=m.device)
        sigma[:, -1] = torch.bmm(u, v.transpose(1, 2)).det()
        return u @ torch.diag_embed(sigma) @ v.transpose(1, 2)
    return svd3_project(torch.cat((x_raw, y_raw, z_raw), dim=1), True)


def weighted_kabsch(src, dst, w):
    """Confidence-weighted rigid fit  min sum_i w_i |R src_i + t - dst_i|^2.
    src, dst (B,N,3), w (B,N) fp32 -> R (B,3,3), t (B,3)."""
    src, dst, w = src.contiguous(), dst.contiguous(), w.contiguous()
    B, N, _ = src.shape
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=src.device)
    t = torch.empty(B, 3, dtype=torch.float32, device=src.device)
    L.check(L.load().dcl_weighted_kabsch(B, N, L.ptr(src), L.ptr(dst), L.ptr(w), L.ptr(R), L.ptr(t), L.stream_ptr()),
            "weighted_kabsch")
    return R, t


class Network(nn.Module):
    def __init__(self, cfg, mode="train", backbone=None, c_m=64, with_backbone=False) -> None:
        """cfg needs n_inp, n_tmp, unit_voxel_extent (reference configs/*.yaml: model section).
        with_backbone: create the two sparse-conv towers `backbone_inp` / `backbone_tmp` (models/DCL_Net.py:47-52, same
        parameter names) and run them on the device (backbone.SparseTowers) in forward(data) / forward_from_points.
        backbone: alternatively a callable data -> (levels_inp, levels_tmp, points_inp (b*n,3), points_tmp (b*n,3), b),
        each `levels_*` a list of four objects with .features (Mv,C_l) / .indices (Mv,4).
        c_m: width of the pose-insensitive branch (64 in the reference; 128 in BASELINE.json's config)."""
        super().__init__()
        self.mode = mode
        self.n_inp, self.n_tmp = cfg.n_inp, cfg.n_tmp
        self.unit_voxel_extent = np.array(cfg.unit_voxel_extent)
        self.backbone = backbone
        if with_backbone:
            from .backbone import Backbone_SPCONV
            self.backbone_inp = Backbone_SPCONV()
            self.backbone_tmp = Backbone_SPCONV()
        self._towers = None
        self.stage1_get_point_feats = Ops_GetPointFeat_spconv(
            scale_lists=[2, 4, 6, 8], unit_voxel_extent=self.unit_voxel_extent, voxel_num_limit=[64, 64, 64])
        blk = partial(BasicBlock_3DCONV, size=1, bias=False, stride=1, padding=0, norm=True, act="relu", drop=0.0)
        for name in ("Xc_p1", "Xc_m1", "Yo_p1", "Yo_m1", "Xc_p2", "Xc_m2", "Yo_p2", "Yo_m2"):
            setattr(self, "disengage_" + name,
                    nn.Sequential(blk(dim_in=480, dim_out=256), blk(dim_in=256, dim_out=256 if "_p" in name else c_m)))
        plain = (["relu", "relu", "none"], [False] * 3, [0.0] * 3)
        fuse = (["relu"] * 3, [True] * 3, [0.0] * 3)
        self.regressor_Xo = Head_MultiLayerPerceptron([256, 256, 128, 3], *plain)
        self.regressor_Yc = Head_MultiLayerPerceptron([256, 256, 128, 3], *plain)
        self.regressor_conf = Head_MultiLayerPerceptron([c_m * 2, 128, 128, 1], *plain)
        self.regressor_conf_bi = Head_MultiLayerPerceptron([c_m * 2, 128, 128, 1], *plain)
        self.neck_fuser = Head_MultiLayerPerceptron([256 * 2, 512, 512, 1024], *fuse)
        self.neck_fuser_bi = Head_MultiLayerPerceptron([256 * 2, 512, 512, 1024], *fuse)
        self.regressor_rot = Head_MultiLayerPerceptron([1024, 512, 128, 9], *plain)
        self.regressor_trans = Head_MultiLayerPerceptron([1024, 512, 128, 3], *plain)
        self.use_fused_tail = True   # inference: pointwise MLPs on tensor cores (fused_tail.py)
        # operand precision of the tensor-core inference path (fused_tail.py): "fp16" = activations rounded once to
        # fp16, fp16 hi/lo weights, split-operand FDA logits (default; ~4e-5 on features, ~1e-5 deg on poses against
        # the fp32 graph); "fp32-faithful" = every operand a bf16 hi/lo pair (3 MMAs per product, ~1e-6)
        self.precision = "fp16"
        self._fused_tail = None

    def _apply(self, fn, *args, **kwargs):
        self._fused_tail = self._towers = None      # parameters moved / cast: packed copies are stale
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._fused_tail = self._towers = None
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode=True):
        if mode != self.training:    # BatchNorm folding depends on the mode; an unchanged mode keeps the packed weights
            self._fused_tail = self._towers = None
        return super().train(mode)

    def towers(self, b, device, sample_clouds=None, caps=None):
        """The device towers for batches of b instances (backbone.SparseTowers), built on first use.  Buffer
        capacities come from `caps` or are planned from `sample_clouds` (two (b*n,3) clouds) with a 30 % margin."""
        from .backbone import SparseTowers
        if not hasattr(self, "backbone_inp"):
            raise RuntimeError("this Network was built without its sparse-conv towers (with_backbone=True)")
        if self.n_inp != self.n_tmp:
            raise RuntimeError("the device towers take equally sized observed / template clouds")
        tw = self._towers
        rounded = None if caps is None else [int((c + 127) // 128 * 128) for c in caps]
        if tw is None or tw.b != b or (rounded is not None and rounded != tw.caps):
            if caps is None:
                caps = SparseTowers.plan_capacities(sample_clouds, b, self.n_inp, device)
            tw = self._towers = SparseTowers(self.backbone_inp, self.backbone_tmp, device, b, self.n_inp, caps,
                                             unit=float(self.unit_voxel_extent[0]))
        return tw

    @torch.no_grad()
    def forward_from_points(self, points_inp, rgb_inp, points_tmp, rgb_tmp, b, caps=None):
        """Raw clouds (b*n,3) + colours (b*n,3) of the observation and the template -> prediction dict: device
        voxelisation, both towers, point-feature interpolation, FDA, pose.  Inference only."""
        if self.training:
            raise RuntimeError("forward_from_points is an inference path: call eval() first")
        tw = self.towers(b, points_inp.device, (points_inp, points_tmp), caps)
        levels_inp, levels_tmp = tw.run(points_inp.contiguous(), rgb_inp.contiguous(), points_tmp.contiguous(),
                                        rgb_tmp.contiguous())
        return self.forward_from_backbone(levels_inp, levels_tmp, points_inp, points_tmp, b)

    # ---- entry points -----------------------------------------------------------------
    def forward(self, data):
        """The reference's interface (models/DCL_Net.py:155-259).  With the built-in towers the per-point tensors
        data[k]["feats"] (N,7) = [1, rgb, xyz] are voxelised on the device — the same voxels and the same means as
        data[k]["occupied_voxels"] / ["v2p_maps"] describe (pointgroup_ops.voxelization_idx, mode 4), which are
        therefore not read — and the call checks the towers' overflow flags (one host sync)."""
        if hasattr(self, "backbone_inp") and self.backbone is None:
            if self.training:
                raise RuntimeError("the device towers are inference-only: call eval() (training through the sparse "
                                   "convolutions is outside this path)")
            dev = next(self.parameters()).device
            f_inp, f_tmp = data["inp"]["feats"].to(dev), data["tmp"]["feats"].to(dev)
            b = data["batch_offsets"].shape[0] - 1
            points_inp, points_tmp = f_inp[:, 4:7].contiguous(), f_tmp[:, 4:7].contiguous()
            with torch.no_grad():
                pred = self.forward_from_points(points_inp, f_inp[:, 1:4].contiguous(), points_tmp,
                                                f_tmp[:, 1:4].contiguous(), b)
            self._towers.check_errors()
        elif self.backbone is None:
            raise RuntimeError("Network.forward(data) needs the sparse-conv towers: build the Network with "
                               "with_backbone=True (or inject a `backbone` callable), or use forward_from_backbone / "
                               "forward_from_point_feats")
        else:
            levels_inp, levels_tmp, points_inp, points_tmp, b = self.backbone(data)
            pred = self.forward_from_backbone(levels_inp, levels_tmp, points_inp, points_tmp, b)
        if self.mode != "test" and "flags" in data:
            pred["sym_flag"] = data["flags"].to(points_inp.device)
        data.setdefault("labels", {})
        data["labels"]["points_tmp"] = points_tmp.view(b, self.n_tmp, -1)
        data["labels"]["points_inp"] = points_inp.view(b, self.n_inp, -1)
        return pred

    def _fused(self, b):
        """The tensor-core inference path (fused_tail.FusedTail).  None — i.e. the PyTorch layer modules, whose GEMMs
        are library calls — only (i) while autograd is recording or the module is in training mode (BatchNorm
        batch statistics; the training path), (ii) when the caller switched it off, or (iii) for shapes the packed
        kernels do not take, and then with a warning: inference never leaves the tensor-core path silently."""
        from .fused_tail import FusedTail
        if torch.is_grad_enabled() or self.training or not self.use_fused_tail:
            return None
        why = FusedTail.unsupported_reason(self, b)
        if why is not None:
            L.warn_once(("fused_tail", why), f"dcl_net_b200.Network: inference falls back to PyTorch layer modules ({why})")
            return None
        fmt = FusedTail.pick_fmt(self)
        if self._fused_tail is None or self._fused_tail.fmt != fmt:
            self._fused_tail = FusedTail(self, fmt)
        return self._fused_tail

    def invalidate_packed_weights(self):
        """Call after changing parameters in place (load_state_dict, .to()): weights are re-packed lazily."""
        self._fused_tail = None

    def forward_from_backbone(self, levels_inp, levels_tmp, points_inp, points_tmp, b):
        """Pyramid levels of both towers -> prediction dict (models/DCL_Net.py:182-255)."""
        dev = points_inp.device
        ids_inp = torch.arange(b, device=dev).repeat_interleave(points_inp.shape[0] // b)
        ids_tmp = torch.arange(b, device=dev).repeat_interleave(points_tmp.shape[0] // b)
        fused = self._fused(b)
        if fused is not None:
            pm_xc, pm_yo = self.stage1_get_point_feats.forward_pm_pair(points_inp, ids_inp, levels_inp,
                                                                        points_tmp, ids_tmp, levels_tmp, fused.fmt)
            return fused.forward(pm_xc, pm_yo, b)
        F_Xc = self.stage1_get_point_feats(points_inp, ids_inp, *levels_inp)
        F_Yo = self.stage1_get_point_feats(points_tmp, ids_tmp, *levels_tmp)
        return self.forward_from_point_feats(F_Xc, F_Yo, b)

    def forward_from_point_feats(self, F_Xc, F_Yo, b):
        """F_Xc (b*n_inp, 480), F_Yo (b*n_tmp, 480) -> prediction dict (models/DCL_Net.py:187-255)."""
        fused = self._fused(b)
        if fused is not None:
            from .fused_tail import pm_pack_rows
            return fused.forward(pm_pack_rows(F_Xc, fused.fmt), pm_pack_rows(F_Yo, fused.fmt), b)
        if self._train_kernels(F_Xc):
            return self._forward_train(F_Xc, F_Yo, b)
        F_Xc = F_Xc.view(b, self.n_inp, -1).transpose(1, 2)[:, :, :, None, None]
        F_Yo = F_Yo.view(b, self.n_tmp, -1).transpose(1, 2)[:, :, :, None, None]
        sq = lambda t: t.squeeze(-1).squeeze(-1)
        F_Xc_p1, F_Xc_m1 = sq(self.disengage_Xc_p1(F_Xc)), sq(self.disengage_Xc_m1(F_Xc))
        F_Xc_p2, F_Xc_m2 = sq(self.disengage_Xc_p2(F_Xc)), sq(self.disengage_Xc_m2(F_Xc))
        F_Yo_p1, F_Yo_m1 = sq(self.disengage_Yo_p1(F_Yo)), sq(self.disengage_Yo_m1(F_Yo))
        F_Yo_p2, F_Yo_m2 = sq(self.disengage_Yo_p2(F_Yo)), sq(self.disengage_Yo_m2(F_Yo))

        # dual FDA: both attention products of each direction in one fused kernel
        F_Xo_p, F_Xo_m = fda_align(F_Xc_m1, F_Yo_m1, F_Yo_p1)
        F_Yc_p, F_Yc_m = fda_align(F_Yo_m2, F_Xc_m2, F_Xc_p2)
        Xo_pred = self.regressor_Xo(F_Xo_p) if self.mode != "test" else None
        Yc_pred = self.regressor_Yc(F_Yc_p) if self.mode != "test" else None

        conf_1 = self.regressor_conf(torch.cat([F_Xc_m1, F_Xo_m], dim=1))
        conf_2 = self.regressor_conf_bi(torch.cat([F_Yc_m, F_Yo_m2], dim=1))
        conf = torch.sigmoid(torch.cat([conf_1, conf_2], dim=2))
        conf_softmax = torch.softmax(conf, dim=2)

        F_p1 = self.neck_fuser(torch.cat([F_Xc_p1, F_Xo_p], dim=1))
        F_p2 = self.neck_fuser_bi(torch.cat([F_Yc_p, F_Yo_p2], dim=1))
        F_p_wei = torch.sum(torch.cat([F_p1, F_p2], dim=2) * conf_softmax, dim=2, keepdim=True)

        ortho9d_pred = self.regressor_rot(F_p_wei).squeeze(-1)
        rot_pred = ortho9d2matrix(ortho9d_pred[:, :3], ortho9d_pred[:, 3:6], ortho9d_pred[:, 6:])
        trans_pred = self.regressor_trans(F_p_wei).squeeze(-1)

        prediction = {"trans_pred": trans_pred, "rot_pred": rot_pred, "conf": conf.squeeze(1), "F_Xo_p": F_Xo_p}
        if self.mode != "test":
            prediction.update({"Xo_pred": Xo_pred.transpose(1, 2), "Yc_pred": Yc_pred.transpose(1, 2)})
        prediction["_debug"] = {"F_Yc_p": F_Yc_p, "F_Xo_m": F_Xo_m, "F_Yc_m": F_Yc_m, "ortho9d": ortho9d_pred}
        return prediction

    def _train_kernels(self, F_Xc):
        """Training runs on the tensor-core training path (train_tail.py) when the module is in train mode
        (BatchNorm on batch statistics) and the shapes fit its tiles; `use_train_kernels = False` selects the
        PyTorch layer modules (library GEMMs) instead, e.g. for A/B runs."""
        if not (self.training and torch.is_grad_enabled() and getattr(self, "use_train_kernels", True)):
            return False
        ok = (F_Xc.is_cuda and F_Xc.dtype == torch.float32 and self.n_inp == self.n_tmp and self.n_inp % 128 == 0
              and F_Xc.shape[1] % 32 == 0 and self.disengage_Xc_m1[1].layers[0].out_channels in (64, 128))
        if not ok:
            L.warn_once("train_kernels", "dcl_net_b200.Network: training falls back to PyTorch layer modules "
                                         "(needs CUDA fp32 inputs, n_inp == n_tmp, n % 128 == 0, c_m in {64, 128})")
        return ok

    def _forward_train(self, F_Xc, F_Yo, b):
        """models/DCL_Net.py:187-235 in train mode with every pointwise MLP stack — forward and backward — on the
        tcgen05 GEMM (train_tail.mlp_stacks) and the fused FDA kernels; same modules, parameters and BatchNorm
        buffers as the layer path below."""
        from .train_tail import StackSpec, disengage_layers, head_layers, mlp_stacks
        n = self.n_inp
        names = ("Xc_p1", "Xc_m1", "Xc_p2", "Xc_m2", "Yo_p1", "Yo_m1", "Yo_p2", "Yo_m2")
        F_Xc, F_Yo = F_Xc.contiguous(), F_Yo.contiguous()
        dis = mlp_stacks([StackSpec([(F_Xc if k.startswith("Xc") else F_Yo, "rm")],
                                    disengage_layers(getattr(self, "disengage_" + k))) for k in names], b, n)
        F_Xc_p1, F_Xc_m1, F_Xc_p2, F_Xc_m2, F_Yo_p1, F_Yo_m1, F_Yo_p2, F_Yo_m2 = dis

        F_Xo_p, F_Xo_m = fda_align(F_Xc_m1, F_Yo_m1, F_Yo_p1)
        F_Yc_p, F_Yc_m = fda_align(F_Yo_m2, F_Xc_m2, F_Xc_p2)
        Xo_pred = Yc_pred = None
        if self.mode != "test":
            (lx, wx), (ly, wy) = head_layers(self.regressor_Xo), head_layers(self.regressor_Yc)
            xo, yc = mlp_stacks([StackSpec([(F_Xo_p, "cm")], lx), StackSpec([(F_Yc_p, "cm")], ly)], b, n)
            Xo_pred, Yc_pred = xo[:, :wx], yc[:, :wy]

        (l1, w1), (l2, w2) = head_layers(self.regressor_conf), head_layers(self.regressor_conf_bi)
        c1, c2 = mlp_stacks([StackSpec([(F_Xc_m1, "cm"), (F_Xo_m, "cm")], l1),
                             StackSpec([(F_Yc_m, "cm"), (F_Yo_m2, "cm")], l2)], b, n)
        conf = torch.sigmoid(torch.cat([c1[:, :w1], c2[:, :w2]], dim=2))
        conf_softmax = torch.softmax(conf, dim=2)

        (f1, _), (f2, _) = head_layers(self.neck_fuser), head_layers(self.neck_fuser_bi)
        F_p1, F_p2 = mlp_stacks([StackSpec([(F_Xc_p1, "cm"), (F_Xo_p, "cm")], f1),
                                 StackSpec([(F_Yc_p, "cm"), (F_Yo_p2, "cm")], f2)], b, n)
        # confidence-weighted pooling over the 2n correspondences (DCL_Net.py:228) without the (b,1024,2n) concat
        F_p_wei = (torch.bmm(F_p1, conf_softmax[:, 0, :n].unsqueeze(2)) +
                   torch.bmm(F_p2, conf_softmax[:, 0, n:].unsqueeze(2)))

        ortho9d_pred = self.regressor_rot(F_p_wei).squeeze(-1)
        rot_pred = ortho9d2matrix(ortho9d_pred[:, :3], ortho9d_pred[:, 3:6], ortho9d_pred[:, 6:])
        trans_pred = self.regressor_trans(F_p_wei).squeeze(-1)
        prediction = {"trans_pred": trans_pred, "rot_pred": rot_pred, "conf": conf.squeeze(1), "F_Xo_p": F_Xo_p}
        if self.mode != "test":
            prediction.update({"Xo_pred": Xo_pred.transpose(1, 2), "Yc_pred": Yc_pred.transpose(1, 2)})
        prediction["_debug"] = {"F_Yc_p": F_Yc_p, "F_Xo_m": F_Xo_m, "F_Yc_m": F_Yc_m, "ortho9d": ortho9d_pred}
        return prediction

    def attention_maps(self, F_Xc_m1, F_Yo_m1, F_Yo_m2, F_Xc_m2):
        """The two (B,M,N) attention maps the reference keeps in `attention_map` / `attention_map_bi`."""
        _, _, lse1 = fda_align(F_Xc_m1, F_Yo_m1, F_Yo_m1.new_zeros(F_Yo_m1.shape[0], 256, F_Yo_m1.shape[2]), True)
        _, _, lse2 = fda_align(F_Yo_m2, F_Xc_m2, F_Xc_m2.new_zeros(F_Xc_m2.shape[0], 256, F_Xc_m2.shape[2]), True)
        return fda_attention_map(F_Xc_m1, F_Yo_m1, lse1), fda_attention_map(F_Yo_m2, F_Xc_m2, lse2)
