"""TEST INFRASTRUCTURE — pure-PyTorch restatement of the model-level part of the hot path.

Device-agnostic (CPU or CUDA), fp32 unless stated, no custom kernels: this is what the CUDA
path is checked against and what bench.py times as the CPU baseline ("port").  The reference
itself cannot run here: every model file hard-codes .cuda() and imports spconv
(models/DCL_Net.py:4,32,157).  Each piece cites the reference lines it restates; the
restatement is pinned against the reference's own classes, imported from /root/reference,
by oracle/make_golden.py -> tests/golden/model_*.npz.

Pieces
  normalize_vector / ortho9d2matrix   utils/transform3D.py:16-21, models/DCL_Net.py:15-36
  aligner                             models/Modules.py:166-169
  disengage_stack / mlp_head          models/Modules.py:58-97, 173-201
  TailNetwork                         models/DCL_Net.py:56-151 (modules), :187-244 (wiring)
  RefinerNet / stage2_refine          models/refiner.py:57-95, tools/test_YCBV_stage2.py:204-225
  nearest_neighbor_interpolate        models/Modules.py:213-251
  weighted_kabsch                     north_star part 3 (no reference counterpart)
"""
import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------- pose
def normalize_vector(v):
    # utils/transform3D.py:16-21 — epsilon is ADDED to the norm, not a clamp.
    mag = torch.sqrt(v.pow(2).sum(1, keepdim=True)) + 1e-8
    return v / mag


def ortho9d2matrix(x_raw, y_raw, z_raw):
    # models/DCL_Net.py:22-35: the three normalised 3-vectors are the COLUMNS of M;
    # R = U diag(1, 1, det(U V^T)) V^T with torch.svd's descending singular values.
    m = torch.stack((normalize_vector(x_raw), normalize_vector(y_raw), normalize_vector(z_raw)), dim=2)
    u, _, v = torch.svd(m)
    sigma = torch.ones(m.shape[0], 3, dtype=m.dtype, device=m.device)
    sigma[:, -1] = torch.bmm(u, v.transpose(1, 2)).det()
    return u @ torch.diag_embed(sigma) @ v.transpose(1, 2)


def project_so3(m):
    """Same projection for an arbitrary (B,3,3) matrix (no column normalisation), fp64."""
    m64 = m.double()
    u, _, vh = torch.linalg.svd(m64)
    d = torch.det(u @ vh)
    s = torch.ones(m.shape[0], 3, dtype=torch.float64, device=m.device)
    s[:, -1] = d
    return (u @ torch.diag_embed(s) @ vh)


def weighted_kabsch(src, dst, w):
    """argmin_{R,t} sum_i w_i |R src_i + t - dst_i|^2, fp64.  src,dst (B,N,3), w (B,N)."""
    src, dst, w = src.double(), dst.double(), w.double()
    ws = w.sum(1, keepdim=True)
    pm = (w.unsqueeze(-1) * src).sum(1) / ws
    qm = (w.unsqueeze(-1) * dst).sum(1) / ws
    p, q = src - pm.unsqueeze(1), dst - qm.unsqueeze(1)
    mt = torch.einsum("bn,bni,bnj->bij", w, q, p)  # H^T
    r = project_so3(mt)
    t = qm - torch.einsum("bij,bj->bi", r, pm)
    return r, t


def rotation_angle_deg(r_a, r_b):
    """Geodesic distance between rotations in degrees."""
    rel = r_a.double().transpose(1, 2) @ r_b.double()
    tr = rel.diagonal(dim1=1, dim2=2).sum(1)
    skew = rel - rel.transpose(1, 2)
    s = torch.sqrt(skew[:, 2, 1] ** 2 + skew[:, 0, 2] ** 2 + skew[:, 1, 0] ** 2) / 2.0
    return torch.rad2deg(torch.atan2(s, (tr - 1.0) / 2.0))


# --------------------------------------------------------------------------- FDA
def aligner(ri_1, ri_2, re_2):
    # models/Modules.py:166-169
    a = torch.softmax(torch.bmm(ri_2.transpose(1, 2), ri_1), dim=1)
    return torch.bmm(re_2, a), a


def fda_direction(ri_1, ri_2, re_2):
    """Aligner plus the confidence-branch product of models/DCL_Net.py:213/215."""
    re_embed, a = aligner(ri_1, ri_2, re_2)
    return re_embed, torch.bmm(ri_2, a), a


class _Wrap(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = layers

    def forward(self, x):
        return self.layers(x)


def disengage_stack(dim_out):
    # DCL_Net.py:69-100: 480 -> 256 -> dim_out, each Conv3d1x1 + BN3d + ReLU
    def blk(i, o):
        return _Wrap(nn.Sequential(nn.Conv3d(i, o, 1, 1, 0, bias=False), nn.BatchNorm3d(o), nn.ReLU()))
    return nn.Sequential(blk(480, 256), blk(256, dim_out))


def mlp_head(dims, acts, bns):
    # models/Modules.py:173-201: Conv1d(k=1) -> activation -> [BatchNorm1d]  (BN AFTER the ReLU)
    layers = []
    for i, (act, bn) in enumerate(zip(acts, bns)):
        layers.append(nn.Conv1d(dims[i], dims[i + 1], 1))
        if act == "relu":
            layers.append(nn.ReLU())
        elif act != "none":
            raise NotImplementedError(act)
        if bn:
            layers.append(nn.BatchNorm1d(dims[i + 1]))
    return _Wrap(nn.Sequential(*layers))


class TailNetwork(nn.Module):
    """Everything of Network (models/DCL_Net.py) after the point-feature interpolation.

    Module names, creation ORDER (hence default-init RNG consumption) and parameter names
    equal the reference's Network minus its two sparse-conv backbones, so a reference
    checkpoint's tail loads with strict=False and a same-seed construction gives the same
    weights as constructing the reference Network with parameter-free backbones.
    c_m is the width of the pose-insensitive branch: 64 in the reference, 128 in BASELINE.json.
    """

    def __init__(self, mode="test", c_m=64):
        super().__init__()
        self.mode = mode
        self.c_m = c_m
        for name in ("Xc_p1", "Xc_m1", "Yo_p1", "Yo_m1", "Xc_p2", "Xc_m2", "Yo_p2", "Yo_m2"):
            setattr(self, "disengage_" + name, disengage_stack(256 if "_p" in name else c_m))
        r3 = (["relu", "relu", "none"], [False] * 3)
        self.regressor_Xo = mlp_head([256, 256, 128, 3], *r3)
        self.regressor_Yc = mlp_head([256, 256, 128, 3], *r3)
        self.regressor_conf = mlp_head([c_m * 2, 128, 128, 1], *r3)
        self.regressor_conf_bi = mlp_head([c_m * 2, 128, 128, 1], *r3)
        self.neck_fuser = mlp_head([512, 512, 512, 1024], ["relu"] * 3, [True] * 3)
        self.neck_fuser_bi = mlp_head([512, 512, 512, 1024], ["relu"] * 3, [True] * 3)
        self.regressor_rot = mlp_head([1024, 512, 128, 9], *r3)
        self.regressor_trans = mlp_head([1024, 512, 128, 3], *r3)

    def forward(self, f_xc_flat, f_yo_flat, b, n_inp, n_tmp):
        """f_xc_flat (b*n_inp, 480), f_yo_flat (b*n_tmp, 480): outputs of Ops_GetPointFeat_spconv."""
        # DCL_Net.py:188-200
        f_xc = f_xc_flat.view(b, n_inp, -1).transpose(1, 2)[:, :, :, None, None]
        f_yo = f_yo_flat.view(b, n_tmp, -1).transpose(1, 2)[:, :, :, None, None]
        sq = lambda t: t.squeeze(-1).squeeze(-1)
        xc_p1, xc_m1 = sq(self.disengage_Xc_p1(f_xc)), sq(self.disengage_Xc_m1(f_xc))
        xc_p2, xc_m2 = sq(self.disengage_Xc_p2(f_xc)), sq(self.disengage_Xc_m2(f_xc))
        yo_p1, yo_m1 = sq(self.disengage_Yo_p1(f_yo)), sq(self.disengage_Yo_m1(f_yo))
        yo_p2, yo_m2 = sq(self.disengage_Yo_p2(f_yo)), sq(self.disengage_Yo_m2(f_yo))
        # :206-215
        f_xo_p, a = aligner(xc_m1, yo_m1, yo_p1)
        xo_pred = self.regressor_Xo(f_xo_p)
        f_yc_p, a_bi = aligner(yo_m2, xc_m2, xc_p2)
        yc_pred = self.regressor_Yc(f_yc_p)
        f_xo_m = torch.bmm(yo_m1, a)
        f_yc_m = torch.bmm(xc_m2, a_bi)
        # :214-220
        conf_1 = self.regressor_conf(torch.cat([xc_m1, f_xo_m], dim=1))
        conf_2 = self.regressor_conf_bi(torch.cat([f_yc_m, yo_m2], dim=1))
        conf = torch.sigmoid(torch.cat([conf_1, conf_2], dim=2))
        conf_softmax = torch.softmax(conf, dim=2)
        # :223-235
        f_p1 = self.neck_fuser(torch.cat([xc_p1, f_xo_p], dim=1))
        f_p2 = self.neck_fuser_bi(torch.cat([f_yc_p, yo_p2], dim=1))
        f_p_wei = torch.sum(torch.cat([f_p1, f_p2], dim=2) * conf_softmax, dim=2, keepdim=True)
        o9 = self.regressor_rot(f_p_wei).squeeze(-1)
        rot = ortho9d2matrix(o9[:, :3], o9[:, 3:6], o9[:, 6:])
        trans = self.regressor_trans(f_p_wei).squeeze(-1)
        out = {"trans_pred": trans, "rot_pred": rot, "conf": conf.squeeze(1), "F_Xo_p": f_xo_p}
        if self.mode != "test":
            out.update({"Xo_pred": xo_pred.transpose(1, 2), "Yc_pred": yc_pred.transpose(1, 2)})
        out["_debug"] = {"F_Yc_p": f_yc_p, "F_Xo_m": f_xo_m, "F_Yc_m": f_yc_m, "ortho9d": o9}
        return out


class RefinerNet(nn.Module):
    # models/refiner.py:57-95
    def __init__(self):
        super().__init__()
        r3 = (["relu", "relu", "none"], [False] * 3)
        self.MLP_share = mlp_head([259, 512, 512, 1024], ["relu"] * 3, [False] * 3)
        self.regressor_rot2 = mlp_head([1024, 512, 128, 9], *r3)
        self.regressor_trans2 = mlp_head([1024, 512, 128, 3], *r3)

    def forward(self, input_dict):
        feats, conf = input_dict["input_features"], input_dict["conf"]
        conf_softmax = torch.softmax(conf.unsqueeze(1), dim=2)[:, :, :1024]  # refiner.py:81
        shared = (self.MLP_share(feats) * conf_softmax).sum(dim=2, keepdim=True)
        o9 = self.regressor_rot2(shared).squeeze(-1)
        d_t = self.regressor_trans2(shared).squeeze(-1)
        return {"trans_pred": d_t, "rot_pred": ortho9d2matrix(o9[:, :3], o9[:, 3:6], o9[:, 6:])}


def stage2_refine(refiner, points_inp, rot, trans, f_xo_p, conf, iterations=2):
    # tools/test_YCBV_stage2.py:204-225
    rot_cur, trans_cur = rot, trans
    cur = torch.bmm(points_inp - trans_cur.unsqueeze(1), rot_cur)
    inp = torch.cat([cur.transpose(1, 2), f_xo_p], dim=1)
    for _ in range(iterations):
        out = refiner({"input_features": inp, "conf": conf})
        trans_cur = (rot_cur @ out["trans_pred"].unsqueeze(2)).squeeze(2) + trans_cur
        rot_cur = rot_cur @ out["rot_pred"]
        cur = torch.bmm(points_inp - trans_cur.unsqueeze(1), rot_cur)
        inp = torch.cat([cur.transpose(1, 2), f_xo_p], dim=1)
    return rot_cur, trans_cur


# ------------------------------------------------------- point-feature interpolation
def voxel_centres(indices, offset, voxel_extent):
    # models/Modules.py:204-211 (Ops_tensor2points): centre = idx*ext + offset + ext/2, batch id kept
    out = indices.float().clone()
    ext = torch.as_tensor(voxel_extent, dtype=torch.float32, device=out.device)
    off = torch.as_tensor(offset, dtype=torch.float32, device=out.device)
    out[:, 1:] = out[:, 1:] * ext + off + 0.5 * ext
    return out


def sp_three_nn_torch(unknown, known):
    """Batch-id-aware 3-NN, squared distances + indices, chunked; the semantics of
    libs/pointnet_sp/src/interpolate_gpu.cu:9-56 up to fp32 summation order (the C oracle
    is the bit-faithful one).  Ties resolve to the lowest index via a stable sort."""
    n = unknown.shape[0]
    d2 = torch.full((n, 3), float("inf"), dtype=torch.float32, device=unknown.device)
    idx = torch.zeros((n, 3), dtype=torch.int32, device=unknown.device)
    for b in torch.unique(unknown[:, 0]).tolist():
        qs = (unknown[:, 0] == b).nonzero().squeeze(1)
        ks = (known[:, 0] == b).nonzero().squeeze(1)
        if ks.numel() == 0:
            continue
        diff = unknown[qs, None, 1:] - known[None, ks, 1:]
        dd = diff[..., 1] * diff[..., 1] + diff[..., 0] * diff[..., 0] + diff[..., 2] * diff[..., 2]
        k = min(3, ks.numel())
        order = torch.sort(dd, dim=1, stable=True)
        d2[qs, :k] = order.values[:, :k]
        idx[qs, :k] = ks[order.indices[:, :k]].int()
    return d2, idx


def nearest_neighbor_interpolate(target_points, query_points, query_feats, three_nn=sp_three_nn_torch):
    # models/Modules.py:213-226
    d2, idx = three_nn(target_points, query_points)
    dist = torch.sqrt(d2)
    recip = 1.0 / (dist + 1e-8)
    weight = recip / torch.sum(recip, dim=1, keepdim=True)
    f = query_feats[idx.long()]  # (n, 3, C)
    return torch.addcmul(torch.addcmul(weight[:, 1:2] * f[:, 1], weight[:, 0:1], f[:, 0]), weight[:, 2:3], f[:, 2])


def get_point_feats(points, batch_ids, levels, unit_voxel_extent, scale_list=(2, 4, 6, 8), voxel_num_limit=(64, 64, 64),
                    three_nn=sp_three_nn_torch):
    """models/Modules.py:227-251.  levels: list of (features (Mv,C), indices (Mv,4) int bxyz)."""
    unit = np.asarray(unit_voxel_extent, dtype=np.float64)
    offset = -0.5 * unit * np.asarray(voxel_num_limit)
    pts = torch.cat([batch_ids.view(-1, 1).float(), points], 1)
    outs = []
    for scale, (feats, indices) in zip(scale_list, levels):
        centres = voxel_centres(indices, offset, unit * scale)
        outs.append(nearest_neighbor_interpolate(pts, centres, feats, three_nn))
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------- losses / metric
def cd_dis(pred, target):
    # models/DCL_Net.py:307-311
    dis = torch.norm(pred.unsqueeze(2) - target.unsqueeze(1), dim=3)
    return 0.5 * (torch.min(dis, 2)[0] + torch.min(dis, 1)[0])


def adds(points_posed_pred, points_posed_gt):
    # tools/test_YCBV_stage1.py:188
    return torch.mean(torch.min(torch.norm(points_posed_pred.unsqueeze(2) - points_posed_gt.unsqueeze(1), dim=3), 2)[0], dim=1)
