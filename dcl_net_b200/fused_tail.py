"""Inference engine for everything of Network.forward after the backbone (models/DCL_Net.py:182-244) and of the
stage-2 refiner loop (models/refiner.py:57-95), with the pointwise MLP stacks on tensor cores (csrc/pm_gemm.cu)
instead of fp32 cuDNN/cuBLAS calls.

Activations travel between layers as operand images (include/dcl_b200.h) written by the producing kernel's epilogue;
an fp32 channel-major tensor exists only where the reference interface needs one (F_Xo_p, which stage 2 reads).
Two formats (`fmt`):
   1  "PM16": activations rounded once to fp16, weights split into fp16 hi + lo — 2 MMAs per product, 2 bytes per
      activation element; the FDA logits keep bf16 hi/lo operands (3 MMAs), its P V products run on fp16 P and V
      (1 MMA).  The default of the inference path (Network.precision = "fp16"): measured 4e-5 relative on the
      features and ~1e-5 deg on the poses against the fp32 graph (bars 1e-3 / 0.01 deg).
   0  "PM": every operand as a bf16 hi/lo pair, 3 MMAs per product (~2^-17; Network.precision = "fp32-faithful").  Weights are packed once per Network (eval-mode BatchNorm before a ReLU is folded
into the convolution; BatchNorm after a ReLU becomes the GEMM epilogue's per-channel affine).

Dataflow (test mode), b instances of n points per side, R = b*n rows — 17 launches:
   PM(F_Xc), PM(F_Yo)  (R x 480)      <- pointnet_sp fused 3-NN interpolation, or pack of an fp32 (R,480) matrix
   8 x [480 -> 256]  ReLU             one launch, 8 problems
   4 x [256 -> 256], 4 x [256 -> c_m] two launches; epilogues write PM images and the FDA query / key / value images
   dual fused FDA (csrc/fda.cu)       one launch, both directions: F_Xo_p (+ fp32), F_Xo_m, F_Yc_p, F_Yc_m as PM images
   confidence heads [2c_m -> 128 -> 128 (-> 1 as a dot in the epilogue)], then dcl_conf_weights (sigmoid + softmax)
   fusers [512 -> 512 -> 512 -> 1024] ReLU+BN, the last layer pooling rows with the confidence weights
   pose heads on the pooled (b,1024) feature (dcl_pose_head) -> 9-D -> svd3_project, 3-D translation
"""
import ctypes

import torch

from . import _lib as L
from .modules import fda_from_workspaces

_EPS_NAMES = ("Xc_p1", "Xc_m1", "Xc_p2", "Xc_m2", "Yo_p1", "Yo_m1", "Yo_p2", "Yo_m2")


def pm_bytes(rows, c, fmt=0):
    return rows * c * (2 if fmt == L.FMT_F16 else 4)


def pm_empty(rows, c, device, fmt=0):
    return torch.empty(pm_bytes(rows, c, fmt), dtype=torch.uint8, device=device)


_pm_empty = pm_empty


def pm_pack_rows(x, fmt=0):
    """fp32 (R, C) row-major -> PM image (fmt 0) / PM16 image (fmt 1); R % 128 == 0, C % 32 == 0."""
    x = x.contiguous()
    rows, c = x.shape
    out = pm_empty(rows, c, x.device, fmt)
    L.check(L.load().dcl_pm_pack_rows(rows, c, c, L.ptr(x), L.ptr(out), fmt, L.stream_ptr()), "pm_pack_rows")
    return out


def pm_pack_cm(x, fmt=0):
    """fp32 (B, C, N) channel-major -> PM / PM16 image of the (B*N, C) activation."""
    x = x.contiguous()
    b, c, n = x.shape
    out = pm_empty(b * n, c, x.device, fmt)
    L.check(L.load().dcl_pm_pack_cm(b, c, n, L.ptr(x), L.ptr(out), fmt, L.stream_ptr()), "pm_pack_cm")
    return out


def pm_unpack(pm, rows, c, fmt=0):
    out = torch.empty(rows, c, dtype=torch.float32, device=pm.device)
    L.check(L.load().dcl_pm_unpack(rows, c, L.ptr(pm), L.ptr(out), fmt, L.stream_ptr()), "pm_unpack")
    return out


def pick_nt(cout):
    for nt in (256, 128, 64):
        if cout % nt == 0:
            return nt
    raise ValueError(f"pm_gemm: cout={cout} is not a multiple of 64")


def pack_weight(w, nt, fmt=0):
    """(cout, cin) fp32 -> packed hi/lo blobs, one per (n-tile, k-block of 32): bf16 halves (fmt 0) or fp16 (fmt 1)."""
    cout, cin = w.shape
    assert cout % nt == 0 and cin % 32 == 0
    dt = torch.float16 if fmt == L.FMT_F16 else torch.bfloat16
    hi = w.to(dt)
    lo = (w - hi.float()).to(dt)

    def img(x):  # (tile, rg, r, kb, ch, e) -> (tile, kb, rg, ch, r, e)
        return x.view(cout // nt, nt // 8, 8, cin // 32, 4, 8).permute(0, 3, 1, 4, 2, 5)
    packed = torch.stack([img(hi), img(lo)], dim=2).contiguous()  # (tile, kb, half, rg, ch, r, e)
    return packed.view(torch.uint8).reshape(-1)


class Layer:
    """One packed GEMM layer: y = post(relu(x W^T + bias))."""

    def __init__(self, w, bias, relu, post_scale=None, post_shift=None, fmt=0):
        self.cout, self.cin = w.shape
        self.nt = pick_nt(self.cout)
        self.fmt = fmt                 # format of the A operand this layer reads (and of its packed weights)
        self.w = pack_weight(w.float().contiguous(), self.nt, fmt)
        self.bias = None if bias is None else bias.float().contiguous()
        self.relu = int(relu)
        self.post_scale = None if post_scale is None else post_scale.float().contiguous()
        self.post_shift = None if post_shift is None else post_shift.float().contiguous()


def _bn_affine(bn):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s, bn.bias - bn.running_mean * s


def layers_from_disengage(stack, fmt=0):
    """nn.Sequential of two BasicBlock_3DCONV (Conv3d 1x1x1 no bias -> BN3d -> ReLU): BN folded into the conv."""
    out = []
    for block in stack:
        conv, bn = block.layers[0], block.layers[1]
        s, t = _bn_affine(bn)
        w = conv.weight.reshape(conv.out_channels, conv.in_channels) * s[:, None]
        out.append(Layer(w, t, relu=True, fmt=fmt))
    return out


def layers_from_head(head, fmt=0):
    """Head_MultiLayerPerceptron: Conv1d(k=1) -> [ReLU] -> [BN]; returns (gemm layers, trailing torch convs).
    Layers whose width is not a multiple of 64 (the final 1/3/9-wide ones) stay in torch."""
    mods = list(head.layers)
    gemm, rest, i = [], [], 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, torch.nn.Conv1d)
        relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
        j = i + 1 + int(relu)
        bn = mods[j] if j < len(mods) and isinstance(mods[j], torch.nn.BatchNorm1d) else None
        j += int(bn is not None)
        if conv.out_channels % 64 == 0 and conv.in_channels % 32 == 0 and not rest:
            ps, pt = _bn_affine(bn) if bn is not None else (None, None)
            gemm.append(Layer(conv.weight.reshape(conv.out_channels, conv.in_channels), conv.bias, relu, ps, pt, fmt))
        else:
            rest.append((conv, relu, bn))
        i = j
    return gemm, rest


def _addr(x):
    """Device address of a tensor, or a raw integer address (a sub-range of a workspace), or None."""
    return x if x is None or isinstance(x, int) else L.ptr(x)


# bench.py sets this to a list to collect (start, end, algorithmic flops, tile width) around every dcl_pm_gemm launch.
GEMM_EVENTS = None


def run_gemm(problems, rows):
    """problems: list of dicts with keys a0, [a1], layer, [out_pm], [out_cm], [rows_per_inst], [pool_w], [pool_out]."""
    rows_ = []
    for p in problems:
        lay = p["layer"]
        a1 = p.get("a1")
        kb_total = lay.cin // 32
        c0 = p.get("c0", lay.cin if a1 is None else None)
        inst = p.get("inst", (0, 0, 0, 0))
        # strided batch (weight-gradient launches of train_tail.py): inst = (slices, a / w / out_cm strides in bytes)
        rows_.append({"a0": p["a0"], "a1": a1, "kb0": c0 // 32, "kb_total": kb_total, "w": lay.w, "bias": lay.bias,
                      "post_scale": lay.post_scale, "post_shift": lay.post_shift, "relu": lay.relu, "cout": lay.cout,
                      "nt": lay.nt, "out_pm": p.get("out_pm"), "out_cm": p.get("out_cm"),
                      "rows_per_inst": p.get("rows_per_inst", 0), "pool_w": p.get("pool_w"), "pool_out": p.get("pool_out"),
                      "dot_w": p.get("dot_w"), "dot_out": p.get("dot_out"), "out_qk": p.get("out_qk"),
                      "qk_tile_rows": p.get("qk_tile_rows", 0), "out_v": p.get("out_v"), "v_row0": p.get("v_row0", 0),
                      "v_rows": p.get("v_rows", 0), "a_fmt": lay.fmt, "out_fmt": p.get("out_fmt", lay.fmt),
                      "inst_count": inst[0], "a_inst_stride": inst[1], "w_inst_stride": inst[2],
                      "out_cm_inst_stride": inst[3]})
    arr = L.fill_structs(L.PmGemmProblem, rows_)
    if GEMM_EVENTS is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(L.load().dcl_pm_gemm(len(problems), ctypes.cast(arr, ctypes.c_void_p), rows, L.stream_ptr()), "pm_gemm")
    if GEMM_EVENTS is not None:
        ev1.record()
        flops = sum(2.0 * rows * p["layer"].cin * p["layer"].cout for p in problems)
        GEMM_EVENTS.append((ev0, ev1, flops, problems[0]["layer"].nt))


class FusedTail:
    """Packed-weight inference path of a dcl_net.Network (eval mode, test mode, no autograd)."""

    keep_debug = False  # True: also materialise F_Xo_m / F_Yc_p / F_Yc_m in the reference layout (tests)

    def __init__(self, net, fmt=0):
        self.net = net
        self.fmt = fmt
        self.c_m = net.disengage_Xc_m1[1].layers[0].out_channels
        with torch.no_grad():
            self.dis = {name: layers_from_disengage(getattr(net, "disengage_" + name), fmt) for name in _EPS_NAMES}
            self.conf, self.conf_rest = layers_from_head(net.regressor_conf, fmt)
            self.conf_bi, self.conf_bi_rest = layers_from_head(net.regressor_conf_bi, fmt)
            self.fuser, rest_a = layers_from_head(net.neck_fuser, fmt)
            self.fuser_bi, rest_b = layers_from_head(net.neck_fuser_bi, fmt)
        assert not rest_a and not rest_b and len(self.fuser) == 3 and len(self.conf) == 2
        # train-mode outputs evaluated without autograd (Xo_pred / Yc_pred, models/DCL_Net.py:207-210): the
        # 256 -> 256 -> 128 layers as they are, the trailing 128 -> 3 layer zero-padded to one 64-wide n-tile
        self.coord = {}
        if net.mode != "test":
            with torch.no_grad():
                for key, head in (("Xo", net.regressor_Xo), ("Yc", net.regressor_Yc)):
                    gemm, rest = layers_from_head(head, fmt)
                    (conv, relu, bn), = rest
                    assert len(gemm) == 2 and conv.out_channels <= 64 and not relu and bn is None
                    w = conv.weight.new_zeros(64, conv.in_channels)
                    w[:conv.out_channels] = conv.weight.reshape(conv.out_channels, conv.in_channels)
                    bias = conv.bias.new_zeros(64)
                    bias[:conv.out_channels] = conv.bias
                    self.coord[key] = (gemm + [Layer(w, bias, relu=False, fmt=fmt)], conv.out_channels)
        self.conf_dot, self.conf_dot_bias = [], []
        for rest in (self.conf_rest, self.conf_bi_rest):
            (conv, relu, bn), = rest
            assert conv.out_channels == 1 and not relu and bn is None
            self.conf_dot.append(conv.weight.detach().reshape(-1).float().contiguous())
            self.conf_dot_bias.append(conv.bias.detach().float().reshape(1, 1).clone())

    @staticmethod
    def unsupported_reason(net, b):
        """None when the packed path takes this network / batch, else a short reason."""
        c_m = net.disengage_Xc_m1[1].layers[0].out_channels
        if net.training:
            return "module is in training mode (BatchNorm uses batch statistics)"
        if net.n_inp != net.n_tmp or net.n_inp % 128 != 0:
            return f"n_inp={net.n_inp}, n_tmp={net.n_tmp}: need equal sizes that are multiples of 128"
        if c_m not in (64, 128):
            return f"c_m={c_m}: the fused FDA kernel takes 64 or 128"
        return None

    @staticmethod
    def supported(net, b):
        return FusedTail.unsupported_reason(net, b) is None

    @staticmethod
    def pick_fmt(net):
        """PM16 when the network asks for it and the fp16 form of the fused FDA kernel applies (CTA pairs: an even
        number of 128-query tiles per instance); the bf16 hi/lo format otherwise."""
        want16 = getattr(net, "precision", "fp16") == "fp16"
        return L.FMT_F16 if want16 and (net.n_inp // 128) % 2 == 0 else L.FMT_BF16X2

    @torch.no_grad()
    def forward(self, pm_xc, pm_yo, b):
        net, c_m = self.net, self.c_m
        n = net.n_inp
        rows = b * n
        dev = pm_xc.device
        f32 = dict(dtype=torch.float32, device=dev)
        fmt = self.fmt
        pm_empty = lambda r, c, d: _pm_empty(r, c, d, fmt)      # every activation image of this pass has format `fmt`

        # ---- disengage layer 1: eight 480 -> 256 problems in one launch
        h1 = {name: pm_empty(rows, 256, dev) for name in _EPS_NAMES}
        run_gemm([{"a0": pm_xc if name.startswith("Xc") else pm_yo, "layer": self.dis[name][0], "out_pm": h1[name]}
                  for name in _EPS_NAMES], rows)

        # ---- disengage layer 2: outputs in the formats their consumers read — point-major images for the MLPs, and
        # the query / key / value operand images of the two FDA launches, written by the GEMM epilogue straight into
        # the FDA workspaces (no channel-major fp32 round trip, no pack pass)
        lib = L.load()
        ws_bytes = lib.dcl_fda_workspace_bytes(b, c_m, 256, n, n)
        offs = (ctypes.c_size_t * 3)()
        L.check(lib.dcl_fda_workspace_layout(b, c_m, 256, n, n, ctypes.cast(offs, ctypes.c_void_p)), "fda layout")
        ws = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]   # [Xc->Yo direction, Yo->Xc]
        q = [w.data_ptr() + offs[0] for w in ws]
        k = [w.data_ptr() + offs[1] for w in ws]
        v = [w.data_ptr() + offs[2] for w in ws]
        vrows = 256 + c_m
        pm_out = {name: pm_empty(rows, 256 if "_p" in name else c_m, dev) for name in ("Xc_p1", "Yo_p2", "Xc_m1", "Yo_m2")}
        fda_out = {
            "Xc_p2": {"out_v": v[1], "v_row0": 0, "v_rows": vrows},
            "Yo_p1": {"out_v": v[0], "v_row0": 0, "v_rows": vrows},
            "Xc_m1": {"out_qk": q[0], "qk_tile_rows": 128},
            "Yo_m2": {"out_qk": q[1], "qk_tile_rows": 128},
            "Yo_m1": {"out_qk": k[0], "qk_tile_rows": 64, "out_v": v[0], "v_row0": 256, "v_rows": vrows},
            "Xc_m2": {"out_qk": k[1], "qk_tile_rows": 64, "out_v": v[1], "v_row0": 256, "v_rows": vrows},
        }
        for group in (("Xc_p1", "Xc_p2", "Yo_p1", "Yo_p2"), ("Xc_m1", "Xc_m2", "Yo_m1", "Yo_m2")):
            run_gemm([dict({"a0": h1[name], "layer": self.dis[name][1], "out_pm": pm_out.get(name)},
                           **fda_out.get(name, {})) for name in group], rows)
        del h1

        # ---- dual FDA: both directions in ONE launch (both attention products of a direction in one pass over the
        # keys); the aligned features leave the
        # kernel as point-major images for the MLPs below, F_Xo_p also in the reference's layout (stage 2 reads it)
        dbg = self.keep_debug
        (F_Xo_p, F_Xo_m, pm_Xo_p, pm_Xo_m, _), (F_Yc_p, F_Yc_m, pm_Yc_p, pm_Yc_m, _) = fda_from_workspaces(
            [(ws[0], True, dbg, True, True, False), (ws[1], dbg, dbg, True, True, False)], b, c_m, n, n, pv_fmt=fmt)
        del ws

        # ---- coordinate regressors of the train-mode interface (regressor_Xo on F_Xo_p, regressor_Yc on F_Yc_p)
        coords = {}
        if self.coord:
            (lx, dx), (ly, dy) = self.coord["Xo"], self.coord["Yc"]
            t1 = [pm_empty(rows, lx[0].cout, dev) for _ in range(2)]
            run_gemm([{"a0": pm_Xo_p, "layer": lx[0], "out_pm": t1[0]}, {"a0": pm_Yc_p, "layer": ly[0], "out_pm": t1[1]}], rows)
            t2 = [pm_empty(rows, lx[1].cout, dev) for _ in range(2)]
            run_gemm([{"a0": t1[0], "layer": lx[1], "out_pm": t2[0]}, {"a0": t1[1], "layer": ly[1], "out_pm": t2[1]}], rows)
            xyz = [torch.empty(b, 64, n, **f32) for _ in range(2)]
            run_gemm([{"a0": t2[0], "layer": lx[2], "out_cm": xyz[0], "rows_per_inst": n},
                      {"a0": t2[1], "layer": ly[2], "out_cm": xyz[1], "rows_per_inst": n}], rows)
            coords = {"Xo_pred": xyz[0][:, :dx].transpose(1, 2), "Yc_pred": xyz[1][:, :dy].transpose(1, 2)}

        # ---- confidence heads: cat([F_Xc_m1, F_Xo_m]) / cat([F_Yc_m, F_Yo_m2]) -> 128 -> 128 -> 1
        c1 = [pm_empty(rows, 128, dev) for _ in range(2)]
        run_gemm([{"a0": pm_out["Xc_m1"], "a1": pm_Xo_m, "c0": c_m, "layer": self.conf[0], "out_pm": c1[0]},
                  {"a0": pm_Yc_m, "a1": pm_out["Yo_m2"], "c0": c_m, "layer": self.conf_bi[0], "out_pm": c1[1]}], rows)
        # the trailing 128 -> 1 convolution is a per-row dot product in the epilogue of the second layer
        logits = torch.empty(2, b, n, **f32)
        run_gemm([{"a0": c1[0], "layer": self.conf[1], "dot_w": self.conf_dot[0], "dot_out": logits[0]},
                  {"a0": c1[1], "layer": self.conf_bi[1], "dot_w": self.conf_dot[1], "dot_out": logits[1]}], rows)
        # sigmoid + softmax over the 2n correspondences (DCL_Net.py:219-220), bias of the dot layer added inside
        conf = torch.empty(b, 2 * n, **f32)
        w1, w2 = torch.empty(rows, **f32), torch.empty(rows, **f32)
        L.check(L.load().dcl_conf_weights(b, n, L.ptr(logits[0]), L.ptr(logits[1]), L.ptr(self.conf_dot_bias[0]),
                                          L.ptr(self.conf_dot_bias[1]), L.ptr(conf), L.ptr(w1), L.ptr(w2),
                                          L.stream_ptr()), "conf_weights")

        # ---- fusers: cat([F_Xc_p1, F_Xo_p]) / cat([F_Yc_p, F_Yo_p2]) -> 512 -> 512 -> 1024, pooled with conf_softmax
        f1 = [pm_empty(rows, 512, dev) for _ in range(2)]
        run_gemm([{"a0": pm_out["Xc_p1"], "a1": pm_Xo_p, "c0": 256, "layer": self.fuser[0], "out_pm": f1[0]},
                  {"a0": pm_Yc_p, "a1": pm_out["Yo_p2"], "c0": 256, "layer": self.fuser_bi[0], "out_pm": f1[1]}], rows)
        f2 = [pm_empty(rows, 512, dev) for _ in range(2)]
        run_gemm([{"a0": f1[0], "layer": self.fuser[1], "out_pm": f2[0]},
                  {"a0": f1[1], "layer": self.fuser_bi[1], "out_pm": f2[1]}], rows)
        parts = [torch.empty(rows // 32, 1024, **f32) for _ in range(2)]
        run_gemm([{"a0": f2[0], "layer": self.fuser[2], "pool_w": w1, "pool_out": parts[0]},
                  {"a0": f2[1], "layer": self.fuser_bi[2], "pool_w": w2, "pool_out": parts[1]}], rows)
        pooled = torch.empty(b, 1024, **f32)
        lib = L.load()
        L.check(lib.dcl_pm_pool_reduce(b, 1024, n // 32, L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(pooled), 0,
                                       L.stream_ptr()), "pool")
        F_p_wei = pooled.unsqueeze(-1)

        # ---- pose regressors on the pooled feature (b x 1024: tiny) and the SO(3) projection
        from .dcl_net import pose_heads, svd3_project
        ortho9d, trans = pose_heads(pooled, net.regressor_rot, net.regressor_trans)
        rot = svd3_project(ortho9d, True)
        return dict({"trans_pred": trans, "rot_pred": rot, "conf": conf, "F_Xo_p": F_Xo_p, "F_Xo_p_pm": pm_Xo_p,
                     "F_Xo_p_pm_fmt": fmt,
                     "_debug": {"F_Yc_p": F_Yc_p, "F_Xo_m": F_Xo_m, "F_Yc_m": F_Yc_m, "ortho9d": ortho9d}}, **coords)


class FusedRefiner:
    """Tensor-core inference path of refiner.Refiner (models/refiner.py:57-95) inside the stage-2 loop of
    tools/test_YCBV_stage2.py:204-225.  The refiner input cat([points_cano (3), F_Xo_p (256)]) is never built: the
    first layer reads X = [F_Xo_p image | 32-channel image of the canonicalised points] (its weight columns are
    permuted and zero-padded to match), the points image being rewritten by dcl_pose_compose_pm every iteration;
    the last layer's epilogue does the confidence-weighted pooling."""

    def __init__(self, refiner, fmt=0):
        self.refiner = refiner
        self.fmt = fmt
        convs = [m for m in refiner.MLP_share.layers if isinstance(m, torch.nn.Conv1d)]
        others = [m for m in refiner.MLP_share.layers if not isinstance(m, (torch.nn.Conv1d, torch.nn.ReLU))]
        assert len(convs) == 3 and not others and convs[0].in_channels == 259
        with torch.no_grad():
            w1 = convs[0].weight.reshape(convs[0].out_channels, 259)
            w1p = w1.new_zeros(w1.shape[0], 288)
            w1p[:, :256] = w1[:, 3:]          # F_Xo_p channels first (8 k-blocks of the first image)
            w1p[:, 256:259] = w1[:, :3]       # then x, y, z (k-block 9, channels 3..31 are zero)
            if fmt == L.FMT_F16:
                # PM16 points image: channels 3-5 hold the fp16 remainder of the coordinates (dcl_pose_compose_pm16),
                # so the coordinates enter the product with ~22 bits: the same weight columns once more
                w1p[:, 259:262] = w1[:, :3]
            self.layers = [Layer(w1p, convs[0].bias, True, fmt=fmt),
                           Layer(convs[1].weight.reshape(convs[1].out_channels, -1), convs[1].bias, True, fmt=fmt),
                           Layer(convs[2].weight.reshape(convs[2].out_channels, -1), convs[2].bias, True, fmt=fmt)]

    @staticmethod
    def supported(refiner, b, n, F_Xo_p=None, conf=None, pm_feat=None):
        """The pooling epilogue writes one partial per 32 global rows and dcl_pm_pool_reduce folds n/32 of them per
        instance, so an instance must be a whole number of 128-row tiles; the feature must be 256 wide and the
        confidence row must cover the n observed points."""
        if refiner.training or torch.is_grad_enabled() or n % 128 != 0 or (b * n) % 256 != 0:
            return False
        if F_Xo_p is None and pm_feat is None:
            return False
        if F_Xo_p is not None and (F_Xo_p.shape[1] != 256 or F_Xo_p.shape[2] != n):
            return False
        return conf is not None and conf.shape[1] >= n

    @torch.no_grad()
    def refine(self, points_inp, rot_pred, trans_pred, F_Xo_p, conf, iteration, pm_feat=None):
        """pm_feat: the point-major image of F_Xo_p in THIS object's format (stage 1 returns it with its format in
        prediction["F_Xo_p_pm_fmt"]); packed here from F_Xo_p when absent."""
        from .dcl_net import pose_heads, svd3_project
        lib = L.load()
        fmt = self.fmt
        pm_empty = lambda r, c, d: _pm_empty(r, c, d, fmt)
        b, n, _ = points_inp.shape
        rows, dev = b * n, points_inp.device
        points_inp = points_inp.contiguous()
        rot, trans = rot_pred.clone().contiguous(), trans_pred.clone().contiguous()
        if pm_feat is None:
            pm_feat = pm_pack_cm(F_Xo_p, fmt)
        pm_pts = torch.zeros(pm_bytes(rows, 32, fmt), dtype=torch.uint8, device=dev)
        pool_w = torch.softmax(conf, dim=1)[:, :n].reshape(-1).contiguous()     # refiner.py:60
        st = L.stream_ptr
        compose_fn = lib.dcl_pose_compose_pm16 if fmt == L.FMT_F16 else lib.dcl_pose_compose_pm

        def compose(dR, dt):
            L.check(compose_fn(b, n, L.ptr(rot), L.ptr(trans), L.ptr(dR), L.ptr(dt), L.ptr(points_inp), None, 0,
                               L.ptr(pm_pts), st()), "pose_compose")

        compose(None, None)
        l1, l2, l3 = self.layers
        for _ in range(iteration):
            h1, h2 = pm_empty(rows, l1.cout, dev), pm_empty(rows, l2.cout, dev)
            run_gemm([{"a0": pm_feat, "a1": pm_pts, "c0": 256, "layer": l1, "out_pm": h1}], rows)
            run_gemm([{"a0": h1, "layer": l2, "out_pm": h2}], rows)
            parts = torch.empty(rows // 32, l3.cout, dtype=torch.float32, device=dev)
            run_gemm([{"a0": h2, "layer": l3, "pool_w": pool_w, "pool_out": parts}], rows)
            pooled = torch.empty(b, l3.cout, dtype=torch.float32, device=dev)
            L.check(lib.dcl_pm_pool_reduce(b, l3.cout, n // 32, L.ptr(parts), None, L.ptr(pooled), 0, st()), "pool")
            ortho9d, dt = pose_heads(pooled, self.refiner.regressor_rot2, self.refiner.regressor_trans2)
            dR = svd3_project(ortho9d, True)
            compose(dR.contiguous(), dt.contiguous())
        return rot, trans
