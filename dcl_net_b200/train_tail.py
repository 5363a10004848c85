"""Training path of the pointwise MLP stacks of Network.forward (models/DCL_Net.py:56-151,187-235 in train mode):
forward AND backward on the tcgen05 GEMM of csrc/pm_gemm.cu instead of the cuDNN/cuBLAS calls autograd makes for
nn.Conv3d / nn.Conv1d / nn.BatchNorm (reference step: tools/train_YCBV_stage1.py:168-191).

`mlp_stacks(stacks)` runs S parallel stacks of pointwise layers as ONE autograd node.  Per layer, three GEMMs on
bf16 hi/lo operand images (3 MMAs per product: fp32-faithful, so gradients need no loss scaling):

    forward   U  = X W^T (+ bias, ReLU)      dcl_pm_gemm, A = PM image of X
    dgrad     dX = dZ W                      dcl_pm_gemm, A = PM image of dZ, packed W^T as the weights
    wgrad     dW = sum_b dZ_b^T X_b          dcl_pm_gemm strided batch over the per-instance transposed images

and the HBM-bound passes of csrc/train_ops.cu around them (train-mode BatchNorm statistics, the pointwise
transforms and their backward, operand-image writers, deterministic bias / weight gradient reductions).  The module
tree, parameter names and BatchNorm running statistics are the reference's; only the kernels differ.

Layer kinds (what follows the 1x1 convolution):
    "linear"   conv (+bias)                               last layer of the regressors
    "relu"     conv + bias -> ReLU                        Head_MultiLayerPerceptron without BatchNorm
    "bn_relu"  conv -> BatchNorm -> ReLU                  BasicBlock_3DCONV (disengage blocks)
    "relu_bn"  conv + bias -> ReLU -> BatchNorm           Head_MultiLayerPerceptron with BatchNorm (neck fusers)
"""
import ctypes
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib as L
from .fused_tail import pick_nt, pm_empty, run_gemm

TR_COPY, TR_AFFINE, TR_AFFINE_RELU, TR_BWD_RELU, TR_BWD_BN_RELU, TR_BWD_RELU_BN = range(6)
_FWD_MODE = {"linear": TR_COPY, "relu": TR_COPY, "bn_relu": TR_AFFINE_RELU, "relu_bn": TR_AFFINE}
_BWD_MODE = {"linear": TR_COPY, "relu": TR_BWD_RELU, "bn_relu": TR_BWD_BN_RELU, "relu_bn": TR_BWD_RELU_BN}


def _pad(x, m):
    return (x + m - 1) // m * m


def _wgrad_group(b):
    """Instances per split-K slice of the weight-gradient GEMMs (K = group * n points)."""
    return next(g for g in (4, 2, 1) if b % g == 0)


def _batched(fn_name, struct, items, what):
    """Launch `items` (lists of field dicts) through a dcl_tr_* entry point, 8 per launch."""
    fn = getattr(L.load(), fn_name)
    st = L.stream_ptr()
    for i in range(0, len(items), 8):
        arr = L.fill_structs(struct, items[i:i + 8])
        L.check(fn(len(arr), ctypes.cast(arr, ctypes.c_void_p), st), what)


def tile_pass(items):
    """csrc/train_ops.cu:tr_tile_kernel.  items: dicts with x (tensor), fmt ('cm' (b,c,n) view | 'rm' (b*n,c) matrix),
    b, c, n, mode and the optional fields of dcl_tr_tile."""
    out = []
    for it in items:
        x, b, c, n = it["x"], it["b"], it["c"], it["n"]
        if it["fmt"] == "cm":
            assert x.shape == (b, c, n) and x.stride(2) == 1
            sb, sc, sn = x.stride(0), x.stride(1), 1
        else:
            assert x.shape == (b * n, c) and x.stride(1) == 1
            sb, sc, sn = n * x.stride(0), 1, x.stride(0)
        f = {"x": x.data_ptr(), "x_sb": sb, "x_sc": sc, "x_sn": sn, "b": b, "c": c, "n": n,
             "mode": it["mode"], "t_row0": it.get("t_row0", 0), "t_rows": it.get("t_rows", 0),
             "k_col0": it.get("k_col0", 0), "k_cols": it.get("k_cols", 0), "t_group": it.get("t_group", 0)}
        for k in ("u", "scale", "shift", "mean", "rstd", "s1", "s2", "out_k", "out_t", "out_cm", "col_partial"):
            f[k] = it.get(k)
        out.append(f)
    _batched("dcl_tr_tile_pass", L.TrTile, out, "tr_tile")


def pack_weights(items):
    """items: (src (rows, cols) fp32 contiguous — or its transpose when `transpose` —, rows, cols, rows_pad, k_pad, nt,
    transpose[, dst, k_col0, k_total]) -> list of packed uint8 tensors.  With dst / k_col0 / k_total the matrix becomes
    a column block of a wider packed matrix (weights concatenated along the reduction axis)."""
    outs, fields = [], []
    for src, rows, cols, rows_pad, k_pad, nt, transpose, *into in items:
        dst, k_col0, k_total = into if into else (None, 0, 0)
        if dst is None:
            dst = torch.empty(rows_pad * k_pad * 4, dtype=torch.uint8, device=src.device)
        outs.append(dst)
        fields.append({"src": src, "dst": dst, "rows": rows, "cols": cols, "rows_pad": rows_pad, "k_pad": k_pad,
                       "nt": nt, "transpose": int(transpose), "k_col0": k_col0, "k_total": k_total})
    _batched("dcl_tr_pack_weights", L.TrWpack, fields, "tr_pack_weights")
    return outs


def _gemm_layer(w_packed, cout, cin, nt, bias=None, relu=False):
    return SimpleNamespace(cout=cout, cin=cin, nt=nt, fmt=L.FMT_BF16X2, w=w_packed, bias=bias, relu=int(relu),
                           post_scale=None, post_shift=None)


def _run_gemm_groups(problems, rows):
    """dcl_pm_gemm takes up to 8 problems of equal (cout, nt) per launch."""
    groups = {}
    for p in problems:
        groups.setdefault((p["layer"].cout, p["layer"].nt), []).append(p)
    for plist in groups.values():
        for i in range(0, len(plist), 8):
            run_gemm(plist[i:i + 8], rows)


# Tests set this to a list to record the ReLU gates of a forward pass: entries (layer, stack, bool mask (b, c, n),
# the nn.ReLU module of the layer or None).
GATE_LOG = None


class StackSpec:
    """One stack: inputs = [(tensor, 'cm' | 'rm')] (one tensor, or two concatenated along channels), layers =
    [(kind, weight (cout, cin), bias | None, bn module | None[, the layer's nn.ReLU module — only for GATE_LOG])]."""

    def __init__(self, inputs, layers):
        self.inputs, self.layers = inputs, layers


class _MlpStacksFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *tensors):
        b, n = plan.b, plan.n
        rows = b * n
        dev = tensors[0].device
        f32 = dict(dtype=torch.float32, device=dev)
        nl = plan.nlayers
        # ---- K-images of the distinct input tensors
        in_imgs = {}
        items = []
        for st in plan.stacks:
            for ti, fmt, c in st.inputs:
                if ti not in in_imgs:
                    in_imgs[ti] = pm_empty(rows, c, dev)
                    items.append({"x": tensors[ti], "fmt": fmt, "b": b, "c": c, "n": n, "mode": TR_COPY,
                                  "out_k": in_imgs[ti]})
        tile_pass(items)
        cur = [[in_imgs[ti] for ti, _, _ in st.inputs] for st in plan.stacks]   # per stack: the A images
        cur_c0 = [st.inputs[0][2] for st in plan.stacks]
        saved_u, saved_bn, saved_t = [], [], []
        grp = _wgrad_group(b)
        for l in range(nl):
            lays = [st.layers[l] for st in plan.stacks]
            wp = pack_weights([(tensors[la.w], la.cout, la.cin, la.cout, la.cin, pick_nt(la.cout), False) for la in lays])
            last = l == nl - 1
            us, nxt, probs = [], [], []
            for s, (la, w) in enumerate(zip(lays, wp)):
                u = torch.empty(b, la.cout, n, **f32)
                bias = tensors[la.bias] if la.bias is not None else None
                lay = _gemm_layer(w, la.cout, la.cin, pick_nt(la.cout), bias, la.kind in ("relu", "relu_bn"))
                p = {"a0": cur[s][0], "layer": lay, "out_cm": u, "rows_per_inst": n}
                if len(cur[s]) == 2:
                    p.update(a1=cur[s][1], c0=cur_c0[s])
                if la.kind in ("linear", "relu") and not last:
                    p["out_pm"] = pm_empty(rows, la.cout, dev)     # Y = U: the next layer's operand straight away
                    nxt.append([p["out_pm"]])
                else:
                    nxt.append(None)
                us.append(u)
                probs.append(p)
            _run_gemm_groups(probs, rows)
            # ---- train-mode BatchNorm: batch statistics, then the affine (+ReLU) pass writing the next operand
            bn_items, tiles, bns = [], [], []
            for s, la in enumerate(lays):
                if la.kind not in ("bn_relu", "relu_bn"):
                    bns.append(None)
                    continue
                bn = la.bn
                stat = torch.empty(4, la.cout, **f32)            # mean, rstd, scale, shift
                upd = bn.training and bn.track_running_stats
                bn_items.append({"u": us[s], "b": b, "c": la.cout, "n": n, "gamma": tensors[la.gamma],
                                 "beta": tensors[la.beta], "eps": float(bn.eps),
                                 "momentum": float(bn.momentum if bn.momentum is not None else 0.0),
                                 "running_mean": bn.running_mean if upd else None,
                                 "running_var": bn.running_var if upd else None,
                                 "mean": stat[0], "rstd": stat[1], "scale": stat[2], "shift": stat[3]})
                bns.append(stat)
                t = {"x": us[s], "fmt": "cm", "b": b, "c": la.cout, "n": n, "mode": _FWD_MODE[la.kind],
                     "scale": stat[2], "shift": stat[3]}
                if last or (GATE_LOG is not None and la.kind == "bn_relu"):
                    t["out_cm"] = torch.empty(b, la.cout, n, **f32)
                if not last:
                    t["out_k"] = pm_empty(rows, la.cout, dev)
                    nxt[s] = [t["out_k"]]
                    # the same pass writes the transposed images the NEXT layer's weight gradient reads (kept for the
                    # backward: one read of U saved per activation)
                    cpad = _pad(la.cout, 128)
                    alloc = torch.empty if cpad == la.cout else torch.zeros
                    t.update(out_t=alloc(b * cpad * n * 4, dtype=torch.uint8, device=dev), t_rows=cpad, t_group=grp)
                tiles.append((s, t))
            if bn_items:
                _batched("dcl_tr_bn_stats", L.TrBn, bn_items, "tr_bn_stats")
                tile_pass([t for _, t in tiles])
            saved_u.append(us)
            saved_bn.append(bns)
            t_of = {s: t.get("out_t") for s, t in tiles}
            saved_t.append([t_of.get(s) for s in range(len(lays))])
            if GATE_LOG is not None:
                ys = {s: t["out_cm"] for s, t in tiles if "out_cm" in t}
                for s, la in enumerate(lays):
                    if la.kind != "linear":
                        GATE_LOG.append((l, s, (ys[s] if la.kind == "bn_relu" else us[s]) > 0, la.relu_mod))
            if last:
                outs = list(us)
                for s, t in tiles:
                    outs[s] = t["out_cm"]
            else:
                cur = nxt
                cur_c0 = [la.cout for la in lays]
        for st in plan.stacks:
            for la in st.layers:
                if la.bn is not None and la.bn.training and la.bn.track_running_stats:
                    la.bn.num_batches_tracked += 1
        ctx.plan = plan
        # everything the backward reads goes through save_for_backward (outputs included: no reference cycles)
        flat, ctx.u_idx, ctx.bn_idx, ctx.t_idx = list(tensors), [], [], []
        for us, bns, ts in zip(saved_u, saved_bn, saved_t):
            ctx.u_idx.append([len(flat) + i for i in range(len(us))])
            flat += us
            for dst, src in ((ctx.bn_idx, bns), (ctx.t_idx, ts)):
                row = []
                for st in src:
                    row.append(None if st is None else len(flat))
                    if st is not None:
                        flat.append(st)
                dst.append(row)
        ctx.ntensors = len(tensors)
        ctx.save_for_backward(*flat)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        plan, flat = ctx.plan, ctx.saved_tensors
        tensors = flat[:ctx.ntensors]
        saved_u = [[flat[i] for i in row] for row in ctx.u_idx]
        saved_bn = [[None if i is None else flat[i] for i in row] for row in ctx.bn_idx]
        saved_t = [[None if i is None else flat[i] for i in row] for row in ctx.t_idx]
        b, n = plan.b, plan.n
        rows = b * n
        dev = tensors[0].device
        f32 = dict(dtype=torch.float32, device=dev)
        nl, ns = plan.nlayers, len(plan.stacks)
        lib = L.load()
        # wgrad split-K: one slice per group of `grp` instances (K = grp*n points): partial sums to write and reduce
        # shrink by grp while the launch still has several waves of tiles
        grp = _wgrad_group(b)
        out_grads = [None] * len(tensors)

        def acc(idx, g):
            out_grads[idx] = g if out_grads[idx] is None else out_grads[idx] + g

        dys = []
        for s, g in enumerate(grads):
            c = plan.stacks[s].layers[-1].cout
            dys.append(torch.zeros(b, c, n, **f32) if g is None else (g if g.stride(2) == 1 else g.contiguous()))
        for l in range(nl - 1, -1, -1):
            lays = [st.layers[l] for st in plan.stacks]
            us, bns = saved_u[l], saved_bn[l]
            need_dx = l > 0 or any(ctx.needs_input_grad[1 + ti] for st in plan.stacks for ti, _, _ in st.inputs)
            # ---- BatchNorm backward sums (= d beta, d gamma)
            red, sums = [], [None] * ns
            for s, la in enumerate(lays):
                if bns[s] is None:
                    continue
                sums[s] = torch.empty(2, la.cout, **f32)
                dy = dys[s]
                assert dy.stride(2) == 1
                red.append({"dy": dy.data_ptr(), "u": us[s], "dy_sb": dy.stride(0), "dy_sc": dy.stride(1),
                            "b": b, "c": la.cout, "n": n, "mode": _BWD_MODE[la.kind], "mean": bns[s][0], "rstd": bns[s][1],
                            "scale": bns[s][2], "shift": bns[s][3], "s1": sums[s][0], "s2": sums[s][1]})
            if red:
                _batched("dcl_tr_bn_bwd_reduce", L.TrBnBwd, red, "tr_bn_bwd_reduce")
            # ---- dZ: PM image (dgrad), transposed images (wgrad), bias-gradient partials
            # layer 0: stacks reading the same input share ONE input-gradient GEMM, dX = [dZ_1 | dZ_2 | ..] [W_1; W_2; ..]
            # (their dZ images are column blocks of one image), instead of one GEMM each and a sum of the results
            share = {}
            if l == 0 and need_dx:
                for s, st in enumerate(plan.stacks):
                    share.setdefault(tuple(ti for ti, _, _ in st.inputs), []).append(s)
                share = {k: v for k, v in share.items() if len(v) > 1}
            cat_img, cat_of = {}, {}
            for key, members in share.items():
                ktot = sum(lays[s].cout for s in members)
                img = pm_empty(rows, ktot, dev)
                col = 0
                for s in members:
                    cat_of[s] = (key, col, ktot)
                    col += lays[s].cout
                cat_img[key] = img
            items, dz_k, dz_t, colp = [], [], [], []
            for s, la in enumerate(lays):
                cp_rows = _pad(la.cout, 128)
                alloc = torch.empty if cp_rows == la.cout else torch.zeros
                dzt = alloc(b * cp_rows * n * 4, dtype=torch.uint8, device=dev)
                it = {"x": dys[s], "fmt": "cm", "b": b, "c": la.cout, "n": n, "mode": _BWD_MODE[la.kind],
                      "out_t": dzt, "t_rows": cp_rows, "t_group": grp}
                if la.kind != "linear":
                    it["u"] = us[s]
                if bns[s] is not None:
                    it.update(scale=bns[s][2], shift=bns[s][3], mean=bns[s][0], rstd=bns[s][1], s1=sums[s][0], s2=sums[s][1])
                if need_dx and s in cat_of:
                    key, col, ktot = cat_of[s]
                    it.update(out_k=cat_img[key], k_col0=col, k_cols=ktot)
                elif need_dx:
                    it["out_k"] = pm_empty(rows, la.cout, dev)
                if la.bias is not None:
                    it["col_partial"] = torch.empty(rows // 128, la.cout, **f32)
                items.append(it)
                dz_k.append(it.get("out_k"))
                dz_t.append((dzt, cp_rows))
                colp.append(it.get("col_partial"))
            tile_pass(items)
            # ---- transposed images of the layer inputs (recomputed from what the forward saved)
            items, x_t, shared = [], [], {}
            for s, st in enumerate(plan.stacks):
                la = lays[s]
                cin_pad = _pad(la.cin, 128)
                key = tuple(ti for ti, _, _ in st.inputs)
                if l == 0 and key in shared:                  # stacks reading the same input share its image
                    x_t.append(shared[key])
                    continue
                if l > 0 and saved_t[l - 1][s] is not None:   # written by the forward's BatchNorm pass
                    x_t.append((saved_t[l - 1][s], cin_pad))
                    continue
                alloc = torch.empty if cin_pad == la.cin else torch.zeros
                xt = alloc(b * cin_pad * n * 4, dtype=torch.uint8, device=dev)
                x_t.append((xt, cin_pad))
                if l == 0:
                    shared[key] = (xt, cin_pad)
                    row0 = 0
                    for ti, fmt, c in st.inputs:
                        items.append({"x": tensors[ti], "fmt": fmt, "b": b, "c": c, "n": n, "mode": TR_COPY,
                                      "out_t": xt, "t_row0": row0, "t_rows": cin_pad, "t_group": grp})
                        row0 += c
                else:
                    pl = st.layers[l - 1]
                    it = {"x": saved_u[l - 1][s], "fmt": "cm", "b": b, "c": pl.cout, "n": n,
                          "mode": _FWD_MODE[pl.kind], "out_t": xt, "t_rows": cin_pad, "t_group": grp}
                    pbn = saved_bn[l - 1][s]
                    if pbn is not None:
                        it.update(scale=pbn[2], shift=pbn[3])
                    items.append(it)
            tile_pass(items)
            # ---- wgrad: per (stack, instance) slice  dW_b^T (cin_pad x cout_pad) = X_b^T-image x dZ_b^T-image, then the
            # fixed-order sum over instances
            # stacks of equal (cin_pad, cout_pad) share one partial buffer and one reduction launch
            shape_groups = {}
            for s, la in enumerate(lays):
                shape_groups.setdefault((x_t[s][1], dz_t[s][1]), []).append(s)
            probs, parts = [], {}
            for (cin_pad, cp_rows), members in shape_groups.items():
                part = torch.empty(len(members), b // grp, cin_pad, cp_rows, **f32)
                for k, s in enumerate(members):
                    parts[s] = (part, k)
                    lay = _gemm_layer(x_t[s][0], cin_pad, grp * n, 128)   # the X^T image read as packed weights (n-tile 128)
                    probs.append({"a0": dz_t[s][0], "layer": lay, "out_cm": part[k], "rows_per_inst": cp_rows,
                                  "inst": (b // grp, cp_rows * grp * n * 4, cin_pad * grp * n * 4, cin_pad * cp_rows * 4),
                                  "_rows": cp_rows})
            by_rows = {}
            for p in probs:
                by_rows.setdefault(p["_rows"], []).append(p)
            for r, plist in by_rows.items():
                _run_gemm_groups(plist, r)
            for (cin_pad, cp_rows), members in shape_groups.items():
                part = parts[members[0]][0]
                dwt = torch.empty(len(members), cin_pad, cp_rows, **f32)
                L.check(lib.dcl_pm_pool_reduce(len(members), cin_pad * cp_rows, b // grp, L.ptr(part), None, L.ptr(dwt), 0,
                                               L.stream_ptr()), "wgrad reduce")
                for k, s in enumerate(members):
                    la = lays[s]
                    acc(la.w, dwt[k, :la.cin, :la.cout].t().reshape(tensors[la.w].shape))
            sums_items = []
            for s, la in enumerate(lays):
                if la.bias is not None:
                    db = torch.empty(la.cout, **f32)
                    sums_items.append({"partial": colp[s], "out": db, "parts": rows // 128, "c": la.cout})
                    acc(la.bias, db)
            if sums_items:
                _batched("dcl_tr_colsum_reduce", L.TrColsum, sums_items, "tr_colsum")
            for s, la in enumerate(lays):
                if bns[s] is not None:
                    acc(la.beta, sums[s][0])
                    acc(la.gamma, sums[s][1])
            # ---- dgrad: dX (b, cin_pad, n) = dZ W
            if need_dx:
                solo = [s for s in range(ns) if s not in cat_of]
                wt = pack_weights([(tensors[lays[s].w], lays[s].cin, lays[s].cout, _pad(lays[s].cin, 64), lays[s].cout,
                                    pick_nt(_pad(lays[s].cin, 64)), True) for s in solo])
                probs, dxs = [], {}
                for s, w in zip(solo, wt):
                    la = lays[s]
                    cin_pad = _pad(la.cin, 64)
                    dxs[s] = torch.empty(b, cin_pad, n, **f32)
                    probs.append({"a0": dz_k[s], "layer": _gemm_layer(w, cin_pad, la.cout, pick_nt(cin_pad)),
                                  "out_cm": dxs[s], "rows_per_inst": n})
                shared_dx = {}
                for key, members in share.items():
                    la0 = lays[members[0]]
                    cin_pad, ktot = _pad(la0.cin, 64), cat_of[members[0]][2]
                    wcat = torch.empty(cin_pad * ktot * 4, dtype=torch.uint8, device=dev)
                    pack_weights([(tensors[lays[s].w], la0.cin, lays[s].cout, cin_pad, lays[s].cout, pick_nt(cin_pad), True,
                                   wcat, cat_of[s][1], ktot) for s in members])
                    shared_dx[key] = torch.empty(b, cin_pad, n, **f32)
                    probs.append({"a0": cat_img[key], "layer": _gemm_layer(wcat, cin_pad, ktot, pick_nt(cin_pad)),
                                  "out_cm": shared_dx[key], "rows_per_inst": n})
                _run_gemm_groups(probs, rows)
                if l > 0:
                    dys = [dxs[s][:, :lays[s].cin] for s in range(ns)]
                else:
                    done = set()
                    for s, st in enumerate(plan.stacks):
                        key = tuple(ti for ti, _, _ in st.inputs)
                        if key in shared_dx:
                            if key in done:
                                continue
                            done.add(key)
                            dx = shared_dx[key]
                        else:
                            dx = dxs[s]
                        c0 = 0
                        for ti, fmt, c in st.inputs:
                            if ctx.needs_input_grad[1 + ti]:
                                g = dx[:, c0:c0 + c]
                                acc(ti, g if fmt == "cm" else g.permute(0, 2, 1).reshape(rows, c))
                            c0 += c
        return (None,) + tuple(out_grads)


def mlp_stacks(stacks, b, n):
    """stacks: list of StackSpec.  Returns one fp32 (b, cout, n) tensor per stack (the last layer's output)."""
    tensors, index = [], {}

    def reg(t):
        if t is None:
            return None
        k = id(t)
        if k not in index:
            index[k] = len(tensors)
            tensors.append(t)
        return index[k]

    plan = SimpleNamespace(b=b, n=n, stacks=[], nlayers=len(stacks[0].layers))
    for st in stacks:
        assert len(st.layers) == plan.nlayers and 1 <= len(st.inputs) <= 2
        ins = []
        for t, fmt in st.inputs:
            c = t.shape[1]
            assert c % 32 == 0 and t.dtype == torch.float32 and t.is_cuda
            if fmt == "cm" and t.stride(2) != 1 or fmt == "rm" and t.stride(1) != 1:
                t = t.contiguous()
            ins.append((reg(t), fmt, c))
        cin = sum(c for _, _, c in ins)
        lays = []
        for kind, w, bias, bn, *rest in st.layers:
            cout = w.shape[0]
            w2 = w.reshape(cout, -1)
            assert w2.shape[1] == cin and cout % 64 == 0 and cin % 32 == 0, (cout, cin, w.shape)
            assert (bn is not None) == (kind in ("bn_relu", "relu_bn"))
            if bn is not None:
                assert bn.affine and bn.momentum is not None
            lays.append(SimpleNamespace(kind=kind, cout=cout, cin=cin, w=reg(w2.contiguous() if not w2.is_contiguous() else w2),
                                        bias=reg(bias), bn=bn, relu_mod=rest[0] if rest else None,
                                        gamma=reg(bn.weight) if bn is not None else None,
                                        beta=reg(bn.bias) if bn is not None else None))
            cin = cout
        plan.stacks.append(SimpleNamespace(inputs=ins, layers=lays))
    assert n % 128 == 0 and (b * n) % 128 == 0
    return _MlpStacksFn.apply(plan, *tensors)


# ------------------------------------------------------------------ module tree -> stack specs
def disengage_layers(stack):
    """nn.Sequential of BasicBlock_3DCONV (Conv3d 1x1x1 without bias -> BatchNorm3d -> ReLU)."""
    out = []
    for block in stack:
        mods = list(block.layers)
        conv, bn = mods[0], mods[1]
        assert isinstance(conv, nn.Conv3d) and conv.bias is None and isinstance(bn, nn.BatchNorm3d) and \
            isinstance(mods[2], nn.ReLU) and len(mods) == 3
        out.append(("bn_relu", conv.weight, None, bn, mods[2]))
    return out


def head_layers(head, pad_last_to=64):
    """Head_MultiLayerPerceptron: Conv1d(k=1) -> [ReLU] -> [BatchNorm1d] per layer.  A last layer narrower than 64
    outputs is zero-padded to `pad_last_to` rows (differentiable torch ops); returns (layers, true output width)."""
    mods = list(head.layers)
    out, i, width = [], 0, None
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, nn.Conv1d)
        relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
        j = i + 1 + int(relu)
        bn = mods[j] if j < len(mods) and isinstance(mods[j], nn.BatchNorm1d) else None
        j += int(bn is not None)
        w, bias = conv.weight.reshape(conv.out_channels, conv.in_channels), conv.bias
        width = conv.out_channels
        if width % 64 != 0:
            assert j == len(mods) and bn is None
            padn = _pad(width, pad_last_to) - width
            w = torch.nn.functional.pad(w, (0, 0, 0, padn))
            bias = torch.nn.functional.pad(bias, (0, padn)) if bias is not None else None
        kind = ("relu_bn" if bn is not None else "relu") if relu else "linear"
        assert not (bn is not None and not relu)
        out.append((kind, w, bias, bn, mods[i + 1] if relu else None))
        i = j
    return out, width
