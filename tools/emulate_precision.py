"""CPU experiment (no GPU): how much operand precision do the tensor-core contractions of the stage-1 tail need?

The product runs every contraction as 3 bf16 MMAs on hi/lo-split operands (~2^-17 per operand).  This script
emulates cheaper operand formats inside the fp64 restated graph (oracle/torch_oracle.py) by rounding the A operand of
every pointwise layer (= the stored activation image) and the FDA operands, and reports the deviation of the
outputs from the unrounded fp64 graph next to the tolerances of BASELINE.json (features 1e-3 relative, poses
0.01 deg / 1e-5 m).  Weight-side rounding is not emulated: a B operand can always be split hi/lo at no extra MMA
beyond the count given.  Usage:  python tools/emulate_precision.py [--b 4] [--c_m 128] [--seeds 3]
"""
import argparse
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as T  # noqa: E402  (test infrastructure; this tool is an experiment, not product)


def r_bf16x2(x):
    hi = x.to(torch.bfloat16).to(x.dtype)
    lo = (x - hi).to(torch.bfloat16).to(x.dtype)
    return hi + lo


def r_fp16(x):
    return x.to(torch.float16).to(x.dtype)


def r_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def r_fp16x2(x):
    hi = x.to(torch.float16).to(x.dtype)
    return hi + (x - hi).to(torch.float16).to(x.dtype)


ROUND = {"exact": lambda x: x, "bf16x2": r_bf16x2, "fp16": r_fp16, "bf16": r_bf16, "fp16x2": r_fp16x2}


def run_variant(net64, f_xc, f_yo, b, n, act_fmt, qk_fmt, v_fmt, p_fmt):
    """act_fmt: callable(module_name) -> format of that layer's A operand; FDA: q/k, v, p formats."""
    net = copy.deepcopy(net64)
    handles = []
    for name, mod in net.named_modules():
        if isinstance(mod, (torch.nn.Conv1d, torch.nn.Conv3d)) and not name.startswith(("regressor_rot", "regressor_trans")):
            fmt = ROUND[act_fmt(name)]
            handles.append(mod.register_forward_pre_hook(lambda m, inp, fmt=fmt: (fmt(inp[0]),)))
    exact_aligner = T.aligner

    def aligner(ri_1, ri_2, re_2):
        q, k = ROUND[qk_fmt](ri_1), ROUND[qk_fmt](ri_2)
        a = torch.softmax(torch.bmm(k.transpose(1, 2), q), dim=1)
        # the kernel normalises after the P V product: P = exp(s - max) in (0,1] is what gets rounded
        s = torch.bmm(k.transpose(1, 2), q)
        p = torch.exp(s - s.max(dim=1, keepdim=True).values)
        pr = ROUND[p_fmt](p)
        out = torch.bmm(ROUND[v_fmt](re_2), pr) / p.sum(dim=1, keepdim=True)
        # the second product (RI_2 A) of the same pass uses the same rounded P; returned through `a`
        a_eff = pr / p.sum(dim=1, keepdim=True)
        return out, a_eff

    T.aligner = aligner
    try:
        with torch.no_grad():
            out = net(f_xc, f_yo, b, n, n)
    finally:
        T.aligner = exact_aligner
        for h in handles:
            h.remove()
    return out


def deviation(out, ref):
    rel = lambda a, c: ((a - c).abs().max() / c.abs().max()).item()
    d = ref["_debug"]
    return {"F_Xo_p": rel(out["F_Xo_p"], ref["F_Xo_p"]), "F_Yc_p": rel(out["_debug"]["F_Yc_p"], d["F_Yc_p"]),
            "F_Xo_m": rel(out["_debug"]["F_Xo_m"], d["F_Xo_m"]), "conf": rel(out["conf"], ref["conf"]),
            "rot_deg": T.rotation_angle_deg(out["rot_pred"], ref["rot_pred"]).max().item(),
            "trans_m": (out["trans_pred"] - ref["trans_pred"]).abs().max().item()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=4)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--c_m", type=int, default=128)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    m_branch = lambda name: "_m1" in name or "_m2" in name
    variants = {
        # name: (A-operand format per layer, q/k, v, p, MMAs per product: layers / QK / PV)
        "3mma_bf16x2_everywhere (product, round 1)": (lambda n: "bf16x2", "bf16x2", "bf16x2", "bf16x2"),
        "P single fp16, rest bf16x2": (lambda n: "bf16x2", "bf16x2", "bf16x2", "fp16"),
        "P single bf16, rest bf16x2": (lambda n: "bf16x2", "bf16x2", "bf16x2", "bf16"),
        "activations fp16 after the FDA only (fusers, conf heads)": (
            lambda n: "fp16" if n.startswith(("neck_fuser", "regressor_conf", "regressor_Xo", "regressor_Yc")) else "bf16x2",
            "bf16x2", "bf16x2", "bf16x2"),
        "activations fp16 except the q/k (m) branches; P fp16": (
            lambda n: "bf16x2" if m_branch(n) else "fp16", "bf16x2", "bf16x2", "fp16"),
        "activations fp16 everywhere; q/k bf16x2; P fp16": (lambda n: "fp16", "bf16x2", "bf16x2", "fp16"),
        "activations fp16 everywhere; q/k fp16; P fp16": (lambda n: "fp16", "fp16", "fp16", "fp16"),
        "activations bf16 single everywhere (plain bf16 inference)": (lambda n: "bf16", "bf16", "bf16", "bf16"),
    }
    results = {k: [] for k in variants}
    results["fp32 reference graph itself (torch fp32 vs fp64)"] = []
    for seed in range(args.seeds):
        torch.manual_seed(100 + seed)
        net32 = T.TailNetwork(mode="test", c_m=args.c_m).eval()
        net64 = copy.deepcopy(net32).double()
        g = torch.Generator().manual_seed(seed)
        f_xc = torch.randn(args.b * args.n, 480, generator=g)
        f_yo = torch.randn(args.b * args.n, 480, generator=g)
        with torch.no_grad():
            ref = net64(f_xc.double(), f_yo.double(), args.b, args.n, args.n)
            out32 = net32(f_xc, f_yo, args.b, args.n, args.n)
        out32 = {k: (v.double() if torch.is_tensor(v) else {kk: vv.double() for kk, vv in v.items()}) for k, v in out32.items()}
        results["fp32 reference graph itself (torch fp32 vs fp64)"].append(deviation(out32, ref))
        for name, (act, qk, v, p) in variants.items():
            out = run_variant(net64, f_xc.double(), f_yo.double(), args.b, args.n, act, qk, v, p)
            results[name].append(deviation(out, ref))
            print(name, results[name][-1], flush=True)
    summary = {name: {k: max(r[k] for r in runs) for k in runs[0]} for name, runs in results.items()}
    print(json.dumps({"b": args.b, "n": args.n, "c_m": args.c_m, "seeds": args.seeds, "worst_over_seeds": summary}, indent=1))
    if args.out:
        json.dump({"b": args.b, "n": args.n, "c_m": args.c_m, "seeds": args.seeds, "worst_over_seeds": summary,
                   "tolerances": {"features_rel": 1e-3, "rot_deg": 0.01, "trans_m": 1e-5}}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
