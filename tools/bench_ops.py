"""Op microbench (BASELINE.json configs[1]): B=32 synthetic clouds — FPS 16384->1024, ball_query r=0.05 k=32,
grouping C=128 fwd/bwd, three_nn 16384x1024 + three_interpolate C=128 fwd/bwd, knn k=16, gather, and the flat
pointnet_sp ops at the stage-1 shape.  Each op: CUDA events, L2 flushed between iterations, median of `iters`;
the reference's own kernel (oracle/_ref, compiled unmodified) is timed the same way next to it; the bit-exact index
check is repeated on the timed inputs.  Prints one JSON line per op and writes gpurun_out/bench_ops.json."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import synthetic                                   # noqa: E402
from dcl_net_b200.pointnet_lib import pointnet2_utils as pu          # noqa: E402
from dcl_net_b200.pointnet_sp import pointnet2_utils as pu_sp        # noqa: E402
from oracle import ref_kernels as R                                  # noqa: E402

dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


results = []


def report(name, ms, ref_ms, alg_bytes, evals=None, exact=None, note=""):
    rec = {"op": name, "ms": ms, "ref_kernel_ms": ref_ms, "speedup_vs_ref_kernel": (ref_ms / ms) if ref_ms else None,
           "algorithmic_GB": alg_bytes / 1e9, "achieved_GBps": alg_bytes / ms / 1e6, "hbm_frac": alg_bytes / ms / 1e6 / HBM,
           "evals_per_s": (evals / (ms * 1e-3)) if evals else None, "bit_exact_vs_ref": exact, "note": note}
    results.append(rec)
    print(json.dumps(rec), flush=True)


def main():
    B, N, NP, NS, C, K = 32, 16384, 1024, 32, 128, 16
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(B, N, 3, generator=g).to(dev)
    have_ref = R.available()

    idx = pu.furthest_point_sample(xyz, NP)
    ref_idx = R.furthest_point_sample(xyz, NP) if have_ref else None
    report("fps 16384->1024", timed(lambda: pu.furthest_point_sample(xyz, NP), 5, 2),
           timed(lambda: R.furthest_point_sample(xyz, NP), 5, 2) if have_ref else None,
           B * (12 * N + 4 * NP), evals=B * N * (NP - 1), exact=bool(torch.equal(idx, ref_idx)) if have_ref else None,
           note="serial over 1023 rounds; evals = distance evaluations")

    xyz_t = xyz.transpose(1, 2).contiguous()
    new_xyz = pu.gather_operation(xyz_t, idx).transpose(1, 2).contiguous()
    report("gather xyz (C=3)", timed(lambda: pu.gather_operation(xyz_t, idx)),
           timed(lambda: R.gather_operation(xyz_t, idx)) if have_ref else None, B * (4 * NP + 4 * 3 * NP * 2))

    bq = pu.ball_query(0.05, NS, xyz, new_xyz)
    ref_bq = R.ball_query(0.05, NS, xyz, new_xyz) if have_ref else None
    report("ball_query r=0.05 ns=32", timed(lambda: pu.ball_query(0.05, NS, xyz, new_xyz)),
           timed(lambda: R.ball_query(0.05, NS, xyz, new_xyz)) if have_ref else None,
           B * (12 * N + 12 * NP + 4 * NP * NS), evals=B * NP * N,
           exact=bool(torch.equal(bq, ref_bq)) if have_ref else None, note="evals upper bound (early exit)")

    feats = torch.randn(B, C, N, generator=g).to(dev)
    out = pu.grouping_operation(feats, bq)
    report("grouping fwd C=128", timed(lambda: pu.grouping_operation(feats, bq)),
           timed(lambda: R.grouping_operation(feats, bq)) if have_ref else None,
           B * (4 * C * N + 4 * NP * NS + 4 * C * NP * NS),
           exact=bool(torch.equal(out, R.grouping_operation(feats, bq))) if have_ref else None)
    go = torch.randn(B, C, NP, NS, generator=g).to(dev)
    from dcl_net_b200 import _lib as L
    lib = L.load()
    gbuf = torch.zeros(B, C, N, device=dev)

    def group_bwd():
        gbuf.zero_()
        L.check(lib.dcl_lib_group_points_grad_kernel_launcher_fast(B, C, N, NP, NS, L.ptr(go), L.ptr(bq), L.ptr(gbuf),
                                                                   L.stream_ptr()), "gg")
    report("grouping bwd C=128 (incl. zero-fill)", timed(group_bwd),
           timed(lambda: R.grouping_operation_grad(go, bq, N)) if have_ref else None,
           B * (4 * C * N + 4 * NP * NS + 4 * C * NP * NS))
    del out, go, gbuf

    dist, i3 = pu.three_nn(xyz, new_xyz)
    if have_ref:
        d2r, i3r = R.three_nn(xyz, new_xyz)
    report("three_nn 16384 x 1024", timed(lambda: pu.three_nn(xyz, new_xyz)),
           timed(lambda: R.three_nn(xyz, new_xyz)) if have_ref else None, B * (12 * N + 12 * NP + 24 * N),
           evals=B * N * NP, exact=bool(torch.equal(i3, i3r) and torch.equal(dist, torch.sqrt(d2r))) if have_ref else None)
    recip = 1.0 / (dist + 1e-8)
    w = (recip / recip.sum(2, keepdim=True)).contiguous()
    f2 = torch.randn(B, C, NP, generator=g).to(dev)
    o = pu.three_interpolate(f2, i3, w)
    report("three_interpolate fwd C=128", timed(lambda: pu.three_interpolate(f2, i3, w)),
           timed(lambda: R.three_interpolate(f2, i3, w)) if have_ref else None, B * (4 * C * NP + 24 * N + 4 * C * N),
           exact=bool(torch.equal(o, R.three_interpolate(f2, i3, w))) if have_ref else None)
    go2 = torch.randn(B, C, N, generator=g).to(dev)
    gb2 = torch.zeros(B, C, NP, device=dev)

    def interp_bwd():
        gb2.zero_()
        L.check(lib.dcl_lib_three_interpolate_grad_kernel_launcher_fast(B, C, N, NP, L.ptr(go2), L.ptr(i3), L.ptr(w),
                                                                        L.ptr(gb2), L.stream_ptr()), "ig")
    report("three_interpolate bwd C=128 (incl. zero-fill)", timed(interp_bwd),
           timed(lambda: R.three_interpolate_grad(go2, i3, w, NP)) if have_ref else None,
           B * (4 * C * NP + 24 * N + 4 * C * N))
    del o, go2, f2

    for (nn, mm, tag) in ((NP, NP, "1024 x 1024"), (N, NP, "16384 x 1024")):
        u = new_xyz if nn == NP else xyz
        dk, ik = pu.knn(K, u, new_xyz)
        if have_ref:
            dkr, ikr = R.knn(K, u, new_xyz)
        report(f"knn k=16 {tag}", timed(lambda: pu.knn(K, u, new_xyz)),
               timed(lambda: R.knn(K, u, new_xyz)) if have_ref else None, B * (12 * nn + 12 * mm + 8 * K * nn),
               evals=B * nn * mm, exact=bool(torch.equal(ik, ikr)) if have_ref else None)

    # pointnet_sp at the stage-1 shape: 32 x 1024 queries against the level-1..4 pyramids
    pts = synthetic.object_clouds(5, B, 1024)
    levels = synthetic.backbone_levels(6, pts, B)
    unknown = torch.cat([torch.arange(B).repeat_interleave(1024).float().unsqueeze(1), pts], 1).to(dev)
    from dcl_net_b200.modules import Ops_tensor2points
    import numpy as np
    for li, (lvl, scale) in enumerate(zip(levels, (2, 4, 6, 8))):
        lv = synthetic.levels_to([lvl], dev)[0]
        fe, known = Ops_tensor2points(lv, -0.5 * 0.006 * 64 * np.ones(3), 0.006 * scale * np.ones(3))
        known, fe = known.contiguous(), fe.contiguous()
        nt, mt, cl = unknown.shape[0], known.shape[0], fe.shape[1]
        d, i = pu_sp.three_nn(unknown, known)
        if have_ref:
            dr, ir = R.sp_three_nn(unknown, known)
        seg_evals = sum(int((known[:, 0] == b).sum()) for b in range(B)) * 1024
        report(f"sp.three_nn L{li + 1} Nt={nt} Mt={mt}", timed(lambda: pu_sp.three_nn(unknown, known)),
               timed(lambda: R.sp_three_nn(unknown, known)) if have_ref else None, 16 * (nt + mt) + 24 * nt,
               evals=seg_evals, exact=bool(torch.equal(i, ir) and torch.equal(d, torch.sqrt(dr))) if have_ref else None,
               note="segmented; evals = sum_b n_b*m_b")
        rr = 1.0 / (d + 1e-8)
        ww = (rr / rr.sum(1, keepdim=True)).contiguous()
        report(f"sp.three_interpolate L{li + 1} C={cl}", timed(lambda: pu_sp.three_interpolate(fe, i, ww)),
               timed(lambda: R.sp_three_interpolate(fe, i, ww)) if have_ref else None, 4 * cl * mt + 24 * nt + 4 * cl * nt)
        report(f"sp.nn_interpolate fused L{li + 1} C={cl}", timed(lambda: pu_sp.nn_interpolate(unknown, known, fe)), None,
               16 * (nt + mt) + 4 * cl * mt + 4 * cl * nt, evals=seg_evals, note="search + weights + interpolation in one call")

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
