"""Runs on the GPU box: executes the REFERENCE's own kernels (oracle/_ref) on seeded inputs and stores their
outputs as fixtures (gpurun_out/golden/neighbour_ref.npz -> committed as tests/golden/neighbour_ref.npz).
The CPU test suite then pins the C oracle against them (tests/test_cpu_oracle.py).  Inputs are regenerated
from the seeds by tests/util.py, so only outputs are stored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_kernels as R  # noqa: E402
from dcl_testutil import cad_like_cloud, flat_bxyz, uniform_cloud  # noqa: E402


def cases():
    """name -> (callable taking a device, returning dict of numpy outputs)."""
    g = lambda s: torch.Generator().manual_seed(s)
    out = {}

    def fps(dev):
        a = cad_like_cloud(1, 2, 1000).to(dev)
        b = uniform_cloud(2, 1, 4096).to(dev)
        i1, t1 = R.furthest_point_sample(a, 64, return_temp=True)
        i2 = R.furthest_point_sample(b, 128)
        return {"fps_cad_idx": i1, "fps_cad_temp": t1, "fps_uni_idx": i2}

    def bq(dev):
        xyz = uniform_cloud(3, 2, 2048).to(dev)
        cad = cad_like_cloud(4, 2, 1500).to(dev)
        return {"bq_uni": R.ball_query(0.1, 16, xyz, xyz[:, :64].contiguous()),
                "bq_cad": R.ball_query(0.02, 8, cad, cad[:, :50].contiguous())}

    def nn(dev):
        u, k = cad_like_cloud(5, 2, 512).to(dev), cad_like_cloud(6, 2, 200).to(dev)
        d3, i3 = R.three_nn(u, k)
        dk, ik = R.knn(8, u, k)
        feats = torch.randn(2, 6, 200, generator=g(7)).to(dev)
        w = torch.rand(2, 512, 3, generator=g(8)).to(dev)
        return {"nn3_d2": d3, "nn3_idx": i3, "knn_d2": dk, "knn_idx": ik,
                "interp": R.three_interpolate(feats, i3, w)}

    def grp(dev):
        feats = torch.randn(2, 5, 300, generator=g(9)).to(dev)
        idx = torch.randint(0, 300, (2, 20, 4), generator=g(10), dtype=torch.int32).to(dev)
        return {"group": R.grouping_operation(feats, idx), "gather": R.gather_operation(feats, idx[:, :, 0].contiguous())}

    def sp(dev):
        u, k = flat_bxyz(11, 3, 100, shuffle=False), flat_bxyz(12, 3, 40)
        u[:, 1:] = (u[:, 1:] * 64).round() / 64
        k[:, 1:] = (k[:, 1:] * 64).round() / 64
        k = k[k[:, 0] != 1].contiguous()  # batch 1 has no known rows
        u, k = u.to(dev), k.to(dev)
        d, i = R.sp_three_nn(u, k)
        feats = torch.randn(k.shape[0], 8, generator=g(13)).to(dev)
        w = torch.rand(300, 3, generator=g(14)).to(dev)
        return {"sp_d2": d, "sp_idx": i, "sp_interp": R.sp_three_interpolate(feats, i, w)}

    for f in (fps, bq, nn, grp, sp):
        out[f.__name__] = f
    return out


def main():
    dev = torch.device("cuda:0")
    res = {}
    for name, fn in cases().items():
        for k, v in fn(dev).items():
            res[k] = v.cpu().numpy()
    torch.cuda.synchronize()
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "neighbour_ref.npz"), **res)
    print("wrote", os.path.join(dst, "neighbour_ref.npz"), {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
