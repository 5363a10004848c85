"""GPU: per-iteration timeline of CTA (0,0,0) of each sparse-convolution launch of one towers pass (clock64 stamps,
dcl_debug_spconv_set_trace).  Prints, per layer, the mean cycles between consecutive events of every role."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dcl_net_b200 import _lib as L  # noqa: E402
from dcl_net_b200 import backbone as BB  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
b = 32
batch = bench.make_host_batch(1017, b, pin=False, entry="points")
caps = BB.SparseTowers.plan_capacities([batch["points_inp"], batch["points_tmp"]], b, 1024, dev)
torch.manual_seed(0)
tw = BB.SparseTowers(BB.Backbone_SPCONV().eval().to(dev), BB.Backbone_SPCONV().eval().to(dev), dev, b, 1024, caps)
args = [batch[k].to(dev) for k in ("points_inp", "rgb_inp", "points_tmp", "rgb_tmp")]
for _ in range(2):
    tw.run(*args)
torch.cuda.synchronize()
# trace every conv launch separately: patch the library call
orig = lib.dcl_spb_conv3
names = []
bufs = []


def traced(bb, cin, cout, nt, convs, st):
    buf = torch.zeros(6 * 512, dtype=torch.int64, device=dev)
    lib.dcl_debug_spconv_set_trace(L.ptr(buf))
    err = orig(bb, cin, cout, nt, convs, st)
    torch.cuda.synchronize()
    names.append(f"{cin}->{cout}")
    bufs.append(buf.cpu().numpy().reshape(6, 512))
    return err


lib.dcl_spb_conv3 = traced
tw.run(*args)
lib.dcl_spb_conv3 = orig
lib.dcl_debug_spconv_set_trace(None)
roles = ["gather: A buffer free", "gather: copies issued", "mma: A landed", "mma: issued", "W issued", "mma: W landed"]
for name, t in zip(names, bufs):
    n_it = int((t[1] > 0).sum())
    n_w = int((t[4] > 0).sum())
    t0 = t[t > 0].min()
    print(f"\n== conv {name}: {n_it} iterations, {n_w} W stages, CTA time {int(t.max() - t0)} cycles")
    for r, label in enumerate(roles):
        n = n_w if r >= 4 else n_it
        if n >= 2:
            d = np.diff(t[r, :n])
            print(f"  {label:24s} first at {int(t[r,0]-t0):7d}  mean step {d.mean():8.0f}  median {np.median(d):8.0f}  max {d.max():8.0f}")
    k = min(n_it, 12)
    print("  it:   free  issued  landed  mma_issued   (cycles since CTA start)")
    for i in range(k):
        print(f"  {i:3d} {int(t[0,i]-t0):7d} {int(t[1,i]-t0):7d} {int(t[2,i]-t0):7d} {int(t[3,i]-t0):7d}")
