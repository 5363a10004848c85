"""Driver for ncu: every neighbourhood op at the microbench shape (BASELINE.json configs[1]), twice (the second pass is
the warm one): FPS, ball_query, knn, three_nn, grouping fwd / bwd, three_interpolate fwd / bwd, gather fwd / bwd,
pointnet_sp three_nn + three_interpolate fwd / bwd.  Summarise the log with tools/summarize_ops_ncu.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import _lib as L                                   # noqa: E402
from dcl_net_b200.pointnet_lib import pointnet2_utils as pu          # noqa: E402
from dcl_net_b200.pointnet_sp import pointnet2_utils as pu_sp        # noqa: E402

dev = torch.device("cuda:0")
B, N, NP, NS, C = 32, 16384, 1024, 32, 128
g = torch.Generator().manual_seed(0)
xyz = torch.rand(B, N, 3, generator=g).to(dev)
idx = pu.furthest_point_sample(xyz, NP)
new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
bq = pu.ball_query(0.05, NS, xyz, new_xyz)
feats = torch.randn(B, C, N, generator=g).to(dev).requires_grad_(True)
known_f = torch.randn(B, C, NP, generator=g).to(dev).requires_grad_(True)
d, i3 = pu.three_nn(xyz, new_xyz)
w = 1.0 / (d + 1e-8)
w = (w / w.sum(2, keepdim=True)).contiguous()
flat_u = torch.cat([torch.arange(B).repeat_interleave(1024).float().unsqueeze(1), torch.rand(B * 1024, 3, generator=g)], 1).to(dev)
flat_k = torch.cat([torch.arange(B).repeat_interleave(300).float().unsqueeze(1), torch.rand(B * 300, 3, generator=g)], 1).to(dev)
kf = torch.randn(B * 300, 128, generator=g).to(dev).requires_grad_(True)
for it in range(2):
    for t in (feats, known_f, kf):
        t.grad = None
    pu.furthest_point_sample(xyz, NP)
    pu.ball_query(0.05, NS, xyz, new_xyz)
    pu.knn(16, xyz, new_xyz)
    pu.three_nn(xyz, new_xyz)
    out = pu.grouping_operation(feats, bq)
    out.backward(torch.ones_like(out))
    o2 = pu.three_interpolate(known_f, i3, w)
    o2.backward(torch.ones_like(o2))
    o3 = pu.gather_operation(feats, idx)
    o3.backward(torch.ones_like(o3))
    dd, ii = pu_sp.three_nn(flat_u, flat_k)
    ww = 1.0 / (dd + 1e-8)
    ww = (ww / ww.sum(1, keepdim=True)).contiguous()
    o4 = pu_sp.three_interpolate(kf, ii, ww)
    o4.backward(torch.ones_like(o4))
torch.cuda.synchronize()
