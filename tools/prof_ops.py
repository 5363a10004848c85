"""Small driver for ncu: FPS, ball_query, knn at the microbench shape (one launch each after a warm-up)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200.pointnet_lib import pointnet2_utils as pu  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
xyz = torch.rand(32, 16384, 3, generator=g).to(dev)
for _ in range(2):
    idx = pu.furthest_point_sample(xyz, 1024)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    bq = pu.ball_query(0.05, 32, xyz, new_xyz)
    d, i = pu.knn(16, xyz, new_xyz)
torch.cuda.synchronize()
