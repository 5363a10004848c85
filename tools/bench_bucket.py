"""GPU diagnostic: cost of the cluster bucket build as a function of the level size (tiny query set, so the search
launch is negligible); reports the two launches together via CUDA events."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200.pointnet_sp import pointnet2_utils as pu
from dcl_net_b200 import fused_tail as FT

dev = torch.device("cuda:0")
b, n_per = 32, 32
g = torch.Generator().manual_seed(0)
unknown = torch.cat([torch.arange(b).repeat_interleave(n_per).float().unsqueeze(1),
                     (torch.rand(b * n_per, 3, generator=g) - 0.5) * 0.3], 1).contiguous().to(dev)
for m in (256, 2048, 8192, 34000):
    ind = torch.cat([torch.randint(0, b, (m, 1), generator=g), torch.randint(0, 32, (m, 3), generator=g)], 1).int()
    ind = torch.unique(ind, dim=0)
    ind = ind[torch.randperm(ind.shape[0], generator=g)].contiguous().to(dev)
    feats = torch.randn(ind.shape[0], 32, generator=g).to(dev)
    out = torch.zeros(FT.pm_bytes(b * n_per, 32), dtype=torch.uint8, device=dev)
    spec = [(ind, [0.6 / 32] * 3, [-0.3] * 3, feats, 0, 32)]
    for nlev in (1, 4):
        specs = spec * nlev
        specs = [(s[0], s[1], s[2], s[3], 0, s[5]) for s in specs]
        for _ in range(3):
            pu.nn_interpolate_vox_levels_pm(unknown, specs[:1] if nlev == 1 else specs, out, 32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pu.nn_interpolate_vox_levels_pm(unknown, specs, out, 32)
        e1.record(); torch.cuda.synchronize()
        print(f"m={ind.shape[0]:6d} levels={nlev}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (build + 1024-query search)")
