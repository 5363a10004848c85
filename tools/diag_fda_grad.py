"""GPU diagnostic: accuracy of FdaAlignFunction's forward/backward against fp64, next to the fp32 torch graph."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200.modules import fda_align
from oracle import torch_oracle as T
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
for scale in (1.0, 2.0):
    g = torch.Generator().manual_seed(31)
    ri1 = (scale * torch.randn(2, 64, 256, generator=g).relu()).to(dev)
    ri2 = (scale * torch.randn(2, 64, 256, generator=g).relu()).to(dev)
    re2 = torch.randn(2, 256, 256, generator=g).to(dev)
    ge, gi = torch.randn(2, 256, 256, generator=g).to(dev), torch.randn(2, 64, 256, generator=g).to(dev)
    def run(fn, dt):
        xs = [t.to(dt).clone().requires_grad_(True) for t in (ri1, ri2, re2)]
        out = fn(*xs)
        e, m = out[0], out[1]
        ((e * ge.to(dt)).sum() + (m * gi.to(dt)).sum()).backward()
        return [e.detach(), m.detach()] + [x.grad for x in xs]
    mine = run(lambda a, b, c: fda_align(a, b, c), torch.float32)
    r32 = run(lambda a, b, c: T.fda_direction(a, b, c), torch.float32)
    r64 = run(lambda a, b, c: T.fda_direction(a, b, c), torch.float64)
    names = ["RE_embed", "RI_embed", "dRI_1", "dRI_2", "dRE_2"]
    for n, a, b, c in zip(names, mine, r32, r64):
        s = c.abs().max().item()
        print(f"scale {scale} {n:9s} mine {(a.double()-c).abs().max().item()/s:.2e}   fp32 graph {(b.double()-c).abs().max().item()/s:.2e}")
