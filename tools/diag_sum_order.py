"""GPU diagnostic: which association does torch.sum(dim=1) use for 3 elements, and how does the rest of the
weight chain of models/Modules.py:222-224 round?  (informs the fused nn_interpolate kernel)"""
import torch
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
d2 = torch.rand(200000, 3, generator=g).to(dev) * 1e-3
dist = torch.sqrt(d2)
r = 1.0 / (dist + 1e-8)
s = torch.sum(r, dim=1, keepdim=True)
r0, r1, r2 = r[:, 0:1], r[:, 1:2], r[:, 2:3]
for name, v in (("(r0+r1)+r2", (r0 + r1) + r2), ("r0+(r1+r2)", r0 + (r1 + r2)), ("(r0+r2)+r1", (r0 + r2) + r1)):
    print(name, "mismatches:", int((v != s).sum()))
r_alt = torch.reciprocal(dist + 1e-8)
print("1.0/x vs reciprocal mismatches:", int((r != r_alt).sum()))
one = torch.ones_like(dist)
print("1.0/x vs ones/x mismatches:", int((r != one / (dist + 1e-8)).sum()))
w = r / s
print("w vs r*(1/s) mismatches:", int((w != r * (1.0 / s)).sum()))
x64 = (dist.double() + 1e-8)
print("dist+1e-8 in fp32 vs fp64-rounded mismatches:", int(((dist + 1e-8) != x64.float()).sum()))
