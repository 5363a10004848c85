"""GPU microbench: ADD-S / CD_Dis distances (dcl_nearest_dist) against the reference's B x N x M x 3 broadcast
(tools/test_YCBV_stage1.py:188, models/DCL_Net.py:307-311) on the same inputs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import losses

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters, out


res = []
for b, n in ((32, 1024), (32, 2620), (8, 8192)):
    g = torch.Generator().manual_seed(n)
    a, c = torch.rand(b, n, 3, generator=g).to(dev), torch.rand(b, n, 3, generator=g).to(dev)
    ms, mine = timed(lambda: losses.adds_metric(a, c))
    ref_ms, ref = timed(lambda: torch.mean(torch.min(torch.norm(a.unsqueeze(2) - c.unsqueeze(1), dim=3), 2)[0], dim=1), 3)
    err = ((mine - ref).abs().max() / ref.abs().max()).item()
    evals = float(b) * n * n
    res.append({"op": "ADD-S (nearest_dist + mean)", "B": b, "N": n, "M": n, "ms": ms, "torch_broadcast_ms": ref_ms,
                "speedup": ref_ms / ms, "evals_per_s": evals / (ms * 1e-3), "rel_err_vs_broadcast": err,
                "broadcast_bytes": evals * 3 * 4})
    print(json.dumps(res[-1]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_adds.json"), "w"), indent=1)
