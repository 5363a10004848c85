"""GPU diagnostic: host->device throughput of one bench batch (PoseEngine.load) on its own, and of the same bytes
as a single pinned buffer — tells whether the end-to-end number is PCIe-bound."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

dev = torch.device("cuda:0")
b = 32
batch = bench.make_host_batch(1234, b, pin=True)
nbytes = sum(f.numel() * 4 + i.numel() * 4 for side in ("inp", "tmp") for f, i in batch[side]) + sum(
    batch["points_" + s].numel() * 4 for s in ("inp", "tmp"))
flat = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for name, fn in (("single pinned buffer", lambda: dst.copy_(flat, non_blocking=True)),):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {nbytes/1e6:.1f} MB in {ms:.3f} ms = {nbytes/ms/1e6:.1f} GB/s")
# the engine's own load (18 copies + 8 fills)
devs = {side: [(f.to(dev), i.to(dev)) for f, i in batch[side]] for side in ("inp", "tmp")}
pts = {s: batch["points_" + s].to(dev) for s in ("inp", "tmp")}
def load():
    for side in ("inp", "tmp"):
        pts[side].copy_(batch["points_" + side], non_blocking=True)
        for (fd, idd), (f, i) in zip(devs[side], batch[side]):
            fd.copy_(f, non_blocking=True); idd.copy_(i, non_blocking=True)
load(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(20):
    load()
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"18 separate copies: {ms:.3f} ms = {nbytes/ms/1e6:.1f} GB/s; host issue time {1e3*(t1-t0)/20:.3f} ms per batch")
