"""GPU diagnostic: clock64 timeline of one CTA of the FDA kernel (roles: MMA issuer, softmax warp, TMA producer).
Usage: python tools/trace_fda.py [C]   (DCL_FDA_SINGLE=1 selects the single-CTA kernel)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import _lib as L
from dcl_net_b200.modules import fda_align
dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, N, M = 32, 1024, 1024
g = torch.Generator().manual_seed(1)
ri1 = (0.2 * torch.randn(B, C, N, generator=g).relu()).to(dev)
ri2 = (0.2 * torch.randn(B, C, M, generator=g).relu()).to(dev)
re2 = torch.randn(B, 256, M, generator=g).to(dev)
for _ in range(3):
    fda_align(ri1, ri2, re2)
torch.cuda.synchronize()
buf = torch.zeros(4 * 1024, dtype=torch.int64, device=dev)
L.check(L.load().dcl_debug_fda_set_trace(L.ptr(buf)), "set trace")
fda_align(ri1, ri2, re2)
torch.cuda.synchronize()
L.check(L.load().dcl_debug_fda_set_trace(None), "clear trace")
life = buf.cpu()[3 * 1024:3 * 1024 + 16].view(2, 8)
t = buf.cpu()[:3 * 1024].view(3, 128, 8)
nb = M // 64
t0 = int(t[t > 0].min())
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
print(f"kernel: {'single' if os.environ.get('DCL_FDA_SINGLE') else 'pair'}  C={C}  (clock cycles relative to first stamp)")
print("MMA issuer: j | S: enter, k_full ok, s_empty ok | PV: enter, p_full ok, v_full[0] ok, v_full[3] ok, issued")
for j in range(nb):
    r = [rel(t[0, j, e]) for e in range(8)]
    print(f"  {j:2d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[3]:7d} {r[4]:7d} {r[6]:7d} {r[7]:7d} {r[5]:7d}")
print("softmax warp 2: j | enter, s_full ok, S read, exp done, o_done ok, P published")
for j in range(nb):
    r = [rel(t[1, j, e]) for e in range(6)]
    print(f"  {j:2d} | " + " ".join(f"{x:7d}" for x in r))
print("producer: j | V(j,0) enter, v_empty ok, V(j,3) v_empty ok")
for j in range(nb):
    r = [rel(t[2, j, e]) for e in range(3)]
    print(f"  {j:2d} | " + " ".join(f"{x:7d}" for x in r))
end = max(int(t[0].max()), int(t[1].max()))
print("span", end - t0, "cycles;  per key block", (end - t0) / nb)
base = int(life[0, 0])
for slot, name in enumerate(("CTA (0,0)", "last leader CTA")):
    print(f"life-cycle of {name} [ns after CTA (0,0) entered]: entry, set-up done, main loop done, epilogue done, exit = "
          + ", ".join(str(int(life[slot, e]) - base) for e in range(5)))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    fda_align(ri1, ri2, re2)
ev1.record()
torch.cuda.synchronize()
print("pack + fwd, avg of 10 (us):", ev0.elapsed_time(ev1) * 100)
