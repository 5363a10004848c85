"""GPU diagnostic: where does the gap between the device-resident step and the host-in/host-out step come from?
Times PipelinedPoseEngine.infer_many with (a) everything, (b) the H2D copies stubbed out, (c) back-to-back graph
replays only."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dcl_net_b200.dcl_net import Network
from dcl_net_b200.engine import PipelinedPoseEngine, PoseEngine

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
b, steps = 32, 60
batches = [bench.make_host_batch(100 + i, b, pin=True) for i in range(3)]
caps = [max(max(bt[s][lv][0].shape[0] for bt in batches for s in ("inp", "tmp")), 1) for lv in range(4)]
torch.manual_seed(0)
net = Network(bench.Cfg(), mode="test", c_m=128).eval().to(dev)
pipe = PipelinedPoseEngine(net, dev, b, caps, depth=2, use_graph=True)

def run(tag):
    for _ in pipe.infer_many(batches[i % 3] for i in range(6)):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rot, trans in pipe.infer_many(batches[i % 3] for i in range(steps)):
        _ = float(trans[0, 0])
    torch.cuda.synchronize()
    print(f"{tag}: {1e3 * (time.perf_counter() - t0) / steps:.4f} ms/step", flush=True)

run("full e2e (H2D + pass + D2H)")
orig_load = PoseEngine.load
def fake_load(self, host_batch):
    self.h2d_bytes = 0
PoseEngine.load = fake_load
run("no H2D (buffers keep the last batch)")
PoseEngine.load = orig_load
eng = pipe.engines[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.run()
e1.record(); torch.cuda.synchronize()
print(f"graph replays back to back: {e0.elapsed_time(e1) / steps:.4f} ms/step")
