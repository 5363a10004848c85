"""GPU microbench: the stage-2 refinement loop (tools/test_YCBV_stage2.py:204-225, 2 iterations, B=32, N=1024) —
tensor-core refiner path against the layer-module (cuDNN/cuBLAS) path of the same drop-in Refiner."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200.refiner import Refiner, refine_poses

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B, N, IT = 32, 1024, 2
torch.manual_seed(0)
ref = Refiner().eval().to(dev)
g = torch.Generator().manual_seed(1)
pts = ((torch.rand(B, N, 3, generator=g) - 0.5) * 0.2).to(dev)
q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
rot = (q * torch.det(q).sign().view(B, 1, 1)).contiguous().to(dev)
trans = ((torch.rand(B, 3, generator=g) - 0.5) * 0.1).to(dev)
f = torch.randn(B, 256, N, generator=g).to(dev)
conf = torch.rand(B, 2 * N, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(iters=10):
    with torch.no_grad():
        for _ in range(3):
            out = refine_poses(ref, pts, rot, trans, f, conf, IT)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = refine_poses(ref, pts, rot, trans, f, conf, IT); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    return tot / iters, out


ms_fused, out_f = timed()
ref.use_fused = False
ms_layers, out_l = timed()
from oracle import torch_oracle as T  # checker only
ang = T.rotation_angle_deg(out_f[0].cpu(), out_l[0].cpu()).max().item()
res = {"what": f"stage-2 refinement, {IT} iterations, B={B}, N={N} (eager launches)", "ms_tensor_core_path": ms_fused,
       "ms_layer_module_path_fp32": ms_layers, "speedup": ms_layers / ms_fused, "max_angle_between_paths_deg": ang,
       "max_trans_diff_m": (out_f[1] - out_l[1]).abs().max().item(),
       "algorithmic_gflop": IT * B * N * 2.0 * (259 * 512 + 512 * 512 + 512 * 1024) / 1e9}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_stage2.json"), "w"), indent=1)
