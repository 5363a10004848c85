"""Kernel-time breakdown of one training step (tools/train_step_ddp.py's step, 1 GPU) with torch.profiler:
prints the kernels by total device time and the step's wall time.  python tools/prof_train.py [--batch 40]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from train_step_ddp import Cfg, losses  # noqa: E402
from dcl_net_b200.dcl_net import Network  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=40)
    ap.add_argument("--rows", type=int, default=45)
    ap.add_argument("--layers", action="store_true", help="PyTorch layer modules instead of the training kernels")
    ap.add_argument("--entry", default="feats", choices=["feats", "backbone"])
    ap.add_argument("--cprofile", action="store_true", help="host-side profile (cProfile) of three steps")
    ap.add_argument("--ncu", action="store_true", help="two plain steps and exit (run under ncu; summarise the second half)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = Network(Cfg, mode="train").to(dev).train()
    net.use_train_kernels = not args.layers
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    b, n = args.batch, 1024
    g = torch.Generator(device=dev).manual_seed(1000)
    f_xc = torch.randn(b * n, 480, device=dev, generator=g)
    f_yo = torch.randn(b * n, 480, device=dev, generator=g)
    pts_tmp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
    pts_inp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
    q, _ = torch.linalg.qr(torch.randn(b, 3, 3, device=dev, generator=g))
    rot_gt = q * torch.det(q).sign().view(b, 1, 1)
    trans_gt = (torch.rand(b, 3, device=dev, generator=g) - 0.5) * 0.1

    if args.entry == "backbone":
        from dcl_net_b200.synthetic import backbone_levels, levels_to, object_clouds
        p_inp, p_tmp = object_clouds(2000, b, n, partial=True), object_clouds(3000, b, n)
        lv_inp = levels_to(backbone_levels(4000, p_inp, b), dev)
        lv_tmp = levels_to(backbone_levels(5000, p_tmp, b), dev)
        for lv in lv_inp + lv_tmp:
            lv.features.requires_grad_(True)
        p_inp, p_tmp = p_inp.to(dev), p_tmp.to(dev)
        pts_inp, pts_tmp = p_inp.view(b, n, 3), p_tmp.view(b, n, 3)

    def step():
        opt.zero_grad(set_to_none=True)
        if args.entry == "backbone":
            for lv in lv_inp + lv_tmp:
                lv.features.grad = None
            out = net.forward_from_backbone(lv_inp, lv_tmp, p_inp, p_tmp, b)
        else:
            out = net.forward_from_point_feats(f_xc, f_yo, b)
        loss = losses(out, pts_tmp, pts_inp, rot_gt, trans_gt)
        loss.backward()
        opt.step()

    if args.ncu:
        step()
        step()
        torch.cuda.synchronize()
        return
    if args.cprofile:
        import cProfile
        import pstats
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            step()
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)
        st.sort_stats("tottime").print_stats(25)
        return
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"ms per step (CUDA events, 5 steps): {e0.elapsed_time(e1) / 5:.3f}")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = {}
    for e in evs:
        t = tot.setdefault(e.name, [0.0, 0])
        t[0] += e.device_time
        t[1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"device time of one step: {total / 1e3:.3f} ms over {sum(v[1] for v in tot.values())} kernels / copies")
    for name, (us, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:args.rows]:
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  x{cnt:<4d} {name[:110]}")


if __name__ == "__main__":
    main()
