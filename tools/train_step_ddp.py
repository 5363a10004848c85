"""Training-step benchmark / smoke (BASELINE.json configs[4]): forward + backward through the FDA section with the
reference's losses, B instances per GPU, DDP over NCCL (one process per GPU; gradients all-reduced over NVLink,
overlapped with backward by DDP's bucket hooks).  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_step_ddp.py --batch 40 --steps 10
    python tools/train_step_ddp.py --batch 40 --steps 10          # single GPU, no process group

Entry = point features (b*n, 480) per tower (the towers themselves are outside this path).  Prints one JSON line on
rank 0: step time (max over ranks), instances/s, and the gradient bytes all-reduced per step."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200.dcl_net import Network  # noqa: E402


class Cfg:
    n_inp = n_tmp = 1024
    unit_voxel_extent = [0.006] * 3


def losses(out, pts_tmp, pts_inp, rot_gt, trans_gt):
    """Non-symmetric branch of models/DCL_Net.py:265-303 (L2 pose loss, Xo / Yc correspondence losses, conf loss)."""
    posed = torch.bmm(pts_tmp, out["rot_pred"].transpose(1, 2)) + out["trans_pred"].unsqueeze(1)
    posed_gt = torch.bmm(pts_tmp, rot_gt.transpose(1, 2)) + trans_gt.unsqueeze(1)
    l_pose = torch.norm(posed - posed_gt, dim=2).mean(dim=1).mean()
    inp_cano_gt = torch.bmm(pts_inp - trans_gt.unsqueeze(1), rot_gt).detach()
    l_xo = torch.norm(out["Xo_pred"] - inp_cano_gt, dim=2)
    l_yc = torch.norm(out["Yc_pred"] - posed_gt, dim=2)
    conf = out["conf"]
    l_conf = torch.mean(torch.cat([l_xo, l_yc], dim=1).detach() * conf - 0.01 * torch.log(conf))
    return l_pose + 5 * l_xo.mean() + l_yc.mean() + l_conf


class FromPointFeats(torch.nn.Module):
    """DDP wraps a module whose forward takes the tensors, so that its reducer arms the gradient hooks."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, f_xc, f_yo, nb):
        return self.inner.forward_from_point_feats(f_xc, f_yo, nb)


class FromBackbone(torch.nn.Module):
    """Entry at the towers' pyramid levels: pointnet_sp three_nn + three_interpolate (with their backward kernels:
    the level features carry gradients, as they do when the towers train) -> FDA section."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, levels_inp, levels_tmp, points_inp, points_tmp, nb):
        return self.inner.forward_from_backbone(levels_inp, levels_tmp, points_inp, points_tmp, nb)


def run(batch, steps, warmup, rank, world, local, contract=False, layers=False, entry="backbone", graph=False):
    """One process per GPU.  Times `steps` training steps (max over ranks), then — multi-GPU — the same steps with
    the gradient all-reduce switched off (DDP.no_sync) and the all-reduce of a gradient-sized buffer on its own:
    exposed all-reduce time = synced step - unsynced step; overlap = 1 - exposed / standalone."""
    import contextlib
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = Network(Cfg, mode="train").to(dev).train()
    net.use_train_kernels = not layers      # layers=True: the nn layer modules on library GEMMs (A/B figure)
    wrapped = FromBackbone(net) if entry == "backbone" else FromPointFeats(net)
    # graph mode: forward + backward (+ Adam on one GPU) replayed as ONE CUDA graph; data parallelism is then a single
    # NCCL all-reduce of the flat gradient buffer after the replay instead of DDP's bucket hooks (the step issues ~750
    # small launches and is otherwise paced by the host; the 19.5 MB all-reduce takes 0.11 ms over NVLink)
    use_ddp = world > 1 and not graph
    model = (torch.nn.parallel.DistributedDataParallel(wrapped, device_ids=[local], gradient_as_bucket_view=True)
             if use_ddp else wrapped)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=graph and world == 1)
    b, n = batch, 1024
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    if entry == "backbone":
        from dcl_net_b200.synthetic import backbone_levels, levels_to, object_clouds
        p_inp, p_tmp = object_clouds(2000 + rank, b, n, partial=True), object_clouds(3000 + rank, b, n)
        lv_inp = levels_to(backbone_levels(4000 + rank, p_inp, b), dev)
        lv_tmp = levels_to(backbone_levels(5000 + rank, p_tmp, b), dev)
        for lv in lv_inp + lv_tmp:
            lv.features.requires_grad_(True)
        p_inp, p_tmp = p_inp.to(dev), p_tmp.to(dev)
        pts_inp, pts_tmp = p_inp.view(b, n, 3), p_tmp.view(b, n, 3)
        fwd = lambda: model(lv_inp, lv_tmp, p_inp, p_tmp, b)
    else:
        f_xc = torch.randn(b * n, 480, device=dev, generator=g)
        f_yo = torch.randn(b * n, 480, device=dev, generator=g)
        pts_tmp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
        pts_inp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
        fwd = lambda: model(f_xc, f_yo, b)
    q, _ = torch.linalg.qr(torch.randn(b, 3, 3, device=dev, generator=g))
    rot_gt = q * torch.det(q).sign().view(b, 1, 1)
    trans_gt = (torch.rand(b, 3, device=dev, generator=g) - 0.5) * 0.1

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        if entry == "backbone":
            for lv in lv_inp + lv_tmp:
                lv.features.grad = None
        ctx = contextlib.nullcontext() if (sync or not use_ddp) else model.no_sync()
        with ctx:
            loss = losses(fwd(), pts_tmp, pts_inp, rot_gt, trans_gt)
            loss.backward()
        opt.step()
        return loss

    if graph:
        from dcl_net_b200.sharding import average_gradients, flat_grad_buffer
        flat = flat_grad_buffer(net.parameters())   # gradients accumulate in place into one flat buffer (static addresses)

        def fwd_bwd():
            flat.zero_()
            if entry == "backbone":
                for lv in lv_inp + lv_tmp:
                    lv.features.grad = None
            loss = losses(fwd(), pts_tmp, pts_inp, rot_gt, trans_gt)
            loss.backward()
            return loss

        def reduce_and_step(sync=True):
            if world > 1 and sync:
                average_gradients(flat)
            opt.step()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fwd_bwd()
                reduce_and_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cuda_graph):
            static_loss = fwd_bwd()
            if world == 1:
                opt.step()

        def step(sync=True):  # noqa: F811
            cuda_graph.replay()
            if world > 1:
                reduce_and_step(sync)
            return static_loss

    def timed(k, sync=True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            loss = step(sync)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), loss

    from dcl_net_b200 import _lib
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    launches0 = _lib.load().dcl_b200_launch_count()
    ms, loss = timed(steps)
    launches = (_lib.load().dcl_b200_launch_count() - launches0) // steps
    nparam = sum(p.numel() for p in net.parameters())
    extra = {}
    if world > 1:
        ms_nosync, _ = timed(steps, sync=False)
        flat = torch.zeros(nparam, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1) / 10
        exposed = max(ms - ms_nosync, 0.0)
        extra = {"ms_per_step_without_allreduce": ms_nosync, "allreduce_ms_exposed": exposed,
                 "allreduce_ms_standalone": ar_ms, "allreduce_bytes_per_step": 4 * nparam,
                 "allreduce_busbw_gbs": 2 * (world - 1) / world * 4 * nparam / (ar_ms * 1e-3) / 1e9,
                 "allreduce_overlap": 1.0 - min(1.0, exposed / ar_ms) if ar_ms > 0 else None}
    if rank == 0:
        value = world * b / (ms / 1e3)
        if contract:
            line = {"metric": "training instances/s (N=M=1024)", "value": value, "unit": "instances/s", "n_gpus": world,
                    "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None,
                    "dtype": ("fp32 layer modules on cuDNN/cuBLAS (FDA: fused kernels)" if layers else
                              "bf16 hi/lo split operands (3 MMAs per product), fp32 accumulate: every pointwise-MLP GEMM "
                              "(forward, dgrad, wgrad) and the FDA forward / backward on tcgen05"),
                    "data": "synthetic",
                    "config": {"workload": "config_LM-shaped training step through the FDA section: fwd + bwd + Adam, "
                                           "train-mode BatchNorm, losses of models/DCL_Net.py:265-303; entry = " +
                                           ("the towers' four pyramid levels per cloud (pointnet_sp three_nn + "
                                            "three_interpolate forward and backward: the level features carry gradients)"
                                            if entry == "backbone" else "point features (b*n, 480) per tower"),
                               "entry": entry,
                               "path": "nn layer modules (A/B)" if layers else "train_tail.mlp_stacks + dcl_fda_bwd",
                               "launch_mode": ("cuda_graph (forward + backward" + (" + Adam" if world == 1 else "") +
                                               " in one graph" + ("; flat-gradient NCCL all-reduce + Adam after the replay)"
                                                                  if world > 1 else ")")) if graph else "eager",
                               "B_per_gpu": b, "N": n, "M": n, "C": 64,
                               "parallelism": (f"data parallel x{world}: one NCCL all-reduce (AVG) of the flat {nparam}-element "
                                               "fp32 gradient buffer per step, after the graph replay" if graph else
                                               f"DDP x{world} (NCCL all-reduce of {nparam} fp32 gradients, bucketed, "
                                               "overlapped with backward)")},
                    "impl": "b200", "loss": float(loss.item()), "params": nparam, "gpu_launches": int(launches)}
            line.update(extra)
        else:
            line = dict({"what": "training step (fwd+bwd+Adam) through the FDA section, train mode", "n_gpus": world,
                         "B_per_gpu": b, "ms_per_step": ms, "instances_per_s": value, "loss": float(loss.item()),
                         "params": nparam}, **extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=40)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", action="store_true", help="nn layer modules on library GEMMs instead of the training kernels")
    ap.add_argument("--entry", default="backbone", choices=["backbone", "feats"])
    ap.add_argument("--graph", action="store_true", help="replay forward + backward as one CUDA graph")
    args = ap.parse_args()
    run(args.batch, args.steps, args.warmup, int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
        int(os.environ.get("LOCAL_RANK", 0)), layers=args.layers, entry=args.entry, graph=args.graph)


if __name__ == "__main__":
    main()
