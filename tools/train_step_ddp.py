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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=40)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = Network(Cfg, mode="train").to(dev).train()
    class FromPointFeats(torch.nn.Module):
        """DDP wraps a module whose forward takes the tensors, so that its reducer arms the gradient hooks."""

        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, f_xc, f_yo, nb):
            return self.inner.forward_from_point_feats(f_xc, f_yo, nb)

    wrapped = FromPointFeats(net)
    model = torch.nn.parallel.DistributedDataParallel(wrapped, device_ids=[local]) if world > 1 else wrapped
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    b, n = args.batch, 1024
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    f_xc = torch.randn(b * n, 480, device=dev, generator=g)
    f_yo = torch.randn(b * n, 480, device=dev, generator=g)
    pts_tmp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
    pts_inp = (torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.2
    q, _ = torch.linalg.qr(torch.randn(b, 3, 3, device=dev, generator=g))
    rot_gt = q * torch.det(q).sign().view(b, 1, 1)
    trans_gt = (torch.rand(b, 3, device=dev, generator=g) - 0.5) * 0.1
    fwd = lambda: model(f_xc, f_yo, b)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = losses(fwd(), pts_tmp, pts_inp, rot_gt, trans_gt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nparam = sum(p.numel() for p in net.parameters())
    if rank == 0:
        print(json.dumps({"what": "training step (fwd+bwd+Adam) through the FDA section, train mode",
                          "n_gpus": world, "B_per_gpu": b, "ms_per_step": float(ms.item()),
                          "instances_per_s": world * b / (float(ms.item()) / 1e3), "loss": float(loss.item()),
                          "allreduce_bytes_per_step": 4 * nparam if world > 1 else 0, "params": nparam}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
