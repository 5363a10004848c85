"""CPU: host-side logic — instance sharding, flat-tensor re-basing, and the N>1 pose gather over gloo (world 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcl_net_b200 import sharding
from dcl_testutil import flat_bxyz


def test_instance_range_partitions():
    for total in (0, 1, 7, 32, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [sharding.instance_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_flat_rebases_batch_ids():
    rows = flat_bxyz(0, 8, 10)
    parts = []
    for r in range(4):
        lo, hi = sharding.instance_range(8, r, 4)
        part, keep = sharding.shard_flat(rows, lo, hi)
        assert part[:, 0].min() >= 0 and part[:, 0].max() < hi - lo
        assert torch.equal(part[:, 1:], rows[keep][:, 1:])
        parts.append(int(keep.sum()))
    assert sum(parts) == rows.shape[0]


def test_gather_poses_single_process_is_identity():
    r, t = torch.eye(3).repeat(5, 1, 1), torch.zeros(5, 3)
    r2, t2 = sharding.gather_poses(r, t)
    assert r2 is r and t2 is t


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.instance_range(total, rank, world)
        ids = torch.arange(lo, hi, dtype=torch.float32)
        rot = ids.view(-1, 1, 1) * torch.ones(1, 3, 3)
        trans = ids.view(-1, 1) + torch.tensor([0.1, 0.2, 0.3])
        r, t = sharding.gather_poses(rot, trans)
        assert r.shape == (total, 3, 3) and t.shape == (total, 3)
        if total % world == 0:  # equal shards: the single asynchronous all_gather gives the same answer
            r2, t2 = sharding.gather_poses(rot, trans, equal_shards=True)
            assert torch.equal(r2, r) and torch.equal(t2, t)
        assert torch.equal(r[:, 0, 0], torch.arange(total, dtype=torch.float32))
        assert torch.allclose(t[:, 2], torch.arange(total, dtype=torch.float32) + 0.3)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_poses_world2_gloo(total):
    mp.spawn(_worker, args=(2, _free_port(), total), nprocs=2, join=True)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) runs without a GPU and prints one
    JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, check=True)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_train_tail_layer_specs():
    """train_tail maps the reference's module trees to layer kinds without touching a GPU: disengage blocks are
    conv -> BN -> ReLU, Head_MultiLayerPerceptron layers conv -> [ReLU] -> [BN], a narrow last layer is zero-padded to
    64 outputs by differentiable torch ops (so its gradient flows back to the unpadded parameter)."""
    import torch
    from dcl_net_b200.dcl_net import Network
    from dcl_net_b200.train_tail import disengage_layers, head_layers

    class Cfg:
        n_inp = n_tmp = 256
        unit_voxel_extent = [0.006] * 3

    net = Network(Cfg, mode="train")
    dis = disengage_layers(net.disengage_Xc_m1)
    assert [l[0] for l in dis] == ["bn_relu", "bn_relu"]
    assert dis[0][1].shape[:2] == (256, 480) and dis[1][1].shape[:2] == (64, 256) and dis[0][2] is None
    assert isinstance(dis[0][4], torch.nn.ReLU)
    lays, width = head_layers(net.regressor_conf)
    assert [l[0] for l in lays] == ["relu", "relu", "linear"] and width == 1
    assert lays[2][1].shape == (64, 128) and lays[2][2].shape == (64,)
    assert torch.equal(lays[2][1][:1], net.regressor_conf.layers[4].weight.reshape(1, 128)) and not lays[2][1][1:].any()
    lays[2][1].sum().backward()
    assert net.regressor_conf.layers[4].weight.grad is not None
    lays, width = head_layers(net.neck_fuser)
    assert [l[0] for l in lays] == ["relu_bn"] * 3 and width == 1024
    assert all(isinstance(l[3], torch.nn.BatchNorm1d) for l in lays)
    lays, width = head_layers(net.regressor_Xo)
    assert width == 3 and lays[-1][1].shape == (64, 128)


def test_so3_projection_backward_closed_form_matches_svd_autograd():
    """dcl_net.so3_projection_backward (what ProjectSO3Function.backward runs) against autograd through the
    reference's torch.svd formula (models/DCL_Net.py:22-35), fp64 on CPU, proper and reflected inputs."""
    import torch
    from dcl_net_b200.dcl_net import so3_projection_backward
    from oracle import torch_oracle as T
    g = torch.Generator().manual_seed(3)
    for reflect in (False, True):
        m = torch.randn(16, 3, 3, generator=g, dtype=torch.float64)
        if reflect:
            m = m * torch.where(torch.det(m) > 0, -1.0, 1.0).view(-1, 1, 1)
        m.requires_grad_(True)
        u, _, v = torch.svd(m)
        sigma = torch.ones(16, 3, dtype=torch.float64)
        sigma[:, -1] = torch.bmm(u, v.transpose(1, 2)).det()
        r = u @ torch.diag_embed(sigma) @ v.transpose(1, 2)
        gr = torch.randn(16, 3, 3, generator=g, dtype=torch.float64)
        (r * gr).sum().backward()
        mine = so3_projection_backward(m.detach(), r.detach(), gr)
        assert (mine - m.grad).abs().max().item() < 1e-10 * max(1.0, m.grad.abs().max().item())
        assert torch.allclose(r.detach(), T.project_so3(m.detach()), atol=1e-12)


def _dp_worker(rank, world, port):
    """Data-parallel step with the flat gradient buffer (sharding.flat_grad_buffer / average_gradients): after the
    all-reduce every rank holds the mean of the per-rank gradients, in the parameters' own .grad views."""
    import torch.distributed as dist
    from dcl_net_b200.sharding import average_gradients, flat_grad_buffer
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                    # same parameters on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        flat = flat_grad_buffer(model.parameters())
        assert flat.numel() == sum(p.numel() for p in model.parameters())
        views = [p.grad.data_ptr() for p in model.parameters()]
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 + rank))   # different data per rank
        for _ in range(2):                                      # second step: the buffer is zeroed, not replaced
            flat.zero_()
            model(x).square().sum().backward()
            assert [p.grad.data_ptr() for p in model.parameters()] == views          # accumulated in place
            local = flat.clone()
            average_gradients(flat)
            gathered = [torch.zeros_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            assert torch.allclose(flat, sum(gathered) / world, atol=1e-6)
            assert torch.equal(torch.cat([p.grad.reshape(-1) for p in model.parameters()]), flat)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_data_parallel_world2_gloo():
    mp.spawn(_dp_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_training_pointwise_backward_formulas():
    """The pointwise backward transforms of csrc/train_ops.cu (modes DCL_TR_BWD_BN_RELU / DCL_TR_BWD_RELU_BN, with the
    sums of tr_bn_bwd_reduce_kernel), restated in torch fp64, against autograd through train-mode BatchNorm + ReLU in
    the two orders the reference uses (models/Modules.py:58-97: conv -> BN -> ReLU; :173-201: conv -> ReLU -> BN)."""
    import torch
    g = torch.Generator().manual_seed(11)
    b, c, n, eps = 3, 8, 40, 1e-5
    gamma = torch.rand(c, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(c, generator=g, dtype=torch.float64) * 0.3
    dy = torch.randn(b, c, n, generator=g, dtype=torch.float64)
    cnt = b * n

    def stats(u):
        mean = u.mean(dim=(0, 2))
        rstd = 1.0 / torch.sqrt(u.var(dim=(0, 2), unbiased=False) + eps)
        scale = gamma * rstd
        return mean, rstd, scale, beta - mean * scale

    v = lambda t: t.view(1, c, 1)
    # conv -> BN -> ReLU: U = conv output
    u = torch.randn(b, c, n, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.relu(torch.nn.functional.batch_norm(u, None, None, gamma, beta, True, 0.1, eps))
    y.backward(dy)
    mean, rstd, scale, shift = stats(u.detach())
    xhat = (u.detach() - v(mean)) * v(rstd)
    gte = torch.where(u.detach() * v(scale) + v(shift) > 0, dy, torch.zeros_like(dy))
    s1, s2 = gte.sum(dim=(0, 2)), (gte * xhat).sum(dim=(0, 2))
    dz = v(scale) * (gte - v(s1) / cnt - xhat * v(s2) / cnt)
    assert torch.allclose(dz, u.grad, atol=1e-12)
    # conv + bias -> ReLU -> BN: U = relu(conv output + bias); z = the pre-activation
    z = torch.randn(b, c, n, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.batch_norm(torch.relu(z), None, None, gamma, beta, True, 0.1, eps)
    y.backward(dy)
    uu = torch.relu(z.detach())
    mean, rstd, scale, shift = stats(uu)
    xhat = (uu - v(mean)) * v(rstd)
    s1, s2 = dy.sum(dim=(0, 2)), (dy * xhat).sum(dim=(0, 2))
    dz = torch.where(uu > 0, v(scale) * (dy - v(s1) / cnt - xhat * v(s2) / cnt), torch.zeros_like(dy))
    assert torch.allclose(dz, z.grad, atol=1e-12)
