"""Training path of the pointwise MLP stacks (dcl_net_b200/train_tail.py, csrc/train_ops.cu, dcl_pm_gemm strided
batch): forward values, input / weight / bias / BatchNorm gradients and running statistics against the same layers
evaluated by PyTorch in fp64 (what autograd gives the reference's nn.Conv1d / nn.Conv3d / nn.BatchNorm modules,
models/DCL_Net.py:56-151, models/Modules.py:58-97,173-201).  Tolerance: 2e-5 of the tensor's max (fp32 level; the
GEMMs run on bf16 hi/lo split operands, ~2^-17 per product)."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _rel(a, b):
    """Normwise relative error.  (A max-norm criterion is not usable for gradients through ReLU: a pre-activation within
    fp32 rounding of zero flips its gate against the fp64 graph and moves a few elements by a finite amount.)"""
    return (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)


def _relmax(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def _mods(kind_list, dims, dev, seed):
    """A Head_MultiLayerPerceptron / disengage-like module list for `kind_list`."""
    from dcl_net_b200.modules import BasicBlock_3DCONV, Head_MultiLayerPerceptron
    torch.manual_seed(seed)
    if kind_list[0] == "bn_relu":
        m = nn.Sequential(*[BasicBlock_3DCONV(dim_in=i, dim_out=o, size=1, bias=False, stride=1, padding=0, norm=True,
                                              act="relu", drop=0.0) for i, o in zip(dims[:-1], dims[1:])])
    else:
        act = ["none" if k == "linear" else "relu" for k in kind_list]
        m = Head_MultiLayerPerceptron(dims, act, [k == "relu_bn" for k in kind_list], [0.0] * len(kind_list))
    for mod in m.modules():                      # non-trivial BatchNorm parameters / buffers
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm3d)):
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.uniform_(-0.3, 0.3)
            mod.running_mean.uniform_(-0.2, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
    return m.to(dev).train()


def _run_ref(m64, x, is3d, gates):
    """The module in fp64 with every ReLU replaced by the gate the implementation under test took (a pre-activation
    within fp32 rounding of zero would otherwise open in one graph and close in the other, which moves whole rows of
    the gradients by a finite amount and — through BatchNorm — everything upstream by ~1e-3)."""
    gates = list(gates)
    x = x[:, :, :, None, None] if is3d else x
    for mod in (m64.modules() if not isinstance(m64, nn.Sequential) else m64.modules()):
        if isinstance(mod, (nn.Conv1d, nn.Conv3d, nn.BatchNorm1d, nn.BatchNorm3d)):
            x = mod(x)
        elif isinstance(mod, nn.ReLU):
            g = gates.pop(0)
            x = x * (g[:, :, :, None, None] if is3d else g).to(x.dtype)
    assert not gates
    return x.squeeze(-1).squeeze(-1) if is3d else x


@pytest.mark.parametrize("case", ["disengage", "head", "fuser", "one_tile"])
def test_mlp_stacks_vs_fp64(cuda_dev, case):
    from dcl_net_b200.train_tail import StackSpec, disengage_layers, head_layers, mlp_stacks
    dev = cuda_dev
    g = torch.Generator().manual_seed(5)
    if case == "disengage":       # two stacks sharing a point-major (b*n, 480) input, widths 256 / 64
        b, n = 3, 256
        x = torch.randn(b * n, 480, generator=g).to(dev).requires_grad_(True)
        mods = [_mods(["bn_relu"] * 2, [480, 256, 256], dev, 1), _mods(["bn_relu"] * 2, [480, 256, 64], dev, 2)]
        specs = lambda ms, xin: [StackSpec([(xin, "rm")], disengage_layers(m)) for m in ms]
        ref_in = lambda xin: [xin.view(b, n, 480).transpose(1, 2)] * 2
        is3d, widths = True, [256, 64]
    elif case == "head":          # conv-ReLU, conv-ReLU, conv (3 outputs, padded) on a concatenated input
        b, n = 2, 256
        xa = torch.randn(b, 64, n, generator=g).to(dev).requires_grad_(True)
        xb = torch.randn(b, 64, n, generator=g).to(dev).requires_grad_(True)
        x = (xa, xb)
        mods = [_mods(["relu", "relu", "linear"], [128, 128, 128, 3], dev, 3),
                _mods(["relu", "relu", "linear"], [128, 128, 128, 1], dev, 4)]
        is3d = False
    elif case == "fuser":         # conv-ReLU-BN x3, concatenated 256 + 256 input
        b, n = 2, 256
        xa = torch.randn(b, 256, n, generator=g).to(dev).requires_grad_(True)
        xb = torch.randn(b, 256, n, generator=g).to(dev).requires_grad_(True)
        x = (xa, xb)
        mods = [_mods(["relu_bn"] * 3, [512, 512, 512, 1024], dev, 6)]
        is3d = False
    else:                         # a single 128-row tile: the unpaired GEMM kernel
        b, n = 1, 128
        xa = torch.randn(b, 256, n, generator=g).to(dev).requires_grad_(True)
        x = (xa,)
        mods = [_mods(["relu", "relu", "linear"], [256, 256, 128, 3], dev, 7)]
        is3d = False

    mods64 = [copy.deepcopy(m).double() for m in mods]
    # ---- this implementation
    from dcl_net_b200 import train_tail
    train_tail.GATE_LOG = log = []
    if case == "disengage":
        outs = mlp_stacks(specs(mods, x), b, n)
    else:
        st = []
        widths = []
        for m in mods:
            lays, w = head_layers(m)
            widths.append(w)
            st.append(StackSpec([(t, "cm") for t in x], lays))
        outs = [o[:, :w] for o, w in zip(mlp_stacks(st, b, n), widths)]
    train_tail.GATE_LOG = None
    gates = [[m for l, s, m, _ in sorted(log, key=lambda e: e[0]) if s == si] for si in range(len(mods))]
    gouts = [torch.randn(o.shape, generator=g).to(dev) for o in outs]
    loss = sum((o * go).sum() for o, go in zip(outs, gouts))
    loss.backward()
    # ---- fp64 layers
    if case == "disengage":
        x64 = x.detach().double().requires_grad_(True)
        ins64 = ref_in(x64)
        outs64 = [_run_ref(m, i, True, gt) for m, i, gt in zip(mods64, ins64, gates)]
        leaves = [(x, x64)]
    else:
        xs64 = [t.detach().double().requires_grad_(True) for t in x]
        cat = torch.cat(xs64, dim=1)
        outs64 = [_run_ref(m, cat, False, gt) for m, gt in zip(mods64, gates)]
        leaves = list(zip(x, xs64))
    sum((o * go.double()).sum() for o, go in zip(outs64, gouts)).backward()

    report = []
    for o, o64 in zip(outs, outs64):
        report.append(("forward", _rel(o, o64), _relmax(o, o64), TOL))
    for t, t64 in leaves:
        report.append(("input grad", _rel(t.grad, t64.grad), _relmax(t.grad, t64.grad), TOL))
    for m, m64 in zip(mods, mods64):
        for (name, p), (_, p64) in zip(m.named_parameters(), m64.named_parameters()):
            assert p.grad is not None, name
            report.append((name, _rel(p.grad, p64.grad), _relmax(p.grad, p64.grad), 5 * TOL))
    print("\n".join(f"{case:10s} {w:28s} fro {e:.2e}  max {em:.2e}" for w, e, em, _ in report))
    bad = [r for r in report if not r[1] < r[3]]
    assert not bad, bad
    for m, m64 in zip(mods, mods64):
        for (name, buf), (_, buf64) in zip(m.named_buffers(), m64.named_buffers()):
            if buf.dtype.is_floating_point:
                assert _rel(buf, buf64) < TOL, f"{case} buffer {name}"
            else:
                assert int(buf) == int(buf64), name


def test_network_train_path_matches_layer_path(cuda_dev):
    """Network.forward_from_point_feats in train mode: tensor-core training path vs the nn layer modules (fp32
    library GEMMs) on the same parameters — outputs and every parameter gradient."""
    from dcl_net_b200.dcl_net import Network

    class Cfg:
        n_inp = n_tmp = 256
        unit_voxel_extent = [0.006] * 3

    torch.manual_seed(11)
    net = Network(Cfg, mode="train").to(cuda_dev).train()
    ref = copy.deepcopy(net)
    ref.use_train_kernels = False
    b, n = 4, 256
    g = torch.Generator().manual_seed(12)
    f_xc = torch.randn(b * n, 480, generator=g).to(cuda_dev)
    f_yo = torch.randn(b * n, 480, generator=g).to(cuda_dev)

    wr = torch.randn(b, 3, 3, generator=g).to(cuda_dev)

    def loss_fn(out):
        return (out["Xo_pred"].square().mean() + out["Yc_pred"].square().mean() + out["conf"].mean() +
                out["trans_pred"].square().mean() + (out["rot_pred"] * wr).sum(dim=(1, 2)).mean())

    a = f_xc.clone().requires_grad_(True), f_yo.clone().requires_grad_(True)
    r = f_xc.clone().requires_grad_(True), f_yo.clone().requires_grad_(True)
    out_a = net.forward_from_point_feats(a[0], a[1], b)
    out_r = ref.forward_from_point_feats(r[0], r[1], b)
    for k in ("Xo_pred", "Yc_pred", "conf", "trans_pred", "rot_pred", "F_Xo_p"):
        assert _rel(out_a[k], out_r[k]) < 1e-3, (k, _rel(out_a[k], out_r[k]))
    loss_fn(out_a).backward()
    loss_fn(out_r).backward()
    # loose, normwise: ReLU gates within fp32 rounding of zero differ between the two graphs (see _run_ref) and
    # train-mode BatchNorm spreads each one over the whole batch; the layer kinds are checked tightly above
    assert _rel(a[0].grad, r[0].grad) < 3e-2 and _rel(a[1].grad, r[1].grad) < 3e-2
    errs = sorted(((_rel(p.grad, q.grad), name) for (name, p), (_, q) in
                   zip(net.named_parameters(), ref.named_parameters()) if q.grad is not None), reverse=True)
    print(errs[:5])
    assert errs[0][0] < 5e-2, errs[:5]
    for (name, p), (_, q) in zip(net.named_buffers(), ref.named_buffers()):
        assert _rel(p.float(), q.float()) < 1e-4, name


def test_training_step_graph_replay_equals_eager(cuda_dev):
    """The whole training step (forward, backward, optimizer) captured in one CUDA graph (tools/train_step_ddp.py
    --graph) moves the parameters and BatchNorm buffers as the eagerly launched step does.  Plain SGD on purpose: its
    update is linear in the gradient, so rounding-level differences stay rounding-level (Adam's first steps move every
    weight by ~lr whatever the gradient's size, which turns a last-bit difference of a near-zero gradient into 2*lr)."""
    from dcl_net_b200.dcl_net import Network

    class Cfg:
        n_inp = n_tmp = 256
        unit_voxel_extent = [0.006] * 3

    b, n = 4, 256
    g = torch.Generator().manual_seed(5)
    f_xc = torch.randn(b * n, 480, generator=g).to(cuda_dev)
    f_yo = torch.randn(b * n, 480, generator=g).to(cuda_dev)
    tgt = torch.randn(b, 3, 3, generator=g).to(cuda_dev)

    def loss_fn(out):
        return (out["Xo_pred"].square().mean() + out["Yc_pred"].square().mean() + out["conf"].mean() +
                out["trans_pred"].square().mean() + (out["rot_pred"] * tgt).sum(dim=(1, 2)).mean())

    torch.manual_seed(9)
    net_e = Network(Cfg, mode="train").to(cuda_dev).train()
    net_g = copy.deepcopy(net_e)
    opt_e = torch.optim.SGD(net_e.parameters(), lr=1e-2)
    for _ in range(5):
        opt_e.zero_grad(set_to_none=True)
        loss_fn(net_e.forward_from_point_feats(f_xc, f_yo, b)).backward()
        opt_e.step()

    opt_g = torch.optim.SGD(net_g.parameters(), lr=1e-2)
    params = list(net_g.parameters())
    flat = torch.zeros(sum(p.numel() for p in params), device=cuda_dev)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()

    def fwd_bwd():
        flat.zero_()
        loss = loss_fn(net_g.forward_from_point_feats(f_xc, f_yo, b))
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fwd_bwd()
            opt_g.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fwd_bwd()
        opt_g.step()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    for (name, p), (_, q) in zip(net_g.named_parameters(), net_e.named_parameters()):
        assert _relmax(p, q) < 1e-4, (name, _relmax(p, q))
    for (name, p), (_, q) in zip(net_g.named_buffers(), net_e.named_buffers()):
        assert _relmax(p.float(), q.float()) < 1e-5, name
