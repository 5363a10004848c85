"""GPU diagnostic: where does the training-step gradient error of the fused path come from?
Runs the step of tests/test_gpu_pose_model.py::test_training_step_gradients_match_oracle with (a) the fused FDA,
(b) a torch fp32 FDA forward with FdaAlignFunction's backward, and prints forward / gradient errors against fp64."""
import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dcl_net_b200 import modules as M
from dcl_net_b200.dcl_net import Network
import dcl_net_b200.dcl_net as DN
from oracle import torch_oracle as T
from test_gpu_pose_model import Cfg
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
b, n = 2, 256
torch.manual_seed(21)
oracle_net = T.TailNetwork(mode="train").train()
net = Network(Cfg(n), mode="train").train()
net.load_state_dict(oracle_net.state_dict(), strict=False)
oracle_net, net = oracle_net.to(dev), net.to(dev)
g = torch.Generator().manual_seed(22)
f_xc, f_yo = torch.randn(b * n, 480, generator=g).to(dev), torch.randn(b * n, 480, generator=g).to(dev)
pts = ((torch.rand(b, n, 3, generator=g) - 0.5) * 0.2).to(dev)
q, _ = torch.linalg.qr(torch.randn(b, 3, 3, generator=g))
rot_gt = (q * torch.det(q).sign().view(b, 1, 1)).to(dev)
trans_gt = ((torch.rand(b, 3, generator=g) - 0.5) * 0.1).to(dev)

def loss_fn(out, dt):
    p, r, t = pts.to(dt), rot_gt.to(dt), trans_gt.to(dt)
    posed = torch.bmm(p, out["rot_pred"].transpose(1, 2)) + out["trans_pred"].unsqueeze(1)
    posed_gt = torch.bmm(p, r.transpose(1, 2)) + t.unsqueeze(1)
    l_pose = torch.norm(posed - posed_gt, dim=2).mean()
    l_xo = torch.norm(out["Xo_pred"] - p, dim=2)
    l_yc = torch.norm(out["Yc_pred"] - posed_gt, dim=2)
    conf = out["conf"]
    l_conf = torch.mean(torch.cat([l_xo, l_yc], dim=1).detach() * conf - 0.01 * torch.log(conf))
    return l_pose + 5 * l_xo.mean() + l_yc.mean() + l_conf

oracle64 = copy.deepcopy(oracle_net).double()
cc = f_xc.double().requires_grad_(True), f_yo.double().requires_grad_(True)
out64 = oracle64(cc[0], cc[1], b, n, n)
loss_fn(out64, torch.float64).backward()
p64 = dict(oracle64.named_parameters())

def run(tag):
    net.zero_grad()
    a = f_xc.clone().requires_grad_(True), f_yo.clone().requires_grad_(True)
    out = net.forward_from_point_feats(a[0], a[1], b)
    loss_fn(out, torch.float32).backward()
    print(f"== {tag}")
    for k in ("Xo_pred", "Yc_pred", "conf", "rot_pred", "trans_pred"):
        s = out64[k].abs().max().item()
        print(f"  fwd {k:10s} {(out[k].double() - out64[k]).abs().max().item() / s:.2e}")
    errs = []
    for i in range(2):
        s = cc[i].grad.abs().max().item()
        errs.append((((a[i].grad.double() - cc[i].grad).abs().max().item()) / s, f"input{i}"))
    for name, p in net.named_parameters():
        if p.grad is None or p64[name].grad is None:
            continue
        s = p64[name].grad.abs().max().item()
        if s > 0:
            errs.append(((p.grad.double() - p64[name].grad).abs().max().item() / s, name))
    errs.sort(reverse=True)
    for e, nm in errs[:8]:
        print(f"  grad {e:.2e} {nm}")

run("fused FDA forward + FdaAlignFunction backward")

class TorchFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, RI_1, RI_2, RE_2):
        S = torch.bmm(RI_2.transpose(1, 2), RI_1)
        lse = torch.logsumexp(S, dim=1)
        A = torch.exp(S - lse.unsqueeze(1))
        ctx.save_for_backward(RI_1, RI_2, RE_2, lse)
        return torch.bmm(RE_2, A), torch.bmm(RI_2, A), lse
    backward = M.FdaAlignFunction.backward

def torch_fwd(RI_1, RI_2, RE_2, return_lse=False):
    out = TorchFwd.apply(RI_1, RI_2, RE_2)
    return out if return_lse else out[:2]
for mod in (M, DN):
    if hasattr(mod, "fda_align"):
        mod.fda_align = torch_fwd
run("torch fp32 FDA forward + FdaAlignFunction backward")
