"""CPU: the PyTorch restatement (oracle/torch_oracle.py) against the fixtures that oracle/make_golden.py produced
by executing the reference's own Python code (ortho9d2matrix, Aligner, Network.forward's FDA section, Refiner,
Ops_GetPointFeat_spconv).  Same torch build => bit-exact on CPU; a small tolerance is allowed for BLAS threading."""
import types

import numpy as np
import torch

from oracle import cpu_oracle, torch_oracle as T
from oracle.make_golden import param_checksum
from dcl_testutil import GOLDEN


def test_ortho9d_golden():
    gold = np.load(f"{GOLDEN}/model_ortho9d.npz")
    raw = torch.from_numpy(gold["raw"])
    R = T.ortho9d2matrix(raw[:, :3], raw[:, 3:6], raw[:, 6:])
    assert np.allclose(R.numpy(), gold["R"], atol=1e-6)
    eye = torch.eye(3).expand_as(R)
    assert (R @ R.transpose(1, 2) - eye).abs().max() < 1e-5 and (torch.det(R) - 1).abs().max() < 1e-5


def test_project_so3_is_the_same_projection():
    g = torch.Generator().manual_seed(1)
    raw = torch.randn(32, 9, generator=g)
    m = torch.stack([T.normalize_vector(raw[:, i:i + 3]) for i in (0, 3, 6)], dim=2)
    a = T.ortho9d2matrix(raw[:, :3], raw[:, 3:6], raw[:, 6:])
    s = torch.linalg.svdvals(m.double())
    ok = (s[:, 1] - s[:, 2] > 1e-2) & (s[:, 2] > 1e-3)
    assert T.rotation_angle_deg(a, T.project_so3(m))[ok].max() < 0.01


def test_aligner_golden():
    gold = np.load(f"{GOLDEN}/model_aligner.npz")
    g = torch.Generator().manual_seed(int(gold["seed"]))
    ri1, ri2, re2 = (torch.randn(2, 64, 128, generator=g).relu(), torch.randn(2, 64, 192, generator=g).relu(),
                     torch.randn(2, 256, 192, generator=g))
    e, a = T.aligner(ri1, ri2, re2)
    assert np.allclose(e.numpy(), gold["RE_embed"], rtol=1e-5, atol=1e-5)
    assert np.allclose(a.sum(1).numpy(), gold["A_colsum"], atol=1e-5)
    assert np.allclose(a[:, ::16, ::16].numpy(), gold["A_sample"], atol=1e-6)


def test_tail_golden():
    gold = np.load(f"{GOLDEN}/model_tail_b2_n128.npz")
    torch.manual_seed(int(gold["seed_weights"]))
    net = T.TailNetwork(mode="train").eval()
    assert abs(param_checksum(net) - float(gold["param_checksum"])) <= 1e-9 * abs(float(gold["param_checksum"]))
    g = torch.Generator().manual_seed(int(gold["seed_inputs"]))
    f_xc, f_yo = torch.randn(256, 480, generator=g), torch.randn(256, 480, generator=g)
    with torch.no_grad():
        out = net(f_xc, f_yo, 2, 128, 128)
    for k in ("rot_pred", "trans_pred", "conf", "F_Xo_p", "Xo_pred", "Yc_pred"):
        assert np.allclose(out[k].numpy(), gold[k], rtol=1e-4, atol=1e-5), k
    for k in ("F_Yc_p", "F_Xo_m", "F_Yc_m"):
        assert np.allclose(out["_debug"][k].numpy(), gold[k], rtol=1e-4, atol=1e-5), k


def test_refiner_golden():
    gold = np.load(f"{GOLDEN}/model_refiner.npz")
    torch.manual_seed(int(gold["seed_weights"]))
    net = T.RefinerNet().eval()
    g = torch.Generator().manual_seed(int(gold["seed_inputs"]))
    inp = {"input_features": torch.randn(2, 259, 1024, generator=g), "conf": torch.rand(2, 2048, generator=g)}
    with torch.no_grad():
        out = net(inp)
    assert np.allclose(out["rot_pred"].numpy(), gold["rot_pred"], atol=1e-5)
    assert np.allclose(out["trans_pred"].numpy(), gold["trans_pred"], atol=1e-6)


def test_losses_golden():
    """CD_Dis / ADD-S restatements against values produced by the reference's own code (oracle/make_golden.py)."""
    gold = np.load(f"{GOLDEN}/model_losses.npz")
    g = torch.Generator().manual_seed(int(gold["seed"]))
    pa, pb = torch.rand(2, 96, 3, generator=g), torch.rand(2, 96, 3, generator=g)
    pa[:, 0] = pb[:, 2]
    assert np.array_equal(T.cd_dis(pa, pb).numpy(), gold["cd_dis"])
    assert np.array_equal(T.adds(pa, pb).numpy(), gold["adds"])


def test_point_feats_golden():
    gold = np.load(f"{GOLDEN}/model_point_feats.npz")
    g = torch.Generator().manual_seed(int(gold["seed"]))
    bsz, npts = 3, 200
    points = (torch.rand(bsz * npts, 3, generator=g) - 0.5) * 0.2
    batch_ids = torch.arange(bsz).repeat_interleave(npts)
    levels = []
    for li, (scale, ch) in enumerate(zip([2, 4, 6, 8], [32, 64, 128, 256])):
        mv = [90, 40, 20, 6][li] * bsz
        ind = torch.cat([torch.randint(0, bsz, (mv, 1), generator=g),
                         torch.randint(0, 64 // scale, (mv, 3), generator=g)], 1).int()
        ind = torch.unique(ind, dim=0)
        ind = ind[torch.randperm(ind.shape[0], generator=g)]
        levels.append((torch.randn(ind.shape[0], ch, generator=g), ind))
    c_nn = lambda u, k: tuple(map(torch.from_numpy, cpu_oracle.sp_three_nn(u.numpy(), k.numpy())))
    got = T.get_point_feats(points, batch_ids, levels, [0.006] * 3, three_nn=c_nn)
    assert np.allclose(got.numpy(), gold["point_feats"], rtol=1e-5, atol=1e-6)
    # the pure-torch search (what bench.py times as the CPU baseline) finds the same neighbours
    got_t = T.get_point_feats(points, batch_ids, levels, [0.006] * 3)
    assert np.allclose(got_t.numpy(), gold["point_feats"], rtol=1e-4, atol=1e-5)


def test_weighted_kabsch_recovers_motion():
    g = torch.Generator().manual_seed(5)
    src = torch.rand(3, 500, 3, generator=g)
    q, _ = torch.linalg.qr(torch.randn(3, 3, 3, generator=g))
    q = q * torch.det(q).sign().view(3, 1, 1)
    t = torch.rand(3, 3, generator=g)
    dst = src @ q.transpose(1, 2) + t.unsqueeze(1)
    R, tt = T.weighted_kabsch(src, dst, torch.rand(3, 500, generator=g))
    assert T.rotation_angle_deg(R, q).max() < 1e-4 and (tt - t.double()).abs().max() < 1e-6


def test_stage2_loop_composition():
    """With a refiner that returns identity updates the pose is unchanged; composition order is R<-R dR, t<-R dt+t."""
    class Fixed(torch.nn.Module):
        def __init__(self, dR, dt):
            super().__init__()
            self.dR, self.dt = dR, dt

        def forward(self, d):
            b = d["input_features"].shape[0]
            return {"rot_pred": self.dR.expand(b, 3, 3), "trans_pred": self.dt.expand(b, 3)}
    g = torch.Generator().manual_seed(6)
    pts, f, conf = torch.rand(2, 64, 3, generator=g), torch.randn(2, 256, 64, generator=g), torch.rand(2, 128, generator=g)
    q, _ = torch.linalg.qr(torch.randn(2, 3, 3, generator=g))
    t = torch.rand(2, 3, generator=g)
    r1, t1 = T.stage2_refine(Fixed(torch.eye(3), torch.zeros(3)), pts, q, t, f, conf, 3)
    assert torch.allclose(r1, q) and torch.allclose(t1, t)
    dR = torch.tensor([[0., -1, 0], [1, 0, 0], [0, 0, 1]])
    dt = torch.tensor([0.1, 0.2, 0.3])
    r2, t2 = T.stage2_refine(Fixed(dR, dt), pts, q, t, f, conf, 1)
    assert torch.allclose(r2, q @ dR, atol=1e-6) and torch.allclose(t2, (q @ dt) + t, atol=1e-6)
