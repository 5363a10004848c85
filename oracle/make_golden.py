"""TEST INFRASTRUCTURE — writes tests/golden/model_*.npz from the REFERENCE's own Python code.

Runs only where /root/reference exists (the build container, CPU).  The reference model files
import spconv / ipdb / pointgroup_ops / compiled extensions and hard-code .cuda(); we
  * register parameter-free stand-ins for those imports in sys.modules,
  * make Tensor.cuda()/Module.cuda() the identity,
and then execute the reference's code itself:
  * models.DCL_Net.ortho9d2matrix, utils.transform3D.normalize_vector      (pose SVD)
  * models.Modules.Aligner                                                  (FDA)
  * models.DCL_Net.Network: constructed for real; the body of Network.forward from
    "# bi-direction FDA" to the prediction dict (DCL_Net.py:187-235) is extracted with
    inspect and exec'd verbatim on synthetic point features                 (FDA section)
  * models.refiner.Refiner.forward                                          (refiner)
  * models.Modules.Ops_GetPointFeat_spconv with pointnet_sp ops bound to the C oracle (glue)
Inputs are regenerated from seeds by the tests; the fixtures store outputs (+ a parameter
checksum proving that oracle.torch_oracle.TailNetwork / RefinerNet built under the same seed
carry the same weights as the reference modules).

    python -m oracle.make_golden
"""
import inspect
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DCL_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    class _NoParam(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    spconv = types.ModuleType("spconv")
    for name in ("SparseConv3d", "SubMConv3d", "SparseAvgPool3d", "SparseConvTensor"):
        setattr(spconv, name, _NoParam)
    spconv.SparseSequential = lambda *mods: torch.nn.Sequential(*mods)
    sys.modules["spconv"] = spconv
    sys.modules["ipdb"] = types.ModuleType("ipdb")
    for name in ("libs.pointnet_sp.pointnet2_cuda", "libs.pointnet_lib.pointnet2_cuda"):
        sys.modules[name] = types.ModuleType(name)
    pg = types.ModuleType("libs.pointgroup_ops.functions")
    pg.pointgroup_ops = types.ModuleType("pointgroup_ops")
    sys.modules["libs.pointgroup_ops"] = types.ModuleType("libs.pointgroup_ops")
    sys.modules["libs.pointgroup_ops.functions"] = pg
    sys.modules["PG_OP"] = types.ModuleType("PG_OP")
    tbx = types.ModuleType("tensorboardX")  # utils/__init__.py pulls in the training logger
    tbx.SummaryWriter = object
    sys.modules["tensorboardX"] = tbx
    if REF not in sys.path:
        sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import models.DCL_Net as ref_net  # noqa
    import models.Modules as ref_mod  # noqa
    import models.refiner as ref_refiner  # noqa
    import utils.transform3D as ref_t3d  # noqa
    return ref_net, ref_mod, ref_refiner, ref_t3d


def param_checksum(module):
    acc = 0.0
    for k, v in sorted(module.state_dict().items()):
        if v.dtype.is_floating_point:
            acc += float(v.double().abs().sum()) + 3.0 * float(v.double().sum())
    return acc


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def reference_fda_section(ref_net_mod, net, f_xc_flat, f_yo_flat, b, n_inp, n_tmp):
    """Execute DCL_Net.py's own FDA section (source lines between the two markers)."""
    src = inspect.getsource(ref_net_mod.Network.forward)
    start = src.index("# bi-direction FDA")
    stop = src.index("if self.mode == 'test'")
    body = textwrap.dedent(" " * 8 + src[start:stop])
    class _Yo(torch.nn.Module):  # the only use inside the section: F_Yo = self.stage1_get_point_feats(...)
        def forward(self, *a, **k):
            return f_yo_flat
    net.stage1_get_point_feats = _Yo()
    pts = torch.zeros(b * n_tmp, 3)
    env = {"self": net, "torch": torch, "b": b, "F_Xc": f_xc_flat, "points_tmp": pts,
           "points_inp": torch.zeros(b * n_inp, 3), "RGB_tmp": torch.zeros(b * n_tmp, 3),
           "RGB_inp": torch.zeros(b * n_inp, 3), "ortho9d2matrix": ref_net_mod.ortho9d2matrix}
    for k in ("feats1_tmp", "feats2_tmp", "feats3_tmp", "feats4_tmp"):
        env[k] = None
    exec(body, env)
    return env


def loss_inputs(seed, b, n):
    """Deterministic inputs of the loss modules (shared by make_golden and the tests)."""
    g = torch.Generator().manual_seed(seed + 1)
    def rot(k):
        q, _ = torch.linalg.qr(torch.randn(k, 3, 3, generator=g))
        return q * torch.det(q).sign().view(k, 1, 1)
    pred = {"rot_pred": rot(b), "trans_pred": 0.1 * torch.randn(b, 3, generator=g),
            "sym_flag": torch.tensor([0.0, 1.0, 1.0][:b] + [0.0] * max(0, b - 3)),
            "conf": torch.rand(b, 2 * n, generator=g) * 0.9 + 0.05,
            "Xo_pred": 0.2 * torch.randn(b, n, 3, generator=g), "Yc_pred": 0.2 * torch.randn(b, n, 3, generator=g)}
    gt = {"rot_gt": rot(b), "trans_gt": 0.1 * torch.randn(b, 3, generator=g),
          "points_tmp": 0.2 * torch.randn(b, n, 3, generator=g), "points_inp": 0.2 * torch.randn(b, n, 3, generator=g)}
    return {"pred": pred, "gt": gt, "refiner_pred": {"rot_pred": rot(b), "trans_pred": 0.01 * torch.randn(b, 3, generator=g)},
            "rot_cur": rot(b), "trans_cur": 0.1 * torch.randn(b, 3, generator=g)}


def main():
    os.makedirs(GOLD, exist_ok=True)
    sys.path.insert(0, ROOT)
    from oracle import cpu_oracle, torch_oracle as T
    ref_net_mod, ref_mod, ref_refiner, ref_t3d = import_reference()
    torch.set_num_threads(8)

    # ---------------- pose SVD: the reference's ortho9d2matrix itself
    g = torch.Generator().manual_seed(100)
    raw = torch.randn(64, 9, generator=g)
    raw[5] *= 1e-3
    raw[6] *= 50.0
    raw[7, 6:] = raw[7, :3] * 0.7 + 1e-3 * raw[7, 6:]  # nearly coplanar columns
    with torch.no_grad():
        r_ref = ref_net_mod.ortho9d2matrix(raw[:, :3], raw[:, 3:6], raw[:, 6:])
        r_mine = T.ortho9d2matrix(raw[:, :3], raw[:, 3:6], raw[:, 6:])
    assert torch.equal(r_ref, r_mine), "torch_oracle.ortho9d2matrix deviates from the reference"
    np.savez_compressed(os.path.join(GOLD, "model_ortho9d.npz"), raw=raw.numpy(), R=r_ref.numpy())

    # ---------------- Aligner
    g = torch.Generator().manual_seed(101)
    ri1, ri2, re2 = (torch.randn(2, 64, 128, generator=g).relu(), torch.randn(2, 64, 192, generator=g).relu(),
                     torch.randn(2, 256, 192, generator=g))
    with torch.no_grad():
        e_ref, a_ref = ref_mod.Aligner()(ri1, ri2, re2)
        e_mine, a_mine = T.aligner(ri1, ri2, re2)
    assert torch.equal(e_ref, e_mine) and torch.equal(a_ref, a_mine)
    np.savez_compressed(os.path.join(GOLD, "model_aligner.npz"), seed=101, RE_embed=e_ref.numpy(),
                        A_colsum=a_ref.sum(1).numpy(), A_sample=a_ref[:, ::16, ::16].numpy())

    # ---------------- FDA section of Network.forward, reference code verbatim
    cfg = _Cfg(voxelization_mode=4, unit_voxel_extent=[0.006, 0.006, 0.006], n_inp=128, n_tmp=128,
               backbone=_Cfg(downsample_by_pooling=True, kernel_size=3))
    b, n = 2, 128
    torch.manual_seed(7)
    net = ref_net_mod.Network(cfg, mode="train").eval()
    torch.manual_seed(7)
    mine = T.TailNetwork(mode="train").eval()
    missing = mine.load_state_dict(net.state_dict(), strict=False)
    assert not missing.missing_keys, missing
    torch.manual_seed(7)
    mine_seeded = T.TailNetwork(mode="train").eval()
    for (k1, v1), (k2, v2) in zip(sorted(mine.state_dict().items()), sorted(mine_seeded.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2), f"same-seed construction differs at {k1}"
    g = torch.Generator().manual_seed(102)
    f_xc = torch.randn(b * n, 480, generator=g)
    f_yo = torch.randn(b * n, 480, generator=g)
    with torch.no_grad():
        env = reference_fda_section(ref_net_mod, net, f_xc, f_yo, b, n, n)
        out = mine(f_xc, f_yo, b, n, n)
    pairs = {"rot_pred": env["rot_pred"], "trans_pred": env["trans_pred"], "conf": env["conf"].squeeze(1),
             "F_Xo_p": env["F_Xo_p"], "Xo_pred": env["Xo_pred"].transpose(1, 2),
             "Yc_pred": env["Yc_pred"].transpose(1, 2)}
    for k, v in pairs.items():
        assert torch.allclose(v, out[k], rtol=0, atol=0), f"TailNetwork deviates from reference code at {k}"
    np.savez_compressed(os.path.join(GOLD, "model_tail_b2_n128.npz"), seed_weights=7, seed_inputs=102,
                        param_checksum=param_checksum(mine),
                        rot_pred=env["rot_pred"].numpy(), trans_pred=env["trans_pred"].numpy(),
                        conf=env["conf"].squeeze(1).numpy(), F_Xo_p=env["F_Xo_p"].numpy(),
                        F_Yc_p=env["F_Yc_p"].numpy(), F_Xo_m=env["F_Xo_m"].numpy(), F_Yc_m=env["F_Yc_m"].numpy(),
                        Xo_pred=pairs["Xo_pred"].numpy(), Yc_pred=pairs["Yc_pred"].numpy(),
                        ortho9d=env["ortho9d_pred"].numpy())

    # ---------------- Refiner.forward
    torch.manual_seed(8)
    ref = ref_refiner.Refiner(cfg).eval()
    torch.manual_seed(8)
    mine_r = T.RefinerNet().eval()
    for (k1, v1), (k2, v2) in zip(sorted(ref.state_dict().items()), sorted(mine_r.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2), f"RefinerNet same-seed weights differ at {k1}"
    g = torch.Generator().manual_seed(103)
    inp = {"input_features": torch.randn(2, 259, 1024, generator=g), "conf": torch.rand(2, 2048, generator=g),
           "obj_idx": None}
    with torch.no_grad():
        o_ref, o_mine = ref(inp), mine_r(inp)
    assert torch.equal(o_ref["rot_pred"], o_mine["rot_pred"]) and torch.equal(o_ref["trans_pred"], o_mine["trans_pred"])
    np.savez_compressed(os.path.join(GOLD, "model_refiner.npz"), seed_weights=8, seed_inputs=103,
                        param_checksum=param_checksum(mine_r),
                        rot_pred=o_ref["rot_pred"].numpy(), trans_pred=o_ref["trans_pred"].numpy())

    # ---------------- point-feature interpolation glue (reference Modules code, C-oracle ops)
    class _SpOps:
        @staticmethod
        def three_nn(unknown, known):
            d2, idx = cpu_oracle.sp_three_nn(unknown.numpy(), known.numpy())
            return torch.sqrt(torch.from_numpy(d2)), torch.from_numpy(idx)

        @staticmethod
        def three_interpolate(feats, idx, weight):
            return torch.from_numpy(cpu_oracle.sp_three_interpolate(feats.numpy(), idx.numpy(), weight.numpy()))

    ref_mod.pointnet2_utils_sp = _SpOps
    g = torch.Generator().manual_seed(104)
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
        t = types.SimpleNamespace(features=torch.randn(ind.shape[0], ch, generator=g), indices=ind)
        levels.append(t)
    getter = ref_mod.Ops_GetPointFeat_spconv(scale_lists=[2, 4, 6, 8], unit_voxel_extent=np.array([0.006] * 3),
                                             voxel_num_limit=[64, 64, 64])
    with torch.no_grad():
        pf_ref = getter(points, batch_ids, *levels)
        pf_mine = T.get_point_feats(points, batch_ids, [(l.features, l.indices) for l in levels], [0.006] * 3,
                                    three_nn=lambda u, k: tuple(map(torch.from_numpy, cpu_oracle.sp_three_nn(u.numpy(), k.numpy()))))
    err = (pf_ref - pf_mine).abs().max().item()
    assert err <= 1e-6 * pf_ref.abs().max().item(), f"get_point_feats deviates: {err}"
    np.savez_compressed(os.path.join(GOLD, "model_point_feats.npz"), seed=104, point_feats=pf_ref.numpy())
    # ---------------- loss / metric distances: the reference's own CD_Dis and the ADD-S expression of its test driver
    g = torch.Generator().manual_seed(105)
    pa, pb = torch.rand(2, 96, 3, generator=g), torch.rand(2, 96, 3, generator=g)
    pa[:, 0] = pb[:, 2]
    cd_ref = ref_net_mod.losses.CD_Dis(None, pa, pb)
    adds_ref = torch.mean(torch.min(torch.norm(pa.unsqueeze(2) - pb.unsqueeze(1), dim=3), 2)[0], dim=1)  # test_YCBV_stage1.py:188
    assert torch.equal(cd_ref, T.cd_dis(pa, pb)) and torch.equal(adds_ref, T.adds(pa, pb))
    # the reference's loss modules themselves (.cuda() is an identity here), on a mixed symmetric / asymmetric batch
    lb, ln = 3, 64
    li = loss_inputs(105, lb, ln)
    ref_losses = ref_net_mod.losses(None)(li["pred"], li["gt"])
    ref_losses_ref = ref_refiner.losses_refiner(None)(li["refiner_pred"], li["trans_cur"], li["rot_cur"],
                                                      li["gt"]["points_tmp"], li["pred"]["sym_flag"], li["gt"])
    np.savez_compressed(os.path.join(GOLD, "model_losses.npz"), seed=105, cd_dis=cd_ref.numpy(), adds=adds_ref.numpy(),
                        b=lb, n=ln, **{"s1_" + k: v.numpy() for k, v in ref_losses.items()},
                        **{"s2_" + k: v.numpy() for k, v in ref_losses_ref.items()})
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  ", f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
