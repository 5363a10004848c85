"""PoseEngine / PipelinedPoseEngine: host-in -> host-out inference, CUDA-graph replay, and passes overlapped on two
compute streams must all give the poses of plain sequential eager passes, bit for bit."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                        # noqa: E402  (synthetic batch builder)
from dcl_net_b200.dcl_net import Network                            # noqa: E402
from dcl_net_b200.engine import PipelinedPoseEngine, PoseEngine     # noqa: E402
from dcl_net_b200.refiner import Refiner                            # noqa: E402


@pytest.mark.parametrize("iterations", [0, 2])
def test_pipelined_two_streams_equals_sequential(cuda_dev, iterations):
    b = 4
    torch.manual_seed(3)
    net = Network(bench.Cfg, mode="test", c_m=128).eval().to(cuda_dev)
    refiner = Refiner().eval().to(cuda_dev) if iterations else None
    batches = [bench.make_host_batch(50 + i, b, pin=True) for i in range(5)]
    caps = [max(max(bt[s][lv][0].shape[0] for bt in batches for s in ("inp", "tmp")), 1) for lv in range(4)]
    with torch.no_grad():
        seq = PoseEngine(net, cuda_dev, b, caps, refiner, iterations)
        want = []
        for bt in batches:
            rot, trans = seq.infer(bt)
            want.append((rot.clone(), trans.clone()))
        for streams in (False, True):
            pipe = PipelinedPoseEngine(net, cuda_dev, b, caps, depth=2, refiner=refiner, iterations=iterations,
                                       use_graph=True, compute_streams=streams)
            got = [(r.clone(), t.clone()) for r, t in pipe.infer_many(batches[i % 5] for i in range(15))]
            assert len(got) == 15
            for i, (r, t) in enumerate(got):
                assert torch.equal(r, want[i % 5][0]) and torch.equal(t, want[i % 5][1]), (streams, i)
    # different batches do give different poses (the comparison above is not vacuous)
    assert not torch.equal(want[0][0], want[1][0])


def _oracle_poses(net_state, c_m, batch, b, dev, refiner_state=None, iterations=0):
    """The restated reference graph on the same host batch: C-oracle 3-NN (bit-faithful tie-break) + torch fp32
    interpolation / FDA section / SVD on the GPU with TF32 off (+ the restated stage-2 loop)."""
    from oracle import cpu_oracle, torch_oracle as T
    oracle_net = T.TailNetwork(mode="test", c_m=c_m).eval()
    missing = oracle_net.load_state_dict(net_state, strict=False)
    assert not missing.missing_keys
    c_nn = lambda u, k: tuple(map(torch.from_numpy, cpu_oracle.sp_three_nn(u.numpy(), k.numpy())))
    ids = torch.arange(b).repeat_interleave(bench.N_PTS)
    f_xc = T.get_point_feats(batch["points_inp"], ids, batch["inp"], bench.Cfg.unit_voxel_extent, three_nn=c_nn)
    f_yo = T.get_point_feats(batch["points_tmp"], ids, batch["tmp"], bench.Cfg.unit_voxel_extent, three_nn=c_nn)
    with torch.no_grad():
        out = oracle_net.to(dev)(f_xc.to(dev), f_yo.to(dev), b, bench.N_PTS, bench.N_PTS)
        rot, trans = out["rot_pred"], out["trans_pred"]
        if iterations:
            oref = T.RefinerNet().eval()
            oref.load_state_dict(refiner_state)
            rot, trans = T.stage2_refine(oref.to(dev), batch["points_inp"].view(b, bench.N_PTS, 3).to(dev), rot, trans,
                                         out["F_Xo_p"], out["conf"], iterations)
    return rot.cpu(), trans.cpu()


@pytest.mark.parametrize("precision", ["fp16", "fp32-faithful"])
@pytest.mark.parametrize("iterations", [0, 2])
def test_headline_config_engine_vs_oracle(cuda_dev, iterations, precision):
    """The exact configuration bench.py times — B=32, N=M=1024, C=128, entered at the backbone pyramids, through
    PoseEngine.infer (static buffers, CUDA graph) — against the oracle: 0.01 deg / 1e-5 m (north_star).
    iterations=2: the same with the stage-2 refinement loop appended (BASELINE.json configs[3])."""
    from oracle import torch_oracle as T
    b = 32
    torch.manual_seed(0)
    net = Network(bench.Cfg, mode="test", c_m=128).eval().to(cuda_dev)
    net.precision = precision
    refiner = Refiner().eval().to(cuda_dev) if iterations else None
    batch = bench.make_host_batch(1017, b, pin=True)
    caps = [max(batch[s][lv][0].shape[0] for s in ("inp", "tmp")) for lv in range(4)]
    eng = PoseEngine(net, cuda_dev, b, caps, refiner, iterations)
    eng.load(batch)
    torch.cuda.synchronize()
    eng.capture()
    rot, trans = eng.infer(batch)
    assert eng._graph is not None
    assert net._fused_tail.fmt == (1 if precision == "fp16" else 0)
    state = {k: v.cpu() for k, v in net.state_dict().items()}
    rstate = {k: v.cpu() for k, v in refiner.state_dict().items()} if iterations else None
    want_rot, want_trans = _oracle_poses(state, 128, batch, b, cuda_dev, rstate, iterations)
    ang = T.rotation_angle_deg(rot.clone(), want_rot).max().item()
    dt = (trans.clone() - want_trans).abs().max().item()
    assert ang < 0.01 and dt < 1e-5, (ang, dt)


@pytest.mark.parametrize("c_m", [64, 128])
def test_engine_padded_capacity_vs_oracle(cuda_dev, c_m):
    """Level buffers larger than the batch's row counts: the unused rows carry batch id == B (a bucket no query
    belongs to) and must not change any pose; the second batch is smaller than the first, so stale rows of the
    previous batch sit behind the live ones as well."""
    from oracle import torch_oracle as T
    b = 4
    torch.manual_seed(5)
    net = Network(bench.Cfg, mode="test", c_m=c_m).eval().to(cuda_dev)
    big, small = bench.make_host_batch(77, b, pin=True), bench.make_host_batch(78, b, pin=True)
    for side in ("inp", "tmp"):        # drop a third of the second batch's coarsest-level rows as well
        f, i = small[side][3]
        keep = f.shape[0] - f.shape[0] // 3
        small[side][3] = (f[:keep].clone().pin_memory(), i[:keep].clone().pin_memory())
    caps = [int(1.5 * max(bt[s][lv][0].shape[0] for bt in (big, small) for s in ("inp", "tmp"))) + 7 for lv in range(4)]
    eng = PoseEngine(net, cuda_dev, b, caps)
    eng.load(big)
    torch.cuda.synchronize()
    eng.capture()
    state = {k: v.cpu() for k, v in net.state_dict().items()}
    for batch in (big, small):
        assert all(batch[s][lv][0].shape[0] < caps[lv] for s in ("inp", "tmp") for lv in range(4))
        rot, trans = eng.infer(batch)
        want_rot, want_trans = _oracle_poses(state, c_m, batch, b, cuda_dev)
        ang = T.rotation_angle_deg(rot.clone(), want_rot).max().item()
        dt = (trans.clone() - want_trans).abs().max().item()
        assert ang < 0.01 and dt < 1e-5, (ang, dt)


def test_engine_recaptures_when_weights_are_repacked(cuda_dev):
    """A captured graph holds raw pointers into the packed weights: after load_state_dict (which re-packs them) the
    engine must not replay the stale graph; a second engine on the same network must not invalidate the first."""
    b = 4
    torch.manual_seed(8)
    net = Network(bench.Cfg, mode="test", c_m=64).eval().to(cuda_dev)
    batch = bench.make_host_batch(91, b, pin=True)
    caps = [max(batch[s][lv][0].shape[0] for s in ("inp", "tmp")) for lv in range(4)]
    eng = PoseEngine(net, cuda_dev, b, caps)
    eng.load(batch)
    torch.cuda.synchronize()
    eng.capture()
    rot0, trans0 = (t.clone() for t in eng.infer(batch))
    eng2 = PoseEngine(net, cuda_dev, b, caps)          # shares net: must leave eng's packed weights alone
    eng2.load(batch)
    torch.cuda.synchronize()
    eng2.capture()
    rot0b, trans0b = (t.clone() for t in eng.infer(batch))
    assert torch.equal(rot0, rot0b) and torch.equal(trans0, trans0b)
    torch.manual_seed(9)
    other = Network(bench.Cfg, mode="test", c_m=64).eval()
    net.load_state_dict(other.state_dict())
    rot1, trans1 = (t.clone() for t in eng.infer(batch))
    fresh = PoseEngine(net, cuda_dev, b, caps)
    rot2, trans2 = (t.clone() for t in fresh.infer(batch))
    assert torch.equal(rot1, rot2) and torch.equal(trans1, trans2)
    assert not torch.equal(rot0, rot1)
