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
