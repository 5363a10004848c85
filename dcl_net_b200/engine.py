"""PoseEngine — the call a user of this path makes: a batch of instances on the HOST in, poses on the host out.

    engine = PoseEngine(net, device, batch, refiner=None, iterations=0)
    rot, trans = engine.infer(host_batch)

Two entries (pinned host tensors):
  entry="points" (a Network built with its towers): the raw clouds and colours,
    {"points_inp": (B*N,3), "rgb_inp": (B*N,3), "points_tmp": (B*M,3), "rgb_tmp": (B*M,3)}      ~1.5 MB per 32 instances
    — voxelisation, both sparse-conv towers, interpolation, FDA and pose all run on the device;
  entry="pyramids": what an external backbone provider would hand over,
    {"points_inp": (B*N,3), "points_tmp": (B*M,3),
     "inp": [(features (Mv,C_l), indices (Mv,4) int32 bxyz) x 4 levels], "tmp": [... x 4]}       ~30 MB per 32 instances
Device buffers are allocated once with fixed row capacities; every call copies the batch into them (cudaMemcpyAsync
from pinned memory); unused pyramid rows carry batch id == B, a bucket no query belongs to, so shapes stay static.
Multi-GPU: one engine per process / GPU over its own instance shard (sharding.py).
"""
import types

import torch

from . import _lib as L
from .refiner import refine_poses


class PoseEngine:
    def __init__(self, net, device, batch, capacities, refiner=None, iterations=0, entry="pyramids"):
        """capacities: entry="pyramids": per level, the maximum number of voxel rows of a batch (both towers use the
        same); entry="points": the nine voxel-set capacities of backbone.SparseTowers (plan_capacities)."""
        if net.training or (refiner is not None and refiner.training):
            raise ValueError("PoseEngine is an inference engine: put the network (and the refiner) in eval() mode first")
        self.net, self.refiner, self.iterations = net, refiner, iterations
        self.device, self.b = device, batch
        self.n_inp, self.n_tmp = net.n_inp, net.n_tmp
        L.load()
        f32, i32 = dict(dtype=torch.float32, device=device), dict(dtype=torch.int32, device=device)
        self.entry = entry
        self.points = {"inp": torch.empty(batch * self.n_inp, 3, **f32), "tmp": torch.empty(batch * self.n_tmp, 3, **f32)}
        self.levels, self.rgb, self.towers = {}, {}, None
        if entry == "points":
            self.rgb = {"inp": torch.empty(batch * self.n_inp, 3, **f32), "tmp": torch.empty(batch * self.n_tmp, 3, **f32)}
            from .backbone import SparseTowers
            # every engine owns its towers' buffers (engines run concurrently on different streams)
            self.towers = SparseTowers(net.backbone_inp, net.backbone_tmp, device, batch, self.n_inp, capacities,
                                       unit=float(net.unit_voxel_extent[0]))
        elif entry == "pyramids":
            for side in ("inp", "tmp"):
                self.levels[side] = [types.SimpleNamespace(features=torch.zeros(cap, ch, **f32),
                                                           indices=torch.zeros(cap, 4, **i32))
                                     for cap, ch in zip(capacities, (32, 64, 128, 256))]
        else:
            raise ValueError("entry must be 'points' or 'pyramids'")
        self.out_host = torch.empty(batch, 12, dtype=torch.float32).pin_memory()
        self.h2d_bytes = 0
        self._graph = None
        self._static = None
        self.poses12 = None
        self._packed = (None, None)   # the packed-weight objects the captured graph holds raw pointers into

    def _current_packed(self):
        return (getattr(self.net, "_fused_tail", None), getattr(self.refiner, "_fused_refiner", None))

    def load(self, host_batch):
        """Asynchronous host->device copy of one batch into the static buffers."""
        nbytes = 0
        for side in ("inp", "tmp"):
            src = host_batch["points_" + side]
            self.points[side].copy_(src, non_blocking=True)
            nbytes += src.numel() * 4
            if self.entry == "points":
                rgb = host_batch["rgb_" + side]
                self.rgb[side].copy_(rgb, non_blocking=True)
                nbytes += rgb.numel() * 4
                continue
            for lvl, (feats, ind) in zip(self.levels[side], host_batch[side]):
                m = feats.shape[0]
                if m > lvl.features.shape[0]:
                    raise ValueError("PoseEngine: batch exceeds the level capacity it was built with")
                lvl.features[:m].copy_(feats, non_blocking=True)
                lvl.indices[:m].copy_(ind, non_blocking=True)
                lvl.indices[m:, 0] = self.b  # padding rows: a batch id no query has
                nbytes += feats.numel() * 4 + ind.numel() * 4
        self.h2d_bytes = nbytes

    def capture(self, warmup=2):
        """Capture the whole pass into a CUDA graph (shapes are static): one launch per step instead of a few
        hundred.  The loaded batch must be valid; later load() calls just refill the same buffers."""
        if self.towers is not None:
            self.towers.repack()       # captured pointers must be those of the current tower weights
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run_eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._static = self._run_eager()
        self._graph = graph
        # The graph replays raw pointers into the packed weights of net._fused_tail / refiner._fused_refiner, which
        # live outside the graph's private pool: keep them alive here, and re-capture (run()) if the modules
        # rebuilt them since (load_state_dict, .to(), a train()/eval() round trip).
        self._packed = self._current_packed()
        return self

    def run(self):
        """Device-resident pass over the loaded batch -> (rot (B,3,3), trans (B,3)) on the device; the same poses
        packed as (B,12) rows [R row-major | t] are left in self.poses12 (what goes back to the host)."""
        if self._graph is not None:
            if any(cur is not old for cur, old in zip(self._current_packed(), self._packed)):
                self._graph = None
                self.capture()             # weights were re-packed: the old graph points at the previous copies
            self._graph.replay()
            rot, trans, self.poses12 = self._static
            return rot, trans
        rot, trans, self.poses12 = self._run_eager()
        return rot, trans

    @torch.no_grad()
    def _run_eager(self):
        if self.entry == "points":
            levels_inp, levels_tmp = self.towers.run(self.points["inp"], self.rgb["inp"], self.points["tmp"], self.rgb["tmp"])
        else:
            levels_inp, levels_tmp = self.levels["inp"], self.levels["tmp"]
        pred = self.net.forward_from_backbone(levels_inp, levels_tmp, self.points["inp"], self.points["tmp"], self.b)
        rot, trans = pred["rot_pred"], pred["trans_pred"]
        if self.refiner is not None and self.iterations > 0:
            rot, trans = refine_poses(self.refiner, self.points["inp"].view(self.b, self.n_inp, 3), rot, trans,
                                      pred["F_Xo_p"], pred["conf"], self.iterations, pred.get("F_Xo_p_pm"),
                                      pred.get("F_Xo_p_pm_fmt", 0))
        return rot, trans, torch.cat([rot.reshape(self.b, 9), trans], dim=1)

    def infer(self, host_batch):
        """Host in, host out: H2D copies + pass + D2H of the (B,12) poses; returns after the poses have landed."""
        self.load(host_batch)
        self.run()
        self.out_host.copy_(self.poses12, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.out_host[:, :9].view(self.b, 3, 3), self.out_host[:, 9:]


class PipelinedPoseEngine:
    """Host-in / host-out inference over a stream of batches with the copies overlapped with compute:
    `depth` PoseEngines (each with its own static buffers and CUDA graph) are used round-robin; batch i+1 is
    copied to the device on a copy stream while batch i computes, and each batch's (B,12) poses are read back
    to pinned host memory right after its pass.  Every batch still pays its own H2D and D2H."""

    def __init__(self, net, device, batch, capacities, depth=2, refiner=None, iterations=0, use_graph=True,
                 compute_streams=False, entry="pyramids"):
        self.engines = [PoseEngine(net, device, batch, capacities, refiner, iterations, entry) for _ in range(depth)]
        self.device, self.use_graph = device, use_graph
        self.copy_stream = torch.cuda.Stream(device)
        # compute_streams: every engine slot runs on a stream of its own, so the tail of one pass's kernels
        # overlaps the head of the next pass (each pass still owns its buffers and its graph)
        self.compute_streams = [torch.cuda.Stream(device) for _ in range(depth)] if compute_streams else None
        self._captured = False
        self.h2d_bytes = 0

    def _ensure_graphs(self, first_batch):
        if self.use_graph and not self._captured:
            for eng in self.engines:
                eng.load(first_batch)
                torch.cuda.synchronize(self.device)
                eng.capture()
        self._captured = True

    def infer_many(self, host_batches):
        """Yields (rot (B,3,3), trans (B,3)) host tensors, in order, one per input batch.  The yielded tensors
        alias a per-engine pinned buffer: consume them before `depth` more batches have been requested."""
        host_batches = iter(host_batches)
        try:
            first = next(host_batches)
        except StopIteration:
            return
        self._ensure_graphs(first)
        main = torch.cuda.current_stream(self.device)
        depth = len(self.engines)
        loaded = [torch.cuda.Event() for _ in range(depth)]
        done = [torch.cuda.Event() for _ in range(depth)]
        pending = []   # engine slots whose results have not been yielded yet

        def stage(slot, batch):
            eng = self.engines[slot]
            self.copy_stream.wait_event(done[slot])          # the slot's previous pass has consumed its buffers
            with torch.cuda.stream(self.copy_stream):
                eng.load(batch)
                loaded[slot].record(self.copy_stream)
            self.h2d_bytes = eng.h2d_bytes

        for ev in done:
            ev.record(main)
        for cs in self.compute_streams or []:
            cs.wait_stream(main)
        stage(0, first)
        i = 0
        nxt = next(host_batches, None)
        while True:
            slot = i % depth
            eng = self.engines[slot]
            if nxt is not None:
                stage((i + 1) % depth, nxt)                   # overlaps with the pass launched below
            cs = self.compute_streams[slot] if self.compute_streams else main
            cs.wait_event(loaded[slot])
            with torch.cuda.stream(cs):
                eng.run()
                eng.out_host.copy_(eng.poses12, non_blocking=True)
                done[slot].record(cs)
            pending.append(slot)
            if len(pending) == depth or nxt is None:
                while pending and (len(pending) == depth or nxt is None):
                    s = pending.pop(0)
                    done[s].synchronize()
                    e = self.engines[s]
                    yield e.out_host[:, :9].view(e.b, 3, 3), e.out_host[:, 9:]
            if nxt is None:
                break
            i += 1
            nxt = next(host_batches, None)
