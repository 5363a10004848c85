#!/usr/bin/env python
"""bench.py — headline benchmark of the DCL-Net hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 32] [--c_m 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): pose instances/s at N=M=1024.  One step = one stage-1 inference pass over a batch of
B=32 object instances per GPU (configs[2]: config_YCBV_bs32 shape), entered at the raw clouds (`--entry points`, default):
    device voxelisation -> two sparse-conv towers (tcgen05, output-stationary) ->
    pointnet_sp three_nn + three_interpolate (4 levels x 2 towers, fused) -> 8 disengage stacks ->
    dual fused FDA (tcgen05) -> confidence / fuser / regressor heads -> SVD pose projection.
`--entry pyramids` is round 1's entry (synthetic outputs of the towers as input; 30 MB of host input per step).
Random-init weights (no checkpoints offline), eval mode.  Operand precision `--precision fp16` (default: activations
rounded once to fp16, fp16 hi/lo weights, 2 MMAs per product; FDA logits on bf16 hi/lo operands) or `fp32-faithful`
(bf16 hi/lo everywhere, 3 MMAs); both are held to the same parity bars by the GPU tests.
Multi-GPU: instances are sharded, one process per GPU, B per GPU fixed (weak scaling); no data-path collective — the
poses are gathered once at the end of the timed region (`--gather step` gathers every step).
`--config stage2`: configs[3], stage 1 + the refiner loop, --batch-total instances sharded over the ranks (strong
scaling).  `--config train`: configs[4], one training step (fwd + bwd + Adam; DDP all-reduce for N > 1) at --batch
instances per GPU on the tensor-core training kernels (tools/train_step_ddp.py).

`value`  : device-resident throughput — inputs already in HBM, CUDA events, max over ranks; the timed step is one
           CUDA-graph replay, consecutive steps alternating over two streams.
`e2e`    : the same pass through PipelinedPoseEngine from pinned HOST buffers: H2D copies of that step's batch and
           the D2H read of the poses are inside the timed region.
`roofline`: the dominant kernel — the persistent CTA-pair tensor-core GEMM of the pointwise MLP stacks
           (pm_gemm_pair_kernel<256,8,fp16>, 5 launches per step) — timed live with CUDA events around each of its
           launches in an eager pass of the same K steps; algorithmic FLOPs = sum of 2*rows*cin*cout over the layers in
           those launches; peak = the measured cuBLAS bf16 BURST figure (the kernel is timed alone, the region is
           tens of milliseconds).  `roofline_fda`: the same for the fused FDA kernel, 2*N*M*(C + P + C) FLOPs per
           instance and direction.
`cpu_baseline`: the oracle port (pure PyTorch restatement of the same pass, oracle/torch_oracle.py) timed on this
           box's host cores on a bounded sample.  `--impl reference` prints that arm as its own line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pose instances/s (N=M=1024)"
UNIT = "instances/s"
N_PTS = 1024
P_DIM = 256
ROTATE = 3  # distinct input sets cycled through the timed steps
# DRAM bytes per launch of the dominant kernel in the fp16 configuration (ncu --set full, mean of its launches)
TRAFFIC_FP16, TRAFFIC_FP16_SRC = 1.01e8, "profiles/r02b_gemm_fda_ncu_full.txt (dram read+write, mean of the kernel's five launches)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="instances per GPU per step")
    ap.add_argument("--c_m", type=int, default=128, help="FDA similarity width: 128 = BASELINE.json, 64 = reference")
    ap.add_argument("--entry", default="points", choices=["points", "pyramids"],
                    help="where a step starts: points = raw clouds + colours (voxelisation and both sparse-conv towers run "
                         "on the device, ~1.5 MB of host input per step); pyramids = round 1's entry, synthetic outputs of "
                         "the towers (30 MB of host input per step)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32-faithful"],
                    help="operand format of the tensor-core path: fp16 = activations rounded once to fp16, fp16 hi/lo "
                         "weights (2 MMAs per product), split-operand FDA logits, fp16 P V (the product's default); "
                         "fp32-faithful = every operand a bf16 hi/lo pair (3 MMAs per product)")
    ap.add_argument("--cpu-batch", type=int, default=4, help="instances per step of the CPU arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--e2e-depth", type=int, default=3, help="batches in flight in the host-in/host-out pipeline")
    ap.add_argument("--streams", type=int, default=2,
                    help="CUDA streams the timed steps alternate over (each step is still one whole pass over its own "
                         "batch; with 2 the tail of one step's kernels overlaps the head of the next step's)")
    ap.add_argument("--train-entry", default="backbone", choices=["backbone", "feats"],
                    help="--config train: enter at the towers' pyramid levels (interpolation forward + backward included) "
                         "or at the (b*n, 480) point features")
    ap.add_argument("--train-eager", action="store_true",
                    help="--config train launches the step eagerly (torch DDP for N > 1).  Default: forward + backward "
                         "(+ Adam on one GPU) replayed as one CUDA graph; with N > 1 the gradients are averaged by one "
                         "NCCL all-reduce of a flat buffer after the replay")
    ap.add_argument("--train-layers", action="store_true",
                    help="--config train on the nn layer modules (cuDNN/cuBLAS) instead of the training kernels: A/B figure")
    ap.add_argument("--config", default="stage1", choices=["stage1", "stage2", "train"],
                    help="stage1 = BASELINE.json configs[2] (the headline); stage2 = configs[3]: stage 1 + 2 refiner "
                         "iterations over --batch-total instances sharded across the GPUs (strong scaling); "
                         "train = configs[4]: one DDP training step (fwd + bwd + Adam) at --batch instances per GPU")
    ap.add_argument("--batch-total", type=int, default=0,
                    help="stage2: instances per step over ALL GPUs (32..4096); each GPU runs its shard in passes of "
                         "--batch instances")
    ap.add_argument("--gather", default="end", choices=["end", "step"],
                    help="multi-GPU: all_gather of the poses once at the end of the timed region (every step's poses, "
                         "one collective) or after every step")
    ap.add_argument("--refine-iterations", type=int, default=0,
                    help="append the stage-2 refinement loop (tools/test_YCBV_stage2.py) with this many iterations; "
                         "0 = stage 1 only, the configuration BASELINE.json's metric is quoted on")
    args = ap.parse_args()
    if args.config == "stage2" and args.refine_iterations == 0:
        args.refine_iterations = 2          # scripts/script_eval_YCBV_stage2.sh:11
    if args.config == "train" and args.batch == 32:
        args.batch = 40                      # config_LM: bs40
    return args


def shard_plan(args, n_gpus):
    """(instances per GPU and step, instances per pass, passes per step).  stage1: one pass of --batch instances per
    GPU (weak scaling).  stage2 with --batch-total: the total is split evenly over the GPUs (strong scaling) and
    every GPU walks its shard in passes of at most --batch instances."""
    if args.config == "stage2" and args.batch_total > 0:
        if args.batch_total % n_gpus:
            raise SystemExit("--batch-total must be a multiple of the number of GPUs")
        per_gpu = args.batch_total // n_gpus
        chunk = min(args.batch, per_gpu)
        if per_gpu % chunk:
            raise SystemExit("--batch-total / gpus must be a multiple of --batch (or smaller than it)")
        return per_gpu, chunk, per_gpu // chunk
    return args.batch, args.batch, 1


def workload_config(args, n_gpus, cpu_arm=False):
    stage2 = (f" -> stage-2 refiner x{args.refine_iterations}" if getattr(args, "refine_iterations", 0) else "")
    per_gpu, chunk, passes = shard_plan(args, n_gpus)
    strong = args.config == "stage2" and args.batch_total > 0
    extra = {}
    if cpu_arm:
        # the CPU arm times a bounded sample: --cpu-batch instances per step, not the GPU arm's batch
        extra = {"B_per_step_cpu_arm": args.cpu_batch}
    if strong:
        extra.update({"B_total": args.batch_total, "passes_per_step": passes, "B_per_pass": chunk})
    return dict({
        "workload": ("DCL-Net stage-1 inference (config_YCBV_bs32 shape) from raw clouds + colours: device voxelisation "
                     "-> two sparse-conv towers -> " if getattr(args, "entry", "points") == "points" else
                     "DCL-Net stage-1 inference (config_YCBV_bs32 shape) from synthetic backbone pyramids: ") +
                    "pointnet_sp 3-NN interpolation -> disengage -> dual FDA -> heads -> SVD pose" + stage2,
        "entry": getattr(args, "entry", "points"),
        "B_per_gpu": per_gpu, "N": N_PTS, "M": N_PTS, "C": args.c_m, "P": P_DIM,
        "weights": "random init, eval mode", "precision": getattr(args, "precision", "fp16"),
        "sharding": f"instances x{n_gpus}, " + ("strong scaling (B_total fixed)" if strong else "weak scaling"),
        "l2": f"{ROTATE} rotating input sets; per-step activations (> 1 GB at B=32) exceed the 126 MB L2",
        "streams": f"{max(1, getattr(args, 'streams', 1))} CUDA stream(s): consecutive passes alternate over them, each pass one whole batch",
    }, **extra)


# ----------------------------------------------------------------------------------------------- inputs
def make_host_batch(seed, b, pin, entry="pyramids"):
    """entry="points": raw clouds + colours (what the dataloader hands to Network.forward); entry="pyramids": the
    clouds and a synthetic four-level voxel pyramid per tower standing in for the sparse-conv towers' outputs."""
    import torch
    from dcl_net_b200 import synthetic
    pts_inp = synthetic.object_clouds(seed, b, N_PTS, partial=True)   # observed: single-view half surface
    pts_tmp = synthetic.object_clouds(seed + 7919, b, N_PTS)            # template: closed surface
    batch = {"points_inp": pts_inp, "points_tmp": pts_tmp}
    if entry == "points":
        batch["rgb_inp"] = synthetic.point_colours(seed + 3, b * N_PTS)
        batch["rgb_tmp"] = synthetic.point_colours(seed + 4, b * N_PTS)
        if pin:
            batch = {k: v.pin_memory() for k, v in batch.items()}
        return batch
    batch.update({"inp": [(l.features, l.indices) for l in synthetic.backbone_levels(seed + 1, pts_inp, b)],
                  "tmp": [(l.features, l.indices) for l in synthetic.backbone_levels(seed + 2, pts_tmp, b)]})
    if pin:
        batch["points_inp"], batch["points_tmp"] = pts_inp.pin_memory(), pts_tmp.pin_memory()
        for side in ("inp", "tmp"):
            batch[side] = [(f.pin_memory(), i.pin_memory()) for f, i in batch[side]]
    return batch


class Cfg:
    n_inp = n_tmp = N_PTS
    unit_voxel_extent = [0.006] * 3


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING a timed region.  NVML (pynvml) at ~1 kHz when it is
    importable — a 20-step region lasts ~20 ms, too short for nvidia-smi's 100 ms polling — else `nvidia-smi -lms`."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self._stop, self._thread = None, None, threading.Event(), None
        self.sm, self.reason_bits, self.sm_max = [], 0, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(reasons(self.handle))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            nv, bits = self.nvml, self.reason_bits
            flags = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "samples": len(self.sm), "reasons": sorted(k for k, v in flags.items() if bits & v),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------------- CPU arm
def host_cores():
    """CPUs this process may run on (the GPU arm binds itself to its GPU's NUMA node before it gets here)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def cpu_pass_builder(args):
    """The oracle port of one step on the host: torch 3-NN interpolation + TailNetwork, all host threads."""
    import torch
    from oracle import torch_oracle as T
    torch.set_num_threads(host_cores())
    b = args.cpu_batch
    torch.manual_seed(0)
    net = T.TailNetwork(mode="test", c_m=args.c_m).eval()
    refiner = T.RefinerNet().eval() if args.refine_iterations > 0 else None
    batch = make_host_batch(1234, b, pin=False, entry=args.entry)
    ids = torch.arange(b).repeat_interleave(N_PTS)
    towers = None
    if args.entry == "points":
        from oracle import backbone_oracle as BO
        towers = [BO.BackboneOracle().eval(), BO.BackboneOracle().eval()]

    def pyramids():
        if towers is None:
            return batch["inp"], batch["tmp"]
        out = []
        for tw, side in zip(towers, ("inp", "tmp")):
            x, _ = BO.tower_input(batch["points_" + side], batch["rgb_" + side], b)
            out.append([(lv.features, lv.indices) for lv in tw(x)])
        return out

    def one_pass():
        with torch.no_grad():
            lv_inp, lv_tmp = pyramids()
            f_xc = T.get_point_feats(batch["points_inp"], ids, lv_inp, Cfg.unit_voxel_extent)
            f_yo = T.get_point_feats(batch["points_tmp"], ids, lv_tmp, Cfg.unit_voxel_extent)
            out = net(f_xc, f_yo, b, N_PTS, N_PTS)
            if refiner is not None:
                return T.stage2_refine(refiner, batch["points_inp"].view(b, N_PTS, 3), out["rot_pred"],
                                       out["trans_pred"], out["F_Xo_p"], out["conf"], args.refine_iterations)
        return out["rot_pred"], out["trans_pred"]
    return one_pass, b


def run_reference_arm(args, rank):
    """`--impl reference`: the reference has no CPU path for this workload (every op hard-codes CUDA and the CUDA
    extensions need THC), so the arm is the oracle port on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    one_pass, b = cpu_pass_builder(args)
    for _ in range(max(1, min(args.warmup, 2))):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    dt = time.perf_counter() - t0
    value = b * args.steps / dt
    sample = f"{args.steps} steps x {b} instances of the same workload on the host (oracle/torch_oracle.py)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong" if (args.config == "stage2" and args.batch_total) else "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, args.gpus, cpu_arm=True),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    one_pass, b = cpu_pass_builder(args)
    one_pass()
    t0, n = time.perf_counter(), 0
    while True:
        one_pass()
        n += 1
        dt = time.perf_counter() - t0
        if (dt >= 10.0 and n >= 3) or dt >= 30.0:
            break
    return {"value": b * n / dt, "unit": UNIT, "cores": host_cores(), "kind": "port",
            "sample": f"{n} passes x {b} instances (B={b} slice of the same workload), {dt:.1f} s, "
                      f"torch {host_cores()} threads, oracle/torch_oracle.py"}


# ----------------------------------------------------------------------------------------------- GPU arm
def pin_to_gpu_numa_node(index):
    """Bind this rank's threads to the CPUs NVML lists as local to its GPU, before any pinned host buffer is
    allocated: eight ranks feeding 30 MB per step each otherwise cross the socket interconnect for half of it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def run_b200_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dcl_net_b200 import _lib, fused_tail, modules, sharding
    from dcl_net_b200.dcl_net import Network
    from dcl_net_b200.engine import PipelinedPoseEngine, PoseEngine

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device: the product has no CPU fallback")
    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        pin_to_gpu_numa_node(local_rank)   # eight ranks sharing the host: keep each one's pinned buffers local
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    per_gpu, b, passes = shard_plan(args, world)      # b = instances per pass
    torch.manual_seed(0)
    from_points = args.entry == "points"
    net = Network(Cfg, mode="test", c_m=args.c_m, with_backbone=from_points).eval().to(dev)
    net.precision = args.precision
    fp16 = args.precision == "fp16"
    batches = [make_host_batch(1000 * (rank + 1) + 17 * i, b, pin=True, entry=args.entry) for i in range(ROTATE)]
    if from_points:
        from dcl_net_b200.backbone import SparseTowers
        caps = SparseTowers.plan_capacities([bt[k] for bt in batches for k in ("points_inp", "points_tmp")], b, N_PTS, dev)
    else:
        caps = [max(max(bt[s][lv][0].shape[0] for bt in batches for s in ("inp", "tmp")), 1) for lv in range(4)]
    refiner = None
    if args.refine_iterations > 0:
        from dcl_net_b200.refiner import Refiner
        refiner = Refiner().eval().to(dev)
        refiner.precision = args.precision
    nstreams = max(1, args.streams)
    n_eng = ROTATE if nstreams == 1 else nstreams * ((ROTATE + nstreams - 1) // nstreams)   # an engine stays on one stream
    engines = [PoseEngine(net, dev, b, caps, refiner, args.refine_iterations, entry=args.entry) for _ in range(n_eng)]
    for k, eng in enumerate(engines):
        eng.load(batches[k % ROTATE])
    torch.cuda.synchronize()
    side_streams = [torch.cuda.Stream(dev) for _ in range(nstreams)] if nstreams > 1 else []

    # Poses of every pass of the timed region, rank-local: (steps * passes, b, 12).  --gather end (default): ONE
    # all_gather of this buffer closes the timed region (each rank needs only its own slice until then);
    # --gather step: an all_gather after every pass, as in round 1.
    n_pass_total = args.steps * passes
    poses_local = torch.zeros(n_pass_total, b, 12, dtype=torch.float32, device=dev)
    poses_all = torch.empty(world * n_pass_total, b, 12, dtype=torch.float32, device=dev) if world > 1 else None

    def one_pass(i):
        eng = engines[i % n_eng]
        rot, trans = eng.run()
        if world > 1 and args.gather == "step":
            sharding.gather_poses(rot, trans, equal_shards=True)
        elif world > 1:
            poses_local[i % n_pass_total].copy_(eng.poses12)   # kept for the one all_gather that ends the region

    def step_resident(i, multi=False):
        """One step = `passes` passes of b instances (1 unless --batch-total shards a larger batch)."""
        for k in range(passes):
            j = i * passes + k
            if side_streams and multi:
                with torch.cuda.stream(side_streams[j % nstreams]):
                    one_pass(j)
            else:
                one_pass(j)

    def gather_end():
        if world > 1 and args.gather == "end":
            dist.all_gather_into_tensor(poses_all.view(-1), poses_local.view(-1))

    def fork():
        main = torch.cuda.current_stream(dev)
        for st in side_streams:
            st.wait_stream(main)

    def join():
        main = torch.cuda.current_stream(dev)
        for st in side_streams:
            main.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step_resident(i)
        # ---- timed region 0 (eager launches): per-launch CUDA events around the fused FDA kernel, and the
        #      count of this library's kernel launches per step
        modules.FDA_KERNEL_EVENTS = []
        fused_tail.GEMM_EVENTS = []
        if from_points:
            from dcl_net_b200 import backbone as backbone_mod
            backbone_mod.CONV_EVENTS = []
        launches0 = lib.dcl_b200_launch_count()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            step_resident(i)          # one stream: the per-kernel events below must not see another step's kernels
        gather_end()
        ev1.record()
        barrier()
        ms_eager = max_over_ranks(ev0.elapsed_time(ev1))
        launches = lib.dcl_b200_launch_count() - launches0
        fda_events, modules.FDA_KERNEL_EVENTS = modules.FDA_KERNEL_EVENTS, None
        fda_ms = [a.elapsed_time(bb) for a, bb, _ in fda_events]
        fda_jobs_per_launch = max([nj for _, _, nj in fda_events] or [1])
        conv_events = []
        if from_points:
            conv_events, backbone_mod.CONV_EVENTS = backbone_mod.CONV_EVENTS, None
        gemm_events, fused_tail.GEMM_EVENTS = fused_tail.GEMM_EVENTS, None
        gemm = [(a.elapsed_time(bb), fl) for a, bb, fl, nt in gemm_events if nt == 256]   # the 256-wide-tile kernel
        # ---- timed region 1: device-resident, the step replayed as a CUDA graph (same kernels, one launch)
        use_graph = not args.no_graph
        if use_graph:
            for eng in engines:
                eng.capture()
            for i in range(2 * n_eng):
                step_resident(i, multi=True)
            barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        fork()
        for i in range(args.steps):
            step_resident(i, multi=True)
        join()
        gather_end()
        ev1.record()
        barrier()
        ms_total = max_over_ranks(ev0.elapsed_time(ev1))
        clocks = sampler.stop()

        # ---- timed region 2: end to end through the public host-in / host-out API.  Every step copies ITS batch
        #      from pinned host memory and reads ITS (B,12) poses back; PipelinedPoseEngine overlaps the copy of
        #      batch i+1 with the pass over batch i (two buffer sets, a copy stream).
        del engines[1:]
        pipe = PipelinedPoseEngine(net, dev, b, caps, depth=max(args.e2e_depth, nstreams), refiner=refiner,
                                   iterations=args.refine_iterations, use_graph=use_graph,
                                   compute_streams=nstreams > 1, entry=args.entry)
        checksum = 0.0
        for rot, trans in pipe.infer_many(batches[i % ROTATE] for i in range(max(3, min(args.warmup, 5)))):
            checksum += float(trans[0, 0])
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rot, trans in pipe.infer_many(batches[i % ROTATE] for i in range(args.steps * passes)):
            checksum += float(trans[0, 0]) + float(rot[-1, 2, 2])    # the host consumes every pass's result
        e1.record()
        barrier()
        if from_points:
            for eng in engines + pipe.engines:
                eng.towers.check_errors()      # no voxel set outgrew its buffer, no point fell off the grid
        h2d = pipe.h2d_bytes
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))

    value = world * per_gpu * args.steps / (ms_total / 1e3)
    e2e_value = world * per_gpu * args.steps / (e2e_ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    # The kernels are timed alone (CUDA events around single launches in a region of tens of milliseconds, GPU far
    # from its power limit): the applicable peak is the measured BURST cuBLAS figure, not the sustained one.
    peak_sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops_burst") or peaks.get("bf16_tflops") or 1600.0
    peak_tf = peak_burst
    peak_src = ("MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks
                else "B200_PROFILING.md fallback (of fallback)")
    # Dominant kernel: the persistent CTA-pair GEMM with 256-wide tiles (disengage and fuser layers; 5 launches/step).
    gemm_ms, gemm_flops = sum(t for t, _ in gemm), sum(fl for _, fl in gemm)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm else float("nan")
    mmas = 2 if fp16 else 3
    split_note = ("fp16 activations rounded once x fp16 hi/lo weights: 2 MMAs per product, so the algorithmic fraction is "
                  "bounded by 1/2; tensor-pipe occupancy is ~2x the algorithmic fraction" if fp16 else
                  "every product runs as 3 bf16 MMAs (hi/lo operand split) to stay fp32-faithful, so the algorithmic "
                  "fraction is bounded by 1/3; tensor-pipe occupancy is ~3x the algorithmic fraction")
    # DRAM bytes per launch of this kernel (dram__bytes_read.sum + dram__bytes_write.sum, mean of its five launches)
    # from the `ncu --set full` capture of the default configuration: profiles/r01_gemm_fda_ncu_full.txt (taken on
    # the multicast variant of the kernel; operands and outputs, hence the DRAM bytes, are the same)
    traffic, traffic_src = None, None
    if b == 32 and args.c_m == 128:
        traffic, traffic_src = ((TRAFFIC_FP16, TRAFFIC_FP16_SRC) if fp16 else
                                (2.31e8, "profiles/r01_gemm_fda_ncu_full.txt (bytes per launch)"))
    gemm_kernel = ("pm_gemm_pair_kernel<256,8,fp16>" if fp16 else
                   "pm_gemm_cluster_kernel<256,4>" if os.environ.get("DCL_PM_GEMM_MCAST")
                   else "pm_gemm_kernel<256,2>" if os.environ.get("DCL_PM_GEMM_SIMPLE") else "pm_gemm_pair_kernel<256,6>")
    roofline = {"kernel": gemm_kernel, "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "avg_launch_ms": gemm_ms / len(gemm) if gemm else None, "launches_timed": len(gemm),
                "algorithmic_flops_per_step": gemm_flops / args.steps,
                "mmas_per_product": mmas,
                "executed_mma_flops_per_step": mmas * gemm_flops / args.steps,
                "executed_frac": mmas * achieved / peak_tf,
                "frac_of_sustained_peak": achieved / peak_sustained,
                "share_of_step": gemm_ms / ms_eager if gemm else None,
                "timed_in": "eager pass of the same K steps (the graph-replayed pass launches the identical kernels)",
                "note": split_note}
    flops_per_launch = fda_jobs_per_launch * b * 2.0 * N_PTS * N_PTS * (args.c_m + P_DIM + args.c_m)
    fda_avg_ms = statistics.mean(fda_ms) if fda_ms else float("nan")
    fda_achieved = flops_per_launch / (fda_avg_ms * 1e-3) / 1e12
    # executed MMA work per algorithmic FLOP: logits (share C / (2C + P)) on split operands, P V on fp16 or split ones
    qk_share = args.c_m / (2.0 * args.c_m + P_DIM)
    fda_exec = (3 * qk_share + 1 * (1 - qk_share)) if fp16 else 3.0
    fda_kernel = ("fda_pair_kernel" if (N_PTS // 128) % 2 == 0 and not os.environ.get("DCL_FDA_SINGLE")
                  else "fda_fwd_kernel")
    roofline_fda = {"kernel": f"{fda_kernel}<{args.c_m}>", "bound": "tensor", "achieved": fda_achieved,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": fda_achieved / peak_tf,
                    "executed_frac": fda_exec * fda_achieved / peak_tf, "avg_launch_ms": fda_avg_ms,
                    "mmas_per_product": "logits 3, P V 1" if fp16 else "3",
                    "launches_timed": len(fda_ms), "algorithmic_flops_per_launch": flops_per_launch,
                    "directions_per_launch": fda_jobs_per_launch,
                    "share_of_step": (sum(fda_ms) / ms_eager) if fda_ms else None}
    roofline_spconv = None
    if conv_events:
        # the 16 sparse convolutions of both towers (8 launches per step, one per layer for the pair of towers):
        # algorithmic FLOPs = 2 * (output row, kernel offset) pairs with an input voxel * c_in * c_out, pairs counted from
        # the rulebooks of the timed inputs; fp16 activations x fp16 hi/lo weights = 2 MMAs per product
        tw = engines[0].towers
        pairs = {op: tw.rulebook_pairs(op) for op in sorted({op for _, _, op, _, _ in conv_events})}
        conv_ms = [a.elapsed_time(bb) for a, bb, _, _, _ in conv_events]
        conv_flops = sum(2.0 * pairs[op] * ci * co for _, _, op, ci, co in conv_events)
        gather_bytes = sum(2.0 * pairs[op] * ci for _, _, op, ci, _ in conv_events)     # fp16 input rows gathered
        conv_tf = conv_flops / (sum(conv_ms) * 1e-3) / 1e12
        roofline_spconv = {"kernel": "sparse_conv3_kernel (8 launches per step: both towers per layer)", "bound": "tensor",
                           "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf,
                           "executed_frac": 2 * conv_tf / peak_tf, "mmas_per_product": 2,
                           "algorithmic_flops_per_step": conv_flops / args.steps,
                           "rulebook_pairs_per_step": sum(pairs[op] for _, _, op, _, _ in conv_events) / args.steps,
                           "rulebook_pairs_source": "rulebooks of the last timed input set (the sets rotate; synthetic clouds "
                                                    "of equal statistics)",
                           "gathered_input_GBps": gather_bytes / (sum(conv_ms) * 1e-3) / 1e9,
                           "ms_per_step": sum(conv_ms) / args.steps, "launches_timed": len(conv_ms),
                           "share_of_step": sum(conv_ms) / ms_eager,
                           "note": "the A operand is gathered row by row (16-byte cp.async pieces) from L2: the gather, "
                                   "not the tensor pipe, paces these kernels (profiles/r02_trace_spconv_8warps.txt)"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if (args.config == "stage2" and args.batch_total) else "weak", "vs_baseline": None,
                "dtype": ("fp16 activations x fp16 hi/lo weights, fp32 accumulate (FDA logits: bf16 hi/lo split operands)"
                          if fp16 else "fp32 (tensor-core contractions: bf16 hi/lo split operands, fp32 accumulate)"),
                "data": "synthetic", "config": workload_config(args, world), "impl": "b200",
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * passes,
                        "d2h_bytes_per_step": per_gpu * 12 * 4,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "launch_mode": "cuda_graph" if use_graph else "eager",
                "streams": nstreams, "pose_gather": args.gather if world > 1 else None,
                "ms_per_step_eager": ms_eager / args.steps, "clocks": clocks, "roofline": roofline,
                "roofline_fda": roofline_fda}
        if roofline_spconv is not None:
            line["roofline_spconv"] = roofline_spconv
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.config == "train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_step_ddp
        train_step_ddp.run(args.batch, args.steps, max(args.warmup, 3), rank, world, local_rank, contract=True,
                           layers=args.train_layers, entry=args.train_entry,
                           graph=not (args.train_eager or args.train_layers))
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see the module docstring)")
    run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
