"""Instance sharding across the GPUs of one box.

Every op on the path is per object instance, so inference shards by contiguous instance ranges,
one process per GPU, weights replicated, and needs no data-path collective; the only exchange is
an optional all_gather of the (B/G, 12) poses.  The flat pointnet_sp tensors carry a batch id in
column 0, which is re-based to the rank-local range.  Training is data parallel: every rank steps on its own
instances and the gradients are averaged by ONE all-reduce of a flat buffer (flat_grad_buffer / average_gradients) —
what DDP's bucketed all-reduce computes, without per-bucket host work, so that the step itself can be a CUDA graph.
"""
import torch
import torch.distributed as dist


def instance_range(total, rank, world):
    """Contiguous [lo, hi) of `total` instances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_flat(rows, lo, hi):
    """Rows of a flat (n, 1+k) tensor whose column 0 is a batch id in [lo, hi), re-based to start at 0."""
    keep = (rows[:, 0] >= lo) & (rows[:, 0] < hi)
    out = rows[keep].clone()
    out[:, 0] -= lo
    return out, keep


def gather_poses(rot, trans, total=None, equal_shards=False):
    """all_gather of per-rank (b_r,3,3)/(b_r,3) poses into (B,3,3)/(B,3) on every rank.  Works with
    ragged shards (the shard sizes are exchanged first, which costs a host sync); pass equal_shards=True when
    every rank holds the same number of instances to get a single asynchronous all_gather.
    No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rot, trans
    world = dist.get_world_size()
    packed = torch.cat([rot.reshape(rot.shape[0], 9), trans], dim=1).contiguous()
    if equal_shards:
        full = packed.new_empty(world * packed.shape[0], 12)
        dist.all_gather_into_tensor(full, packed)
        return full[:, :9].reshape(-1, 3, 3), full[:, 9:]
    counts = [torch.zeros(1, dtype=torch.int64, device=packed.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device))
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    padded = packed.new_zeros(width, 12)
    padded[: packed.shape[0]] = packed
    bufs = [packed.new_zeros(width, 12) for _ in range(world)]
    dist.all_gather(bufs, padded)
    full = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return full[:, :9].reshape(-1, 3, 3), full[:, 9:]


def flat_grad_buffer(params):
    """One flat fp32 buffer holding every parameter's gradient: sets p.grad to views of it (autograd then accumulates
    in place, at addresses that never change — what a captured CUDA graph needs) and returns the buffer.  Zero it at
    the start of every step instead of calling zero_grad(set_to_none=True)."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def average_gradients(flat):
    """In-place mean of the flat gradient buffer over the ranks: one all-reduce (NCCL: ReduceOp.AVG; backends without
    AVG, e.g. gloo in the CPU tests: SUM then divide).  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat
