"""Rank / shard arithmetic for multi-GPU inference (one process per GPU, no data-path collective).

Keyframes are independent: rank r of W processes [lo, hi) of a global batch; the only
collectives are the barrier + MAX-reduction of the timed region and an optional gather of results to rank 0.
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def aggregate_throughput(units_this_rank, seconds_this_rank, device=None):
    """Whole-job throughput = units processed by all ranks / slowest rank's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units_this_rank / seconds_this_rank
    t = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item() / max_over_ranks(seconds_this_rank, device)


def gather_rows(local, total_rows, device=None):
    """All ranks contribute their (rows_r, ...) block; every rank gets the (total_rows, ...) concatenation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total_rows, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])
