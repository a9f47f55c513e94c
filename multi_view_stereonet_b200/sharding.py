"""Data-parallel plumbing for the path: image groups are independent end to end
(reference forward never mixes batch items; GroupNorm and the idepth range are per
sample, multi_view_stereonet.py:25-31, 144-158), so ranks own contiguous slices of the
batch and the only collective is one all_gather of per-rank timings."""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Contiguous slice [first, first + count) of `global_batch` items owned by `rank`."""
    assert global_batch % world == 0, "global batch must divide evenly over the ranks"
    per = global_batch // world
    return rank * per, per


def gather_timings(values, device=None):
    """all_gather of a small float64 vector; returns a (world, len) CPU tensor.
    Works without an initialised process group (world = 1)."""
    mine = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
        return torch.stack(out).cpu()
    return mine.cpu().unsqueeze(0)


def aggregate_throughput(units_per_rank_step, steps, per_rank_ms):
    """Whole-job units/s: all ranks' units over the slowest rank's time."""
    world = per_rank_ms.shape[0]
    worst_ms = float(per_rank_ms.max())
    return world * units_per_rank_step * steps / (worst_ms * 1e-3), worst_ms
