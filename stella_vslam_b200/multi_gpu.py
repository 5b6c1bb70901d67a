"""Multi-GPU plumbing of the hot path: streams (and independent BA windows) shard embarrassingly,
one process per GPU; the only collective is an all-gather of small fixed-size per-stream records so that every rank (and
rank 0's report) sees all results.  torch.distributed is used for the plumbing: NCCL over NVLink on the GPU box, gloo in
the CPU tests."""
import torch
import torch.distributed as dist


def assign_streams(n_streams, world_size, rank):
    """Static partition: stream s is owned by rank s % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return [s for s in range(n_streams) if s % world_size == rank]


def owner_of(stream, world_size):
    return stream % world_size


def gather_records(local, world_size, out=None):
    """All-gather a (n_local, k) int32 tensor of per-frame/stream records -> (world_size, n_local, k) on every rank.
    Every rank must pass the same shape (the per-GPU work is fixed: weak scaling)."""
    if world_size == 1:
        return local.unsqueeze(0)
    if out is None:
        out = torch.empty((world_size,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous().unsqueeze(0))
    return out


def max_over_ranks(value, device, world_size):
    """Device-side timing reduction: the step time of the job is the slowest rank's."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def frames_per_second(frames_per_rank_per_step, steps, world_size, ms_total_max):
    """Whole-job aggregate: all ranks' frames over the slowest rank's time."""
    return world_size * frames_per_rank_per_step * steps / (ms_total_max * 1e-3)
