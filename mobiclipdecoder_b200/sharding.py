"""Multi-GPU plumbing for independent streams: stream i belongs to rank i % world; there is no
data-path collective.  torch.distributed is used only to line the ranks up (barrier) and to take the slowest
rank's time, so the same code runs over NCCL on the GPU box and over gloo in the CPU tests."""
import os


def world_from_env():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def streams_of_rank(n_total, rank, world):
    """Global stream ids owned by `rank` (round-robin, MD: one decoder object per stream, no shared state)."""
    return list(range(rank, n_total, world))


def stream_seed(base_seed, global_stream_id):
    return base_seed + global_stream_id


def barrier(dist, device=None):
    if dist is not None and dist.is_initialized():
        if device is not None:
            dist.barrier(device_ids=[device])
        else:
            dist.barrier()


def max_over_ranks(dist, value, torch, device='cpu'):
    """Slowest rank's elapsed time: the job is done when the last rank is."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value, torch, device='cpu'):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_fps(frames_per_rank_total, max_ms):
    return frames_per_rank_total / (max_ms * 1e-3)
