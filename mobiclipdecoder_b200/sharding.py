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


def cores_of_rank(local_rank, local_world, available=None):
    """The host cores rank `local_rank` of `local_world` parses on: an equal, contiguous slice of the cores this process may
    run on.  Contiguous, because Linux numbers the cores of one socket consecutively and GPUs k and k+1 hang off the same
    socket on the 8-GPU boards: the parse threads, the pinned arenas they fill (first touch) and the GPU's PCIe root then
    sit on one NUMA node."""
    cores = sorted(available if available is not None else os.sched_getaffinity(0))
    per = max(1, len(cores) // max(1, local_world))
    lo = (local_rank * per) % len(cores)
    return cores[lo:lo + per] or cores


def pin_to_cores(local_rank, local_world):
    """Restrict this process (and every thread it starts from now on) to its slice.  Returns the slice."""
    mine = cores_of_rank(local_rank, local_world)
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        pass
    return mine


def gather_objects(dist, obj):
    """Every rank's `obj`, in rank order (a list of one when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
