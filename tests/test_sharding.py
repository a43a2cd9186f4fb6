"""N > 1 host-side logic on CPU: world_size-2 gloo processes shard independent streams (no data-path collective),
parse their own shard with the native parser, and agree on the max-over-ranks time and the aggregate count."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mobiclipdecoder_b200 import sharding


def test_round_robin_partition_is_exact():
    for n in (1, 7, 8, 1024):
        for world in (1, 2, 4, 8):
            owned = [sharding.streams_of_rank(n, r, world) for r in range(world)]
            flat = sorted(x for o in owned for x in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mobiclipdecoder_b200 import MobiParser
    from mobiclipdecoder_b200.workloads import CONFIGS, make_stream
    w, h, ver, _ = CONFIGS['moflex_400x240']
    mine = sharding.streams_of_rank(6, rank, world)
    frames = 0
    for g in mine:  # each rank decodes (here: parses) only its own streams; nothing crosses ranks
        s, p = make_stream('moflex_400x240', sharding.stream_seed(50, g)), MobiParser(w, h, ver)
        for _ in range(3):
            assert p.parse(s.next_frame()[0], 0)[0] == 0
            frames += 1
    sharding.barrier(dist)
    local_ms = 10.0 + 5.0 * rank
    worst = sharding.max_over_ranks(dist, local_ms, torch)
    total = sharding.sum_over_ranks(dist, frames, torch)
    everyone = sharding.gather_objects(dist, {'rank': rank, 'streams': mine})
    out[rank] = (mine, worst, total, everyone)
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0] == [0, 2, 4] and out[1][0] == [1, 3, 5]
    for r in range(world):
        assert out[r][1] == 15.0      # the slowest rank's time, on every rank
        assert out[r][2] == 18.0      # all frames of all ranks
        assert [e['rank'] for e in out[r][3]] == [0, 1] and out[r][3][1]['streams'] == [1, 3, 5]
    assert sharding.aggregate_fps(18, 15.0) == pytest.approx(1200.0)


def test_core_slices_are_disjoint_and_cover():
    cores = list(range(32))
    for world in (1, 2, 4, 8):
        slices = [sharding.cores_of_rank(r, world, cores) for r in range(world)]
        assert sorted(c for s in slices for c in s) == cores
        assert all(s == list(range(s[0], s[0] + len(s))) for s in slices)       # contiguous: one socket per rank
    assert sharding.cores_of_rank(5, 8, [0, 1, 2]) != []                          # fewer cores than ranks: shared, never empty
    assert sharding.gather_objects(None, 7) == [7]
