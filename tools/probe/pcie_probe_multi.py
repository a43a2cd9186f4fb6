"""Concurrent host<->device copy ceiling of the box, one rank per GPU (torchrun).  Per rank: pinned D2H of one e2e step's
BGRA result (1024 x 400 x 240 x 4 B) and pinned H2D of one step's packed arrays, first one rank at a time, then every rank
at once; rank 0 prints one JSON object.  Used for DESIGN.md "end-to-end ceiling" (how far 8 GPUs' results can be pulled
through the host at all).   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/probe/pcie_probe_multi.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mobiclipdecoder_b200 import sharding  # noqa: E402


def main():
    rank, local, world = sharding.world_from_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    D = dist if world > 1 else None
    cores = sharding.pin_to_cores(local, world) if '--no-affinity' not in sys.argv else sorted(os.sched_getaffinity(0))
    n_out, n_in, reps = 1024 * 400 * 240 * 4, 29_000_000, 6
    d_out = torch.empty(n_out, dtype=torch.uint8, device='cuda')
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device='cuda')
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(d2h, h2d):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            if d2h:
                with torch.cuda.stream(s1):
                    h_out.copy_(d_out, non_blocking=True)
            if h2d:
                with torch.cuda.stream(s2):
                    d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps

    run(True, True)
    alone = {}
    for r in range(world):               # one rank at a time
        sharding.barrier(D, local)
        if r == rank:
            alone = {'d2h_gbps': n_out / run(True, False) / 1e9, 'h2d_gbps': n_in / run(False, True) / 1e9}
    sharding.barrier(D, local)
    together = {}
    for name, a, b in (('d2h', True, False), ('h2d', False, True), ('both', True, True)):
        sharding.barrier(D, local)
        dt = run(a, b)
        worst = sharding.max_over_ranks(D, dt, torch, torch.device('cuda', local))
        together[name] = {'mine_ms': dt * 1e3, 'worst_ms': worst * 1e3}
    mine = {'rank': rank, 'cores': [cores[0], cores[-1]], 'alone': alone, 'together': together}
    everyone = sharding.gather_objects(D, mine)
    if rank == 0:
        agg = {}
        for name, nbytes in (('d2h', n_out), ('h2d', n_in), ('both', n_out)):
            worst = max(e['together'][name]['worst_ms'] for e in everyone)
            agg[name + '_aggregate_gbps'] = world * nbytes / (worst * 1e-3) / 1e9
        agg['bgra_frames_per_s_ceiling'] = world * 1024 / (max(e['together']['both']['worst_ms'] for e in everyone) * 1e-3)
        print(json.dumps({'n_gpus': world, 'host_cores': len(os.sched_getaffinity(0)) if '--no-affinity' in sys.argv else len(cores) * world,
                          'd2h_bytes': n_out, 'h2d_bytes': n_in, 'aggregate': agg, 'ranks': everyone}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
