"""One process, every visible GPU: MobiMultiBatch shards global streams round-robin over a MobiBatch per device, each driven
from its own host thread, and hands results back in global stream order.  On a one-GPU box the same device is listed twice
(two contexts' worth of batches, two host threads): the sharding, threading and merge logic is what is under test, and it is
identical; the driver's N>1 runs exercise distinct devices."""
import numpy as np
import pytest

from mobiclipdecoder_b200 import MobiBatch, MobiMultiBatch
from mobiclipdecoder_b200.workloads import CONFIGS, make_stream
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n > 1 else [0, 0]


def test_multi_batch_bit_exact_in_global_order():
    name, n_streams, n_frames = 'moflex_400x240', 7, 24      # 7 streams over 2+ devices: ragged shards
    w, h, ver, _ = CONFIGS[name]
    gens = [make_stream(name, 300 + s) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    mb = MobiMultiBatch(w, h, ver, n_streams, _devices(), n_threads=2)
    assert sorted(g for own in mb.owned for g in own) == list(range(n_streams))
    for f in range(n_frames):
        frames = [g.next_frame()[0] for g in gens]
        offs, status = mb.decode(frames)
        assert all(st == 0 for st in status)
        want = [None] * n_streams
        for s in range(n_streams):
            ok, off, want[s] = oracles[s].decode(frames[s], 0, f == n_frames - 1)
            assert ok and off == offs[s]
        if f % 8 == 7:
            yuv = mb.read_yuv()
            for s in range(n_streams):
                assert np.array_equal(yuv[s], oracles[s].i420()), 'stream %d frame %d' % (s, f)
    bgra = mb.read_bgra_all()
    for s in range(n_streams):
        assert np.array_equal(bgra[s], want[s]), 'stream %d' % s
    mb.close()


def test_multi_batch_pipelined():
    name, n_streams, n_frames = 'mods_256x192', 6, 10
    w, h, ver, _ = CONFIGS[name]
    gens = [make_stream(name, 400 + s) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    mb = MobiMultiBatch(w, h, ver, n_streams, _devices(), n_threads=2)
    prev = None
    for f in range(n_frames):
        frames = [g.next_frame()[0] for g in gens]
        mb.submit(frames, fmt=MobiBatch.OUT_I420)
        if prev is not None:
            got = mb.fetch()
            for s in range(n_streams):
                assert np.array_equal(got[s], prev[s]), 'stream %d frame %d' % (s, f - 1)
        for s in range(n_streams):
            assert oracles[s].decode(frames[s], 0, False)[0]
        prev = [o.i420().copy() for o in oracles]
    got = mb.fetch()
    for s in range(n_streams):
        assert np.array_equal(got[s], prev[s])
    mb.close()
