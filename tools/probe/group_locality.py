"""Does it pay to walk the streams in GROUPS whose pictures stay in L2 from one step to the next?  1024 streams as G
batches of 1024 / G; per replay every group runs its K staged steps back to back (chained on the device: the next group's
stream waits for the previous group's), so a step's reference pictures were written a few hundred microseconds earlier by
the same group.  G = 1 is bench.py's order (all streams, step by step: 377 MB of pictures between a write and its read).
    python tools/probe/group_locality.py [workload]"""
import os
import sys
import json

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from mobiclipdecoder_b200 import MobiBatch  # noqa: E402
from mobiclipdecoder_b200.workloads import CONFIGS  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else 'moflex_400x240'
    bench.WORKLOAD = wl
    w, h, ver, _ = CONFIGS[wl]
    S, K, Wm, R = 1024, 20, 5, 25
    streams = bench.gen_streams(wl, S, Wm + K, bench.BASE_SEED, os.cpu_count())
    dev = torch.device('cuda', 0)
    out = {}
    for G in (1, 2, 4, 8, 16):
        n = S // G
        batches = [MobiBatch(w, h, ver, n, device=0, n_threads=os.cpu_count()) for _ in range(G)]
        exts = [torch.cuda.ExternalStream(b.cuda_stream(), device=dev) for b in batches]
        for g, b in enumerate(batches):
            for k in range(Wm + K):
                b.stage([streams[g * n + s][k] for s in range(n)])
            b.sync(); b.reset()
        for fmt in (MobiBatch.OUT_BGRA, 0):
            for b in batches:
                b.reset(); b.replay(0, Wm, fmt); b.sync()

            def one_pass():
                for g, b in enumerate(batches):
                    if g:
                        exts[g].wait_stream(exts[g - 1])
                    b.replay(Wm, K, fmt)
                if G > 1:
                    exts[0].wait_stream(exts[-1])      # the next pass starts behind this one
            one_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(exts[0])
            for _ in range(R):
                one_pass()
            e1.record(exts[0])
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / R / K
            out['G=%d %s' % (G, 'bgra' if fmt else 'recon')] = {'ms_per_step_of_1024': round(ms, 4), 'frames_per_s': round(S / (ms * 1e-3))}
            print('groups', G, 'streams/group', n, 'bgra' if fmt else 'recon', '%.4f ms per step of 1024 frames' % ms, flush=True)
        for b in batches:
            b.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
