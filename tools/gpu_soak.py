"""Longer randomised parity soak than the test suite affords: many seeds, long streams, every workload, single-stream and
lock-step batch paths (incl. steps whose I-pictures go by ticket), Y / U / V / BGRA / Offset against the oracle.
    python tools/gpu_soak.py [seeds] [frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from mobiclipdecoder_b200 import MobiBatch, MobiclipDecoder  # noqa: E402
from mobiclipdecoder_b200.workloads import CONFIGS, make_stream  # noqa: E402
from oracle_lib import Oracle  # noqa: E402


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    t0, checked = time.time(), 0
    for name in ('moflex_400x240', 'mods_256x192', 'moc5_640x480'):
        w, h, ver, _ = CONFIGS[name]
        for seed in range(seeds):
            extra = {} if seed % 2 == 0 else dict(p_oob_mv=0.2, p_split=0.5, p_intra_mb=0.25, gop=7 + seed)
            st, dec, ora = make_stream(name, 9000 + seed, **extra), MobiclipDecoder(w, h, ver), Oracle(w, h, ver)
            for f in range(n_frames if name != 'moc5_640x480' else n_frames // 3):
                data = st.next_frame()[0]
                dec.Data, dec.Offset = data, 0
                bmp = dec.DecodeFrame()
                ok, off, want = ora.decode(data, 0)
                assert ok and bmp is not None and dec.Offset == off, (name, seed, f)
                assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv), (name, seed, f, 'planes')
                assert np.array_equal(bmp, want), (name, seed, f, 'bitmap')
                checked += 1
            dec.close()
        # lock-step batch: 200 streams (more I-pictures than SMs in step 0 and at the common GOP boundary), ragged GOP phases otherwise
        S = 200
        gens = [make_stream(name, 12000 + s, gop=(12 if s < 160 else 5 + s % 9)) for s in range(S)]
        oras = [Oracle(w, h, ver) for _ in range(S)]
        b = MobiBatch(w, h, ver, S, n_threads=os.cpu_count())
        for f in range(26 if name != 'moc5_640x480' else 8):
            frames = [g.next_frame()[0] for g in gens]
            offs, status = b.decode(frames)
            assert all(s == 0 for s in status)
            for s in range(S):
                ok, off, _ = oras[s].decode(frames[s], 0, False)
                assert ok and off == offs[s]
            if f % 4 == 0 or f in (12, 13, 24, 25):
                got = b.read_yuv()
                for s in range(S):
                    assert np.array_equal(got[s], oras[s].i420()), (name, 'batch', s, f)
                    checked += 1
        b.close()
        print('%s ok (%d pictures compared so far, %.0f s)' % (name, checked, time.time() - t0), flush=True)
    print('soak passed: %d pictures compared' % checked)


if __name__ == '__main__':
    main()
