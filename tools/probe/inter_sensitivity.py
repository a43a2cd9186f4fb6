"""Where does the inter kernel's time go?  The bench workload with one ingredient of the mix switched off at a time
(generator parameters, include/mobisynth.h): the kernel's CUDA-event time per launch and per inter macroblock.
    python tools/probe/inter_sensitivity.py"""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mobiclipdecoder_b200 import MobiBatch  # noqa: E402
from mobiclipdecoder_b200.workloads import CONFIGS, make_stream  # noqa: E402

VARIANTS = [
    ('bench mix', {}),
    ('no residuals (p_cbp = 0)', dict(p_cbp=0.0)),
    ('8x8 transforms only (p_blk8 = 1)', dict(p_blk8=1.0)),
    ('4x4 transforms only (p_blk8 = 0)', dict(p_blk8=0.0)),
    ('unsplit macroblocks (p_split = 0)', dict(p_split=0.0)),
    ('reference picture 1 only (p_ref1 = 1)', dict(p_ref1=1.0)),
    ('windows inside the picture (p_oob_mv = 0)', dict(p_oob_mv=0.0)),
    ('no intra macroblocks in P-pictures', dict(p_intra_mb=0.0)),
    ('unsplit, picture 1, no residuals', dict(p_split=0.0, p_ref1=1.0, p_cbp=0.0, p_oob_mv=0.0)),
    ('predicted vectors only (p_zero_mv = 1)', dict(p_zero_mv=1.0)),
]


def main():
    name, S, K, Wm = 'moflex_400x240', 1024, 8, 3
    w, h, ver, ov = CONFIGS[name]
    gop = ov['gop']
    out = []
    for label, extra in VARIANTS:
        def one(i):
            st = make_stream(name, 5000 + i, gop_phase=(i * 37) % gop, **extra)
            return [st.next_frame()[0] for _ in range(Wm + K)]
        with ThreadPoolExecutor(os.cpu_count()) as ex:
            streams = list(ex.map(one, range(S)))
        b = MobiBatch(w, h, ver, S, n_threads=os.cpu_count())
        for k in range(Wm + K):
            b.stage([streams[s][k] for s in range(S)])
        b.sync(); b.reset()
        b.replay(0, Wm + K); b.sync(); b.reset()
        b.replay(0, Wm); b.sync()
        b.clear_stats()
        b.set_kernel_timing(True)
        b.replay(Wm, K)
        kt = b.kernel_times()
        st = b.stats()
        b.close()
        mbs = st['inter_mbs'] / K
        row = {'variant': label, 'inter_ms_per_launch': round(kt['inter_ms'] / kt['inter_launches'], 4), 'inter_mbs_per_launch': round(mbs),
               'ns_per_inter_mb': round(kt['inter_ms'] / kt['inter_launches'] * 1e6 / mbs, 3), 'leaves_per_mb': round(st['parts'] / max(1, st['inter_mbs']), 2),
               'coefs_per_mb': round(st['inter_coefs'] / max(1, st['inter_mbs']), 1), 'intra_p_ms': round(kt['intra_ms'] / max(1, kt['intra_launches']), 4),
               'key_ms': round(kt['key_ms'] / max(1, kt['key_launches']), 4)}
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
