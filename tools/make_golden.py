#!/usr/bin/env python3
"""Generates tests/golden/*.json: SHA-256 of the cropped I420 planes (and of the BGRA bitmap) of every frame of
the seeded BASELINE configs, as decoded by the reference's own decoder source compiled here (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Runs only where /root/reference exists; the JSON files are committed
so that the oracle and the CUDA path can be checked against the reference where it cannot travel (the GPU box).

    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

GOLDEN = {
    # file: (workload, seed, frames, synth overrides)
    'config1_mods_256x192_seedC0FFEE.json': ('mods_256x192', 0xC0FFEE, 64, {}),
    'config2_pframes_256x192_seed1.json': ('pframes_256x192', 1, 17, {}),
    'config3_moflex_400x240_seed1.json': ('moflex_400x240', 1, 100, {}),
    'config4_moc5_640x480_seed1.json': ('moc5_640x480', 1, 34, {}),
    'stress_moflex_400x240_seed12.json': ('moflex_400x240', 12, 30, dict(gop=6, p_split=0.6, p_escape=0.3, p_ref1=0.2, p_oob_mv=0.3, p_intra_mb=0.3,
                                                                         p_sub_mb=0.7, p_cbp=0.7, p_blk8=0.4, mean_coefs=9.0, p_dquant=0.5, mv_range=32, p_zero_mv=0.1)),
}


def main():
    from mobiclipdecoder_b200 import _build
    _build.build_all()
    from mobiclipdecoder_b200.workloads import CONFIGS, frames
    from oracle_lib import Ref, have_ref
    if not have_ref():
        raise SystemExit('oracle/_ref/libmobiref.so is not built: golden vectors come from the compiled reference source only')
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for fn, (name, seed, n, ov) in GOLDEN.items():
        w, h, ver, _ = CONFIGS[name]
        r = Ref(w, h, ver)
        rows = []
        for i, (data, key) in enumerate(frames(name, seed, n, **ov)):
            ok, off, bgra = r.decode(data, 0)
            assert ok, '%s frame %d aborted in the reference' % (fn, i)
            rows.append({'frame': i, 'key': bool(key), 'bytes': len(data), 'offset_after': off, 'quantizer': r.quantizer,
                         'input_sha256': hashlib.sha256(data).hexdigest(), 'i420_sha256': hashlib.sha256(r.i420().tobytes()).hexdigest(),
                         'y_strided_sha256': hashlib.sha256(r.y.tobytes()).hexdigest(), 'uv_strided_sha256': hashlib.sha256(r.uv.tobytes()).hexdigest(),
                         'bgra_sha256': hashlib.sha256(bgra.tobytes()).hexdigest()})
        doc = {'generator': 'tools/make_golden.py', 'decoder': 'oracle/_ref/libmobiref.so (reference source @ c88b67d3, transliterated by oracle/build_ref.py)',
               'workload': name, 'seed': seed, 'width': w, 'height': h, 'version': int(ver), 'synth_overrides': ov, 'frames': rows}
        with open(os.path.join(out_dir, fn), 'w') as f:
            json.dump(doc, f, indent=1)
        print(fn, len(rows), 'frames')


if __name__ == '__main__':
    main()
