"""Golden vectors (tests/golden/*.json, made by tools/make_golden.py from the compiled reference source):
the oracle here on CPU, the CUDA path on the GPU box where /root/reference does not exist."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from mobiclipdecoder_b200.workloads import frames
from oracle_lib import Oracle

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', '*.json'))
                if not os.path.basename(p).startswith(('primitives_', 'tables_')))   # stream-level files only (primitives_*.json, tables_*.json: test_golden_primitives.py)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_golden(path):
    g = json.load(open(path))
    o = Oracle(g['width'], g['height'], g['version'])
    fr = frames(g['workload'], g['seed'], len(g['frames']), **g['synth_overrides'])
    for row, (data, key) in zip(g['frames'], fr):
        assert hashlib.sha256(data).hexdigest() == row['input_sha256'], 'the generator no longer reproduces the golden input'
        ok, off, bgra = o.decode(data, 0)
        assert ok and off == row['offset_after'] and o.quantizer == row['quantizer']
        assert _sha(o.i420()) == row['i420_sha256'], 'frame %d' % row['frame']
        assert _sha(o.y) == row['y_strided_sha256'] and _sha(o.uv) == row['uv_strided_sha256']
        assert _sha(bgra) == row['bgra_sha256']


@pytest.mark.gpu
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_matches_golden(path):
    from mobiclipdecoder_b200 import MobiclipDecoder
    g = json.load(open(path))
    d = MobiclipDecoder(g['width'], g['height'], g['version'])
    fr = frames(g['workload'], g['seed'], len(g['frames']), **g['synth_overrides'])
    for row, (data, key) in zip(g['frames'], fr):
        assert hashlib.sha256(data).hexdigest() == row['input_sha256']
        d.Data, d.Offset = data, 0
        bmp = d.DecodeFrame()
        assert bmp is not None and d.Offset == row['offset_after'] and d.Quantizer == row['quantizer']
        assert _sha(d.Y[0]) == row['y_strided_sha256'] and _sha(d.UV[0]) == row['uv_strided_sha256'], 'frame %d' % row['frame']
        y, u, v = d.ReadYuv()
        assert _sha(np.concatenate([y.ravel(), u.ravel(), v.ravel()])) == row['i420_sha256']
        assert _sha(bmp) == row['bgra_sha256']
    d.close()
