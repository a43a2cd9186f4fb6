"""Pins the oracle (oracle/mobi_oracle.c) against the reference's own decoder source compiled here
(oracle/_ref/libmobiref.so, built by oracle/build_ref.py from /root/reference).  The reference ships no tests or
golden vectors, so this differential run IS the pin; its outputs are frozen as hashes in
tests/golden/ (see tools/make_golden.py) for places where oracle/_ref is not available."""
import numpy as np
import pytest

from mobiclipdecoder_b200.workloads import CONFIGS, frames
from oracle_lib import Oracle, Ref, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref/libmobiref.so not built (needs /root/reference)')


@pytest.mark.parametrize('name,seed,n', [('mods_256x192', 0xC0FFEE, 64), ('mods_256x192', 2, 40), ('pframes_256x192', 9, 16),
                                         ('moflex_400x240', 1, 100), ('moflex_400x240', 4, 30), ('moc5_640x480', 1, 34)])
def test_streams_identical(name, seed, n):
    w, h, ver, _ = CONFIGS[name]
    o, r = Oracle(w, h, ver), Ref(w, h, ver)
    for i, (data, key) in enumerate(frames(name, seed, n)):
        ok1, off1, b1 = o.decode(data, 0)
        ok2, off2, b2 = r.decode(data, 0)
        assert ok1 and ok2, 'frame %d aborted (oracle %s, reference %s)' % (i, ok1, ok2)
        assert off1 == off2
        assert o.quantizer == r.quantizer and o.yuvformat == r.yuvformat
        assert np.array_equal(o.y, r.y), 'frame %d luma' % i
        assert np.array_equal(o.uv, r.uv), 'frame %d chroma' % i
        assert np.array_equal(b1, b2), 'frame %d bitmap' % i


def test_stress_parameters_identical():
    """Heavier syntax mix: deep partition trees, many escapes, all references, table 1 residuals, vectors that
    leave the visible picture (stride padding and row wrap)."""
    for name, seed in (('mods_256x192', 11), ('moflex_400x240', 12)):
        w, h, ver, _ = CONFIGS[name]
        o, r = Oracle(w, h, ver), Ref(w, h, ver)
        fr = frames(name, seed, 30, gop=6, p_split=0.6, p_escape=0.3, p_ref1=0.2, p_oob_mv=0.3, p_intra_mb=0.3, p_sub_mb=0.7,
                    p_cbp=0.7, p_blk8=0.4, mean_coefs=9.0, p_dquant=0.5, mv_range=32, p_zero_mv=0.1)
        for i, (data, key) in enumerate(fr):
            ok1, off1, b1 = o.decode(data, 0)
            ok2, off2, b2 = r.decode(data, 0)
            assert ok1 and ok2 and off1 == off2
            assert np.array_equal(o.y, r.y) and np.array_equal(o.uv, r.uv) and np.array_equal(b1, b2), '%s frame %d' % (name, i)


def test_abort_behaviour_identical():
    """Frames the C# code would throw on (-> null Bitmap, MD:325): both must abort, and leave the same Offset."""
    w, h, ver, _ = CONFIGS['moflex_400x240']
    fr = frames('moflex_400x240', 5, 3)
    for bad in (fr[1][0], fr[0][0][:40], fr[0][0][:2], b'\x00\x00\x00\x00'):
        o, r = Oracle(w, h, ver), Ref(w, h, ver)
        ok1, off1, _ = o.decode(bad, 0)
        ok2, off2, _ = r.decode(bad, 0)
        assert ok1 == ok2 and not ok1
    # random garbage: whatever happens must happen identically
    rng = np.random.default_rng(7)
    for t in range(40):
        o, r = Oracle(w, h, ver), Ref(w, h, ver)
        assert o.decode(fr[0][0], 0)[0] and r.decode(fr[0][0], 0)[0]
        junk = rng.integers(0, 256, size=600, dtype=np.uint8).tobytes()
        ok1, off1, _ = o.decode(junk, 0)
        ok2, off2, _ = r.decode(junk, 0)
        assert ok1 == ok2, 'trial %d' % t
        if ok1:
            assert off1 == off2 and np.array_equal(o.y, r.y) and np.array_equal(o.uv, r.uv)
