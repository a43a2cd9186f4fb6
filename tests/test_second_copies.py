"""Cross-restatement checks: the reference carries TWO independent copies of most
reconstruction primitives -- the decoder's (MobiclipDecoder.cs) and the encoder side's (FrameUtil.cs, MobiEncoder.cs,
MacroBlock.cs).  Both are compiled from the reference's own files (oracle/build_ref.py); here the decoder copy, the
encoder copy and the oracle's restatement are run on the same random inputs and must agree bit for bit:

    CopyBlock MD:418            <->  FrameUtil.GetPBlock FU:105
    IDCT64Px8 / 16 / 3 / 1 MD:3435-3725  <->  MobiEncoder.IDCT64 ME:1012
    IDCT16Px4 / 1 MD:3728-3798  <->  MobiEncoder.IDCT16 ME:1180
    PredictIntra MD:1883-2774   <->  MacroBlock.GetCompvals8x8 / 4x4 MB:630 / 1184
    plane predictors MD:3017-3327  <->  MacroBlock.PredictIntraPlane16x16 / 8x8 / 4x4 MB:1477 / 1630 / 1716

This is the one known-answer source the reference offers that does not come from its decoder."""
import numpy as np
import pytest

from oracle_lib import Oracle, Ref, Ref2, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref/libmobiref.so not built (needs /root/reference)')

W, H, S = 64, 48, 256
SCAN8 = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
SCAN4 = [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]


def _planes(rng, smooth):
    if smooth:
        yy, xx = np.mgrid[0:H, 0:S]
        y = ((yy * 3 + xx * 2) % 200 + rng.integers(0, 16, size=(H, S))).astype(np.uint8).ravel()
    else:
        y = rng.integers(0, 256, size=S * H, dtype=np.uint8)
    return y, rng.integers(0, 256, size=S * H // 2, dtype=np.uint8)


def _pair(rng, smooth=False):
    o, r = Oracle(W, H, 2), Ref(W, H, 2)
    y, uv = _planes(rng, smooth)
    o.set_planes(y, uv)
    r.set_planes(y, uv)
    return o, r


def _window_for_delta(d):
    v = 2 * d if d > 0 else 1 - 2 * d
    k = v.bit_length() - 1
    return int(('0' * k + format(v, 'b')).ljust(32, '0'), 2)


def test_getpblock_equals_copyblock_equals_oracle():
    rng = np.random.default_rng(11)
    n = 0
    for lw in range(4):
        for lh in range(4):
            for phase in range(4):
                o, r = _pair(rng)
                src = rng.integers(0, 256, size=S * H, dtype=np.uint8)
                w, h = 2 << lw, 2 << lh
                dx = 2 * int(rng.integers(-6, 7)) + (phase & 1)
                dy = 2 * int(rng.integers(-4, 5)) + (phase >> 1)
                bx, by = 24, 10
                off = by * S + bx
                assert o.copy_block(0, src, dx, dy, w, h, off) == r.copy_block(0, src, dx, dy, w, h, off) == 1
                twin = Ref2.pblock(src, dx, dy, w, h, off, S)
                assert twin is not None
                assert np.array_equal(twin, r.y.reshape(H, S)[by:by + h, bx:bx + w])
                assert np.array_equal(twin, o.y.reshape(H, S)[by:by + h, bx:bx + w])
                n += 1
    assert n == 64


@pytest.mark.parametrize('n', [8, 4])
def test_encoder_inverse_transform_equals_the_decoders_size_dispatched_ones(n):
    """The decoder picks one of four (8x8) / two (4x4) transforms from the end scan position (MD:2939-2942, 2954-2955);
    the encoder has one general transform.  All of them, and the oracle's, give the same pixels."""
    rng = np.random.default_rng(12 + n)
    scan = SCAN8 if n == 8 else SCAN4
    for trial in range(400):
        o, r = _pair(rng, smooth=True)
        coef = np.zeros(n * n, dtype=np.int32)
        endpos = min(int(rng.choice([1, 2, 3, 5, 10, 11, 14, 21, n * n])), n * n)
        for p in range(endpos):
            if rng.random() < 0.7:
                coef[scan[p]] = int(rng.integers(-40, 41)) * int(rng.integers(8, 40))
        bx, by = 8 * int(rng.integers(0, 5)), 16
        off = by * S + bx
        pred = r.y.reshape(H, S)[by:by + n, bx:bx + n].copy()
        a, b = o.idct(0, n, coef, endpos, off), r.idct(0, n, coef, endpos, off)
        twin = Ref2.idct(n, coef, pred)
        assert a == b and (twin is not None) == bool(b), 'trial %d' % trial   # the clip table aborts all three alike
        if b:
            want = r.y.reshape(H, S)[by:by + n, bx:bx + n]
            assert np.array_equal(twin, want), 'trial %d endpos %d' % (trial, endpos)
            assert np.array_equal(o.y, r.y)


@pytest.mark.parametrize('mode', [0, 1, 3, 4, 5, 6, 7, 8, 10, 11, 13, 14, 15, 16, 17, 18])
def test_encoder_predictors_equal_the_decoders(mode):
    rng = np.random.default_rng(200 + mode)
    n = 8 if mode < 10 else 4
    for trial in range(80):
        o, r = _pair(rng, smooth=trial % 2 == 0)
        plane = int(rng.integers(0, 2)) if mode not in (8, 18) else 0   # 8 / 18 read 13 / 7 pixels of the row above: luma only
        voff = S // 2 if plane and trial % 3 == 0 else 0                 # V lives in the right half of each chroma row
        bx = 16 + n * int(rng.integers(0, 3))
        by = 8 + n * int(rng.integers(0, 2))   # chroma has H/2 = 24 rows
        off = by * S + bx + voff
        before = (r.uv if plane else r.y).copy()
        a, b = o.predict_intra(mode, plane, off, 0), r.predict_intra(mode, plane, off, 0)
        assert a == b == 1
        twin = Ref2.compvals(n, mode, before, bx, by, S, voff)
        assert twin is not None
        want = (r.uv if plane else r.y).reshape(-1, S)[by:by + n, bx + voff:bx + voff + n]
        assert np.array_equal(twin, want), 'mode %d trial %d plane %d' % (mode, trial, plane)
        assert np.array_equal(o.y, r.y) and np.array_equal(o.uv, r.uv)


@pytest.mark.parametrize('n', [16, 8, 4])
def test_encoder_plane_predictors_equal_the_decoders(n):
    """Including deltas large enough to push values outside 0..255, where the decoder ORs unclipped values into a word
    (MD:3064-3074, 3212-3219, 3314-3321)."""
    rng = np.random.default_rng(300 + n)
    agree_overflow = 0
    for trial in range(200):
        o, r = _pair(rng, smooth=trial % 2 == 0)
        d = int(rng.integers(-30, 31)) if trial % 4 else int(rng.integers(-120, 121))
        bx, by = 16, 16
        off = by * S + bx
        before = r.y.copy()
        win = _window_for_delta(d)
        if n == 16:
            a, b = o.plane16(off, win), r.plane16(off, win)
        else:
            a, b = o.predict_intra(2 if n == 8 else 12, 0, off, win), r.predict_intra(2 if n == 8 else 12, 0, off, win)
        assert a == b == 1
        assert np.array_equal(o.y, r.y)
        want = r.y.reshape(H, S)[by:by + n, bx:bx + n]
        twin = Ref2.plane(n, before, off, S, d)
        assert twin is not None
        if np.array_equal(twin, want):
            agree_overflow += 1
        else:
            # the encoder copy stores bytes one by one (truncating each), the decoder packs four unclipped values into a
            # word with OR: they may differ only where a value left 0..255
            assert abs(d) > 30, 'trial %d delta %d: copies differ without overflow' % (trial, d)
    assert agree_overflow >= 150
