"""Primitive-level differential tests, oracle vs compiled reference source: CopyBlock (MD:418) at all 16 sizes x
4 half-pel phases, the size-dispatched inverse transforms (MD:2939-2942, 2954-2955: the 1/3/16-coefficient
variants against one general transform), every intra predictor (MD:1883-2774) incl. the plane predictors with
their unclipped byte packing (MD:3017-3327)."""
import numpy as np
import pytest

from oracle_lib import Oracle, Ref, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref/libmobiref.so not built (needs /root/reference)')

W, H, S = 64, 48, 256


def _pair(rng, smooth=False):
    o, r = Oracle(W, H, 2), Ref(W, H, 2)
    if smooth:
        yy, xx = np.mgrid[0:H, 0:S]
        y = ((yy * 3 + xx * 2) % 200 + rng.integers(0, 16, size=(H, S))).astype(np.uint8).ravel()
    else:
        y = rng.integers(0, 256, size=S * H, dtype=np.uint8)
    uv = rng.integers(0, 256, size=S * H // 2, dtype=np.uint8)
    o.set_planes(y, uv)
    r.set_planes(y, uv)
    return o, r


def _same(o, r):
    return np.array_equal(o.y, r.y) and np.array_equal(o.uv, r.uv)


def test_copy_block_all_shapes_and_phases():
    rng = np.random.default_rng(1)
    for lw in range(4):
        for lh in range(4):
            for phase in range(4):
                for plane in (0, 1):
                    o, r = _pair(rng)
                    src = rng.integers(0, 256, size=S * H // (2 if plane else 1), dtype=np.uint8)
                    w, h = 2 << lw, 2 << lh
                    if plane:
                        w, h = max(1, w // 2), max(1, h // 2)
                    dx = 2 * int(rng.integers(-6, 7)) + (phase & 1)
                    dy = 2 * int(rng.integers(-4, 5)) + (phase >> 1)
                    off = 10 * S + 24
                    assert o.copy_block(plane, src, dx, dy, w, h, off) == r.copy_block(plane, src, dx, dy, w, h, off) == 1
                    assert _same(o, r)


def test_copy_block_out_of_range_aborts_alike():
    rng = np.random.default_rng(2)
    o, r = _pair(rng)
    src = rng.integers(0, 256, size=S * H, dtype=np.uint8)
    for dx, dy, off in ((0, -4, 0), (0, 2 * H, 0), (1, 1, S * (H - 16) + S - 16)):
        assert o.copy_block(0, src, dx, dy, 16, 16, off) == r.copy_block(0, src, dx, dy, 16, 16, off) == 0


@pytest.mark.parametrize('n', [8, 4])
def test_inverse_transform_variants(n):
    rng = np.random.default_rng(3)
    zz8 = None
    for trial in range(300):
        o, r = _pair(rng, smooth=True)
        coef = np.zeros(n * n, dtype=np.int32)
        # number of leading scan positions that may be non-zero; the reference picks its transform from it
        endpos = int(rng.choice([1, 2, 3, 5, 10, 11, n * n]))
        endpos = min(endpos, n * n)
        scan = SCAN8 if n == 8 else SCAN4
        for p in range(endpos):
            if rng.random() < 0.7:
                coef[scan[p]] = int(rng.integers(-40, 41)) * int(rng.integers(8, 40))
        off = 16 * S + 8 * int(rng.integers(0, 5))
        a, b = o.idct(0, n, coef, endpos, off), r.idct(0, n, coef, endpos, off)
        assert a == b
        if a:
            assert _same(o, r), 'trial %d endpos %d' % (trial, endpos)


# zigzag orders (MobiclipDecoder.cs:3836-3882 via mobi_tables.h); only used to place test coefficients
SCAN8 = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
SCAN4 = [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]


def _window_for_delta(d):
    """32-bit window whose leading bits are the signed Elias-gamma code of d (MD:2998-3015)."""
    v = 2 * d if d > 0 else 1 - 2 * d
    k = v.bit_length() - 1
    bits = '0' * k + format(v, 'b')
    return int(bits.ljust(32, '0'), 2)


@pytest.mark.parametrize('mode', list(range(0, 9)) + list(range(10, 19)))
def test_intra_predictors(mode):
    rng = np.random.default_rng(100 + mode)
    for trial in range(60):
        o, r = _pair(rng, smooth=trial % 2 == 0)
        plane = int(rng.integers(0, 2)) if mode not in (8, 18) else 0
        n = 8 if mode < 10 else 4
        bx = int(rng.integers(0, 3 if plane == 0 else 2)) * 16 + n * int(rng.integers(0, 2))
        by = int(rng.integers(1, 3)) * 8
        off = by * S + bx + (S // 2 if plane and trial % 3 == 0 else 0)
        if trial % 7 == 0 and mode in (3, 13):
            off = bx  # top row: DC availability logic
        if trial % 11 == 0 and mode in (3, 13):
            off = by * S  # left edge
        win = _window_for_delta(int(rng.integers(-40, 41)))
        a, b = o.predict_intra(mode, plane, off, win), r.predict_intra(mode, plane, off, win)
        assert a == b
        if a:
            assert _same(o, r), 'mode %d trial %d plane %d off %d' % (mode, trial, plane, off)


def test_plane16_with_byte_overflow():
    rng = np.random.default_rng(5)
    for trial in range(200):
        o, r = _pair(rng, smooth=trial % 2 == 0)
        off = 16 * S + 16 * int(rng.integers(0, 3))
        d = int(rng.integers(-120, 121))  # large deltas push values outside 0..255: bytes bleed (MD:3064-3074)
        win = _window_for_delta(d)
        a, b = o.plane16(off, win), r.plane16(off, win)
        assert a == b == 1
        assert _same(o, r), 'trial %d delta %d' % (trial, d)
