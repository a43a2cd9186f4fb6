"""Seeded primitive-level cases shared by tools/make_golden_primitives.py (which runs them through the reference's
second copies) and tests/test_golden_primitives.py (which runs them through the oracle): the two must produce the same
bytes.  Every case is (kind, arguments); the runner returns the produced pixel block as a uint8 array."""
import hashlib

import numpy as np

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


def window_for_delta(d):
    """32-bit window whose leading bits are the signed Elias-gamma code of d (MobiclipDecoder.cs:2998-3015)."""
    v = 2 * d if d > 0 else 1 - 2 * d
    k = v.bit_length() - 1
    return int(('0' * k + format(v, 'b')).ljust(32, '0'), 2)


def cases(kind):
    """Yields dicts describing the seeded cases of one kind: 'mc', 'idct8', 'idct4', 'pred', 'plane'."""
    rng = np.random.default_rng({'mc': 701, 'idct8': 702, 'idct4': 703, 'pred': 704, 'plane': 705}[kind])
    if kind == 'mc':
        for lw in range(4):
            for lh in range(4):
                for phase in range(4):
                    y, uv = _planes(rng, False)
                    src = rng.integers(0, 256, size=S * H, dtype=np.uint8)
                    yield dict(y=y, uv=uv, src=src, w=2 << lw, h=2 << lh, dx=2 * int(rng.integers(-6, 7)) + (phase & 1),
                               dy=2 * int(rng.integers(-4, 5)) + (phase >> 1), bx=24, by=10)
    elif kind in ('idct8', 'idct4'):
        n = 8 if kind == 'idct8' else 4
        scan = SCAN8 if n == 8 else SCAN4
        for trial in range(300):
            y, uv = _planes(rng, True)
            coef = np.zeros(n * n, dtype=np.int32)
            endpos = min(int(rng.choice([1, 2, 3, 5, 10, 11, 14, 21, n * n])), n * n)
            for p in range(endpos):
                if rng.random() < 0.7:
                    coef[scan[p]] = int(rng.integers(-20, 21)) * int(rng.integers(8, 40))   # stays inside the clip table
            yield dict(y=y, uv=uv, n=n, coef=coef, endpos=endpos, bx=8 * int(rng.integers(0, 5)), by=16)
    elif kind == 'pred':
        for mode in [0, 1, 3, 4, 5, 6, 7, 8, 10, 11, 13, 14, 15, 16, 17, 18]:
            n = 8 if mode < 10 else 4
            for trial in range(30):
                y, uv = _planes(rng, trial % 2 == 0)
                plane = int(rng.integers(0, 2)) if mode not in (8, 18) else 0
                voff = S // 2 if plane and trial % 3 == 0 else 0
                yield dict(y=y, uv=uv, mode=mode, n=n, plane=plane, voff=voff, bx=16 + n * int(rng.integers(0, 3)), by=8 + n * int(rng.integers(0, 2)))
    elif kind == 'plane':
        for n in (16, 8, 4):
            for trial in range(100):
                y, uv = _planes(rng, trial % 2 == 0)
                yield dict(y=y, uv=uv, n=n, delta=int(rng.integers(-12, 13)), bx=16, by=16)   # small deltas: no value leaves 0..255


def run_second_copies(kind, c):
    """The reference's encoder-side copies (oracle/_ref)."""
    from oracle_lib import Ref2
    if kind == 'mc':
        return Ref2.pblock(c['src'], c['dx'], c['dy'], c['w'], c['h'], c['by'] * S + c['bx'], S)
    if kind in ('idct8', 'idct4'):
        n = c['n']
        pred = c['y'].reshape(H, S)[c['by']:c['by'] + n, c['bx']:c['bx'] + n]
        return Ref2.idct(n, c['coef'], pred)
    if kind == 'pred':
        data = c['uv'] if c['plane'] else c['y']
        return Ref2.compvals(c['n'], c['mode'], data, c['bx'], c['by'], S, c['voff'])
    if kind == 'plane':
        return Ref2.plane(c['n'], c['y'], c['by'] * S + c['bx'], S, c['delta'])


def run_oracle(kind, c):
    """The oracle's restatement of the DECODER's copies (oracle/mobi_oracle.c)."""
    from oracle_lib import Oracle
    o = Oracle(W, H, 2)
    o.set_planes(c['y'], c['uv'])
    if kind == 'mc':
        assert o.copy_block(0, c['src'], c['dx'], c['dy'], c['w'], c['h'], c['by'] * S + c['bx']) == 1
        return o.y.reshape(H, S)[c['by']:c['by'] + c['h'], c['bx']:c['bx'] + c['w']]
    if kind in ('idct8', 'idct4'):
        n = c['n']
        assert o.idct(0, n, c['coef'], c['endpos'], c['by'] * S + c['bx']) == 1
        return o.y.reshape(H, S)[c['by']:c['by'] + n, c['bx']:c['bx'] + n]
    if kind == 'pred':
        n = c['n']
        assert o.predict_intra(c['mode'], c['plane'], c['by'] * S + c['bx'] + c['voff'], 0) == 1
        return (o.uv if c['plane'] else o.y).reshape(-1, S)[c['by']:c['by'] + n, c['bx'] + c['voff']:c['bx'] + c['voff'] + n]
    if kind == 'plane':
        n, off, win = c['n'], c['by'] * S + c['bx'], window_for_delta(c['delta'])
        assert (o.plane16(off, win) if n == 16 else o.predict_intra(2 if n == 8 else 12, 0, off, win)) == 1
        return o.y.reshape(H, S)[c['by']:c['by'] + n, c['bx']:c['bx'] + n]


def digests(runner):
    out = {}
    for kind in ('mc', 'idct8', 'idct4', 'pred', 'plane'):
        hsh, n = hashlib.sha256(), 0
        for c in cases(kind):
            r = runner(kind, c)
            assert r is not None, (kind, n)
            hsh.update(np.ascontiguousarray(r, dtype=np.uint8).tobytes())
            n += 1
        out[kind] = '%s:%d' % (hsh.hexdigest(), n)
    return out
