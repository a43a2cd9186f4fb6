"""GPU-vs-oracle per primitive, through mobi_submit_packed: hand-built packed frames
(tests/hand_frames.py) place one primitive at a time -- every leaf shape x half-pel phase x reference picture incl.
windows that wrap around a pixel row, every transform class of the reference's size dispatch (MD:2938-2942, 2954),
every directional / DC / plane predictor on luma and chroma, and the three decode-order hazards of SURVEY.md 8a(1) --
and the expected picture comes from the oracle's primitives replayed in decode order.  Bit-exact."""
import numpy as np
import pytest

from hand_frames import HandFrame, Leaf, Mb
from mobiclipdecoder_b200 import MobiclipDecoder, MobiParser, MobiclipVersion
from mobiclipdecoder_b200.synth import SynthParams, SynthStream
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

GEOMETRIES = [(240, 64, MobiclipVersion.Moflex3DS), (256, 64, MobiclipVersion.ModsDS)]   # Width < Stride; Width == Stride (flat addressing wraps onto pixels)


class Rig:
    """A GPU decoder and an oracle holding the same six pictures, plus the quantiser tables they were decoded with."""

    def __init__(self, w, h, ver, seed=5, gpu=True):
        self.w, self.h, self.ver = w, h, ver
        self.dec, self.ora, par = (MobiclipDecoder(w, h, ver) if gpu else None), Oracle(w, h, ver), MobiParser(w, h, ver)
        self.S = self.ora.S
        st = SynthStream(SynthParams(w, h, ver, seed, gop=0))
        self.refs = []
        for i in range(6):
            data, _ = st.next_frame()
            if self.dec:
                self.dec.Data, self.dec.Offset = data, 0
                assert self.dec.DecodeFrame(False) is not None, self.dec.last_error()
            assert self.ora.decode(data, 0, False)[0]
            rc, _, pf = par.parse(data, 0)
            assert rc == 0
            self.qtab, self.quant = list(pf.hdr.contents.qtab), pf.hdr.contents.quantizer
            assert self.dec is None or np.array_equal(self.dec.Y[0], self.ora.y)
            self.refs.insert(0, (self.ora.y, self.ora.uv))
        st.close()

    def frame(self, key=False):
        return HandFrame(self.w, self.h, self.S, self.qtab, self.quant, key)

    def run(self, hf, what):
        pf, keep = hf.packed()
        if self.dec is None:
            # host only (tests/test_hand_frames.py): the frame must pass the checks mobi_submit_packed applies, and the oracle's
            # primitives must accept every read it asks for
            import ctypes as C
            from mobiclipdecoder_b200 import _native
            err = C.create_string_buffer(512)
            rc = _native.mobicuda().mobi_packed_validate(self.w, self.h, int(self.ver), C.byref(pf), min(len(self.refs), 6), err, 512)
            assert rc == 0, '%s: %s' % (what, err.value.decode())
            want_y, want_uv = hf.expected(self.ora, self.refs)
            self.refs.insert(0, (want_y, want_uv))
            del self.refs[5:]
            return
        self.dec.SubmitPacked(pf)
        want_y, want_uv = hf.expected(self.ora, self.refs)
        y, uv = self.dec.Y[0], self.dec.UV[0]
        bad = np.flatnonzero(y != want_y)
        assert bad.size == 0, '%s: luma differs at flat %d (row %d, col %d): got %d want %d' % (what, bad[0], bad[0] // self.S, bad[0] % self.S, y[bad[0]], want_y[bad[0]])
        bad = np.flatnonzero(uv != want_uv)
        assert bad.size == 0, '%s: chroma differs at flat %d (row %d, col %d)' % (what, bad[0], bad[0] // self.S, bad[0] % self.S)
        self.refs.insert(0, (want_y, want_uv))
        del self.refs[5:]

    def close(self):
        if self.dec:
            self.dec.close()


def _reads_inside(S, H, off, w, h, dx, dy):
    """CopyBlock's reads (MD:418-456) stay inside the luma and the chroma array: what the reference needs not to throw."""
    first = off + (dy >> 1) * S + (dx >> 1)
    last = first + (h - 1 + (dy & 1)) * S + w - 1 + (dx & 1)
    cdx, cdy = dx >> 1, dy >> 1
    cfirst = off // 2 + (cdy >> 1) * S + (cdx >> 1)
    clast = cfirst + S // 2 + ((h >> 1) - 1 + (cdy & 1)) * S + (w >> 1) - 1 + (cdx & 1)
    return first >= 0 and last < S * H and cfirst >= 0 and clast < S * H // 2


def _vector(rng, rig, mbx, mby, lf_x, lf_y, w, h, phase, wrap):
    """A half-pel vector of the given phase whose reads stay inside the plane arrays (flat addressing); with `wrap` the
    window is pushed out of its pixel row at the picture's left / right edge (it then reads the neighbouring row)."""
    S, H = rig.S, rig.h
    x, y = mbx * 16 + lf_x, mby * 16 + lf_y
    off = y * S + x
    for attempt in range(200):
        ix, iy = int(rng.integers(-7, 8)), int(rng.integers(-7, 8))
        if wrap and mbx == 0 and attempt < 100:
            ix = -int(rng.integers(x + 1, x + 9))
        if wrap and mbx == rig.w // 16 - 1 and attempt < 100:
            ix = int(rng.integers(S - (x + w), S - (x + w) + 9))
        dx, dy = 2 * ix + (phase & 1), 2 * iy + (phase >> 1)
        if y + (dy >> 1) >= 0 and y + (dy >> 1) + h + 1 <= H and _reads_inside(S, H, off, w, h, dx, dy):
            return dx, dy
    raise AssertionError('no vector fits')


@pytest.mark.parametrize('w,h,ver', GEOMETRIES)
def test_motion_compensation_every_shape_phase_reference(w, h, ver, gpu=True):
    """CopyBlock (MD:418-456) for all 16 leaf shapes x 4 half-pel phases x 5 reference pictures, luma and both chroma
    planes, incl. windows that leave their pixel row (row-wrap: load-per-lane path) and macroblocks of up to 64 leaves."""
    rig = Rig(w, h, ver, gpu=gpu)
    rng = np.random.default_rng(11)
    seen, count = set(), [0] * 16   # (per shape: leaves placed so far -> phase and reference are enumerated, not drawn)
    for f in range(10):
        hf = rig.frame()
        for m in range(hf.mbw * hf.mbh):
            mbx, mby = m % hf.mbw, m // hf.mbw
            si = (m + 5 * f) % 16
            lw, lh = 2 << (si & 3), 2 << (si >> 2)
            leaves = []
            for ly in range(0, 16, lh):
                for lx in range(0, 16, lw):
                    phase, ref = count[si] % 4, 1 + (count[si] // 4) % 5
                    count[si] += 1
                    mvx, mvy = _vector(rng, rig, mbx, mby, lx, ly, lw, lh, phase, wrap=(f % 3 == 1))
                    leaves.append(Leaf(lx, ly, lw, lh, ref, mvx, mvy))
                    seen.add((si, phase, ref))
            hf.add(Mb('inter', leaves=leaves, inline=(len(leaves) == 1 and m % 2 == 0)))
        rig.run(hf, 'frame %d' % f)
    assert len(seen) == 16 * 4 * 5, 'not every shape x phase x reference was placed: %d of 320' % len(seen)
    rig.close()


def _recs(rng, positions, small=True):
    return [(p, int(rng.choice([-2, -1, 1, 2] if small else [-3, -2, -1, 1, 2, 3]))) for p in positions]


@pytest.mark.parametrize('w,h,ver', GEOMETRIES)
def test_inverse_transform_size_classes(w, h, ver, gpu=True):
    """The reference dispatches on the scan position of the last coefficient (MD:2938-2942: <= 0 IDCT1Px8, <= 2 IDCT3Px8,
    <= 9 IDCT16Px8, else IDCT64Px8; MD:2954: IDCT1Px4 / IDCT16Px4): every class, on luma and chroma, 8x8- and
    4x4-transformed blocks mixed inside one macroblock (the inter kernel pools the blocks of four macroblocks)."""
    rig = Rig(w, h, ver, gpu=gpu)
    rng = np.random.default_rng(12)
    last8 = [0, 1, 2, 5, 9, 10, 20, 40, 63]     # last scan position of an 8x8-transformed block: all four classes and their edges
    last4 = [0, 1, 3, 15]
    for f in range(4):
        hf = rig.frame()
        for m in range(hf.mbw * hf.mbh):
            blocks = {}
            for blk in range(6):
                pick = (m * 7 + blk * 3 + f) % 5
                if pick == 0:
                    continue   # not coded
                if pick in (1, 2, 3):
                    last = last8[(m + blk + f) % len(last8)]
                    mid = sorted(set(int(x) for x in rng.integers(0, last + 1, size=min(3, last))))
                    blocks[blk] = ('8', _recs(rng, sorted(set(mid + [last]))))
                else:
                    body = {}
                    for sub in range(4):
                        if (m + sub + blk + f) % 3:
                            last = last4[(m + sub + f) % len(last4)]
                            mid = sorted(set(int(x) for x in rng.integers(0, last + 1, size=min(2, last))))
                            body[sub] = _recs(rng, sorted(set(mid + [last])))
                    if body:
                        blocks[blk] = ('4', body)
            mv = (0, 0) if m % 3 else _vector(rng, rig, m % hf.mbw, m // hf.mbw, 0, 0, 16, 16, 1 + 2 * (m % 2), False)
            hf.add(Mb('inter', leaves=[Leaf(0, 0, 16, 16, 1, mv[0], mv[1])], blocks=blocks, inline=True))
        rig.run(hf, 'frame %d' % f)
    rig.close()


def _intra_mb(rng, mbx, mby, hazards, top_row):
    """One intra macroblock: per 8x8 luma block one 8x8 op or four 4x4 ops, chroma ops, residuals on some of them."""
    ops, blocks = [], {}
    legal8 = [3] if top_row else [0, 1, 2, 3, 4, 5, 6, 7, 8]
    if not top_row and rng.random() < 0.15:
        ops.append((20, 0, 0, 0, int(rng.integers(-40, 41))))   # 16x16 plane predictor (MD:3017), then per-block residuals only
        for k in range(4):
            ops.append((9, 0, (k & 1) * 2, (k >> 1) * 2, 0))
            if rng.random() < 0.5:
                blocks[k] = ('8', _recs(rng, [0, int(rng.integers(1, 12))]))
    else:
        for k in range(4):
            x4, y4 = (k & 1) * 2, (k >> 1) * 2
            if rng.random() < 0.5:
                mode = int(rng.choice(legal8))
                if hazards and k == 3:
                    mode = 8            # hazard (i): block 3 reads row y+7, columns x+16..x+20 of the right-hand macroblock
                ops.append((mode, 0, x4, y4, int(rng.integers(-30, 31)) if mode == 2 else 0))
                if rng.random() < 0.6:
                    blocks[k] = ('8', _recs(rng, sorted(set([0, int(rng.integers(0, 20))]))))
            else:
                body = {}
                for j in range(4):
                    mode = 13 if top_row else int(rng.choice([10, 11, 12, 13, 14, 15, 16, 17, 18]))
                    if hazards and j == 3:
                        mode = 18       # hazard (ii): sub-block 3 reads row 3 of the unit to its right
                    if hazards and k == 3 and j == 1:
                        mode = 18       # hazard (iii): sub-block 1 of block 3 reads columns x+16..x+18 of the right-hand macroblock
                    ops.append((mode, 0, x4 + (j & 1), y4 + (j >> 1), int(rng.integers(-20, 21)) if mode == 12 else 0))
                    if rng.random() < 0.4:
                        body[j] = _recs(rng, sorted(set([0, int(rng.integers(0, 8))])))
                if body:
                    blocks[k] = ('4', body)
    for plane in (1, 2):
        blk = 3 + plane
        if not top_row and rng.random() < 0.25:
            ops.append((2, plane, 0, 0, int(rng.integers(-30, 31))))   # 8x8 plane predictor on chroma (MD:3168), residual-only op after it
            ops.append((9, plane, 0, 0, 0))
            if rng.random() < 0.5:
                blocks[blk] = ('8', _recs(rng, [0, 3]))
        elif rng.random() < 0.6:
            mode = 3 if top_row else int(rng.choice([0, 1, 3, 4, 5, 6, 7]))
            ops.append((mode, plane, 0, 0, 0))
            if rng.random() < 0.5:
                blocks[blk] = ('8', _recs(rng, sorted(set([0, int(rng.integers(0, 10))]))))
        else:
            mode = 13 if top_row else int(rng.choice([10, 11, 13, 14, 15, 16, 17]))
            body = {}
            for j in range(4):
                ops.append((mode, plane, j & 1, j >> 1, 0))
                if rng.random() < 0.5:
                    body[j] = _recs(rng, sorted(set([0, int(rng.integers(0, 6))])))
            if body:
                blocks[blk] = ('4', body)
    return Mb('intra', ops=ops, blocks=blocks)


@pytest.mark.parametrize('w,h,ver', GEOMETRIES)
def test_intra_predictors_and_decode_order_hazards(w, h, ver, gpu=True):
    """PredictIntra (MD:1883-2774) modes 0-8 / 10-18, the three plane predictors (MD:3017-3327), predict-then-residual per
    block (MD:2898-2956), in P-pictures where intra macroblocks sit between inter ones -- so the 'future pixels are zero'
    hazards of SURVEY.md 8a(1) meet pixels the inter kernel has ALREADY written -- and in I-pictures (k_intra_key)."""
    rig = Rig(w, h, ver, gpu=gpu)
    rng = np.random.default_rng(13)
    for f in range(6):
        key = f in (2, 5)
        hf = rig.frame(key=key)
        for m in range(hf.mbw * hf.mbh):
            mbx, mby = m % hf.mbw, m // hf.mbw
            if key or (mby > 0 and (mbx + mby + f) % 2 == 0):
                hf.add(_intra_mb(rng, mbx, mby, hazards=(f % 2 == 1 or key) and mby > 0, top_row=(mby == 0)))
            else:
                hf.add(Mb('inter', leaves=[Leaf(0, 0, 16, 16, 1 + (m + f) % 3, *_vector(rng, rig, mbx, mby, 0, 0, 16, 16, 0, False))]))
        rig.run(hf, 'frame %d (%s)' % (f, 'I' if key else 'P'))
    rig.close()
