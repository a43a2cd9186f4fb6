"""TEST INFRASTRUCTURE: hand-built packed frames (include/mobicuda.h) and their expected pictures.

`HandFrame` assembles the arrays mobi_submit_packed takes -- macroblock descriptors, leaves, intra ops, coefficient
records -- from a per-macroblock description, with no bitstream and no parser in between, so that a test can place ONE
primitive (a leaf shape / half-pel phase / reference, a transform class, a predictor, a decode-order hazard) exactly
where it wants it.  `expected()` replays the same description in decode order through the oracle's PRIMITIVES
(CopyBlock, PredictIntra, the plane predictors, the inverse transforms: oracle/mobi_oracle.c hooks, each pinned to the
compiled reference by tests/test_oracle_primitives.py), following the reference's macroblock drivers:
    inter  MD:400-416 (leaf: luma, then U and V at (dx >> 1, dy >> 1), half size), MD:1818-1833 / 2909-2968 (residuals)
    intra  MD:1759-1880, 2776-2902 (per block: predict, then add the residual)
"""
import ctypes as C

import numpy as np

from mobiclipdecoder_b200 import _native as N

SCAN8 = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def window_for_delta(d):
    """32-bit window whose leading bits are the signed Elias-gamma code of d (MD:2998-3015)."""
    v = 2 * d if d > 0 else 1 - 2 * d
    k = v.bit_length() - 1
    return int(('0' * k + format(v, 'b')).ljust(32, '0'), 2)


class Leaf:
    def __init__(self, x, y, w, h, ref, mvx, mvy):
        self.x, self.y, self.w, self.h, self.ref, self.mvx, self.mvy = x, y, w, h, ref, mvx, mvy


class Mb:
    """kind 'inter': leaves = [Leaf]; 'intra': ops = [(mode, plane, x4, y4, delta)], a residual is attached to the op of the
    same (plane, block).  blocks: {blk: ('8', [(pos, level)])} or {blk: ('4', {sub: [(pos, level)]})}, blk 0-3 luma, 4 U, 5 V."""

    def __init__(self, kind, leaves=None, ops=None, blocks=None, inline=False):
        self.kind, self.leaves, self.ops, self.blocks, self.inline = kind, leaves or [], ops or [], blocks or {}, inline


class HandFrame:
    def __init__(self, w, h, stride, qtab, quantizer, key=False):
        self.W, self.H, self.S, self.qtab, self.quant, self.key = w, h, stride, list(qtab), quantizer, key
        self.mbw, self.mbh = w // 16, h // 16
        self.mbs = []

    def add(self, mb):
        self.mbs.append(mb)
        return self

    # ---- packed arrays ------------------------------------------------------------------------------
    def packed(self):
        assert len(self.mbs) == self.mbw * self.mbh
        mbs, parts, ops, coefs, intra = [], [], [], [], []
        max_ref = n_inter_coefs = 0
        for m, mb in enumerate(self.mbs):
            first_coef, mask, m8 = len(coefs), 0, 0
            for blk in sorted(mb.blocks):
                kind, body = mb.blocks[blk]
                mask |= 1 << blk
                units = [(0, body)] if kind == '8' else sorted(body.items())
                if kind == '8':
                    m8 |= 1 << blk
                for sub, recs in units:
                    assert recs, 'a coded transform unit holds at least one coefficient'
                    for k, (pos, level) in enumerate(recs):
                        last = k == len(recs) - 1
                        coefs.append((level, pos | sub << 6, blk | (m & 3) << 3 | (0x40 if last else 0) | (0x80 if kind == '8' else 0)))
            nco = len(coefs) - first_coef
            if mb.kind == 'inter':
                first = len(parts)
                for lf in mb.leaves:
                    lw, lh = lf.w.bit_length() - 2, lf.h.bit_length() - 2
                    parts.append(((lf.x >> 1) | (lf.y >> 1) << 4, lw | lh << 2 | lf.ref << 4, lf.mvx, lf.mvy))
                    max_ref = max(max_ref, lf.ref)
                info = 0 | len(mb.leaves) << 2 | nco << 9 | mask << 18 | (m8 & 15) << 24 | (m8 >> 4) << 29
                rank = 0
                if mb.inline and len(mb.leaves) == 1:
                    lf = mb.leaves[0]
                    info |= 1 << 28
                    rank = (lf.mvx & 0x3FFF) | (lf.mvy & 0x3FFF) << 14 | lf.ref << 28
                mbs.append((info, first, first_coef, rank))
                n_inter_coefs += nco
            else:
                first = len(ops)
                for (mode, plane, x4, y4, delta) in mb.ops:
                    res = self._has_residual(mb, mode, plane, x4, y4)
                    if mode in (9, 19) and not res:
                        continue
                    ops.append((mode & 31) | (32 if res else 0) | plane << 6 | x4 << 8 | y4 << 10 | (delta & 0xFFFF) << 16)
                # neighbours whose pixels the predictors may read: all four (the runtime keeps the ones that exist and are intra)
                mbs.append((1 | (len(ops) - first) << 2 | nco << 9 | mask << 18 | 0xF << 24, first, first_coef, len(intra)))
                intra.append(m)
        hd = N.FrameHdr()
        hd.flags = 1 if self.key else 0
        hd.n_mb, hd.n_parts, hd.n_ops, hd.n_coefs, hd.n_intra = len(mbs), len(parts), len(ops), len(coefs), len(intra)
        hd.quantizer, hd.yuv_format, hd.bytes_consumed, hd.max_ref, hd.n_inter_coefs = self.quant, 0, 0, max_ref, n_inter_coefs
        for i in range(80):
            hd.qtab[i] = self.qtab[i]
        keep = {'hdr': hd}
        a = (N.Mb * max(1, len(mbs)))()
        for i, (info, fs, fc, rk) in enumerate(mbs):
            a[i].info, a[i].first_sub, a[i].first_coef, a[i].intra_rank = info, fs, fc, rk
        keep['mbs'] = a
        p = (N.Part * max(1, len(parts)))()
        for i, (xy, shape, mvx, mvy) in enumerate(parts):
            p[i].xy, p[i].shape, p[i].mvx, p[i].mvy, p[i].pad = xy, shape, mvx, mvy, 0
        keep['parts'] = p
        o = (C.c_uint32 * max(1, len(ops)))(*ops)
        keep['ops'] = o
        c = (N.Coef * (len(coefs) + 40))()   # (+ slack: the intra kernel prefetches past the end of the array it was given a copy of)
        for i, (level, pos, blk) in enumerate(coefs):
            c[i].level, c[i].pos, c[i].blk = level, pos, blk
        keep['coefs'] = c
        il = (C.c_uint32 * max(1, len(intra)))(*intra)
        keep['intra'] = il
        pf = N.PackedFrame(C.pointer(hd), C.cast(a, C.POINTER(N.Mb)), C.cast(p, C.POINTER(N.Part)), C.cast(o, C.POINTER(C.c_uint32)),
                           C.cast(c, C.POINTER(N.Coef)), C.cast(il, C.POINTER(C.c_uint32)))
        return pf, keep

    @staticmethod
    def _has_residual(mb, mode, plane, x4, y4):
        """Does the block this op predicts carry a residual?  8x8 ops: an '8' block; 4x4 ops: that sub-block of a '4' block."""
        if mode == 20:
            return False
        blk = (y4 >> 1) * 2 + (x4 >> 1) if plane == 0 else 3 + plane
        if blk not in mb.blocks:
            return False
        kind, body = mb.blocks[blk]
        if mode < 10:
            return kind == '8'
        return kind == '4' and ((y4 & 1) * 2 + (x4 & 1)) in body

    # ---- expected picture, through the oracle's primitives in decode order ---------------------------
    def _coef_array(self, n, recs):
        """Dequantised coefficients of one transform unit as ReadDCTMatrix leaves them (MD:3424-3429): c[zigzag] = level * scale."""
        out = np.zeros(n * n, dtype=np.int32)
        for pos, level in recs:
            word = self.qtab[pos] if n == 8 else self.qtab[64 + pos]
            out[word & 0xFF] = np.int32(level * (word >> 8))
        return out, recs[-1][0] + 1

    def _residual(self, ora, mb, blk, sub, plane, offset):
        kind, body = mb.blocks[blk]
        if kind == '8':
            coef, end = self._coef_array(8, body)
            assert ora.idct(plane, 8, coef, end, offset)
        else:
            coef, end = self._coef_array(4, body[sub])
            assert ora.idct(plane, 4, coef, end, offset)

    def expected(self, ora, refs):
        """refs[k - 1] = (y, uv) of ring picture k.  Returns (y, uv) of the new picture."""
        S, H = self.S, self.H
        ora.set_planes(np.zeros(S * H, dtype=np.uint8), np.zeros(S * H // 2, dtype=np.uint8))
        for m, mb in enumerate(self.mbs):
            mbx, mby = m % self.mbw, m // self.mbw
            off = mby * 16 * S + mbx * 16
            if mb.kind == 'inter':
                for lf in mb.leaves:
                    o = off + lf.y * S + lf.x
                    ry, ruv = refs[lf.ref - 1]
                    assert ora.copy_block(0, ry, lf.mvx, lf.mvy, lf.w, lf.h, o), 'luma vector outside the plane'
                    assert ora.copy_block(1, ruv, lf.mvx >> 1, lf.mvy >> 1, lf.w // 2, lf.h // 2, o // 2)
                    assert ora.copy_block(1, ruv, lf.mvx >> 1, lf.mvy >> 1, lf.w // 2, lf.h // 2, o // 2 + S // 2)
                for blk in sorted(mb.blocks):
                    kind, body = mb.blocks[blk]
                    plane = 0 if blk < 4 else 1
                    base = off + (blk >> 1) * 8 * S + (blk & 1) * 8 if blk < 4 else off // 2 + (S // 2 if blk == 5 else 0)
                    if kind == '8':
                        self._residual(ora, mb, blk, 0, plane, base)
                    else:
                        for sub in sorted(body):
                            self._residual(ora, mb, blk, sub, plane, base + (sub >> 1) * 4 * S + (sub & 1) * 4)
            else:
                for (mode, plane, x4, y4, delta) in mb.ops:
                    o = off + y4 * 4 * S + x4 * 4 if plane == 0 else off // 2 + (S // 2 if plane == 2 else 0) + y4 * 4 * S + x4 * 4
                    pl = 0 if plane == 0 else 1
                    if mode == 20:
                        assert ora.plane16(o, window_for_delta(delta))
                    elif mode not in (9, 19):
                        assert ora.predict_intra(mode, pl, o, window_for_delta(delta) if mode in (2, 12) else 0), 'predictor %d reads outside the plane' % mode
                    if self._has_residual(mb, mode, plane, x4, y4):
                        blk = (y4 >> 1) * 2 + (x4 >> 1) if plane == 0 else 3 + plane
                        self._residual(ora, mb, blk, (y4 & 1) * 2 + (x4 & 1), pl, o)
        return ora.y, ora.uv
