"""I-pictures whose bits are written by the REFERENCE's own writer: BitWriter.cs (bit packing, Elias-gamma varints)
and MobiEncoder.EncodeDCT (ME:675-765, the coefficient entropy coder with its choice of the three escape forms), both
compiled from the reference's files (oracle/build_ref.py -> mobiref2_bw_*).  The frame syntax around the coefficient blocks
(header, coded-block patterns, predictor modes: MD:222-249, 1759-1880, 2869-2896) is spelled out here with the
reference writer's WriteBits / WriteVarIntUnsigned; every macroblock uses the full-block mode with DC predictors, which
are legal everywhere in the picture.  Returns the frame bytes and the list of coefficient records a parser must find."""
import ctypes as C
import re
import os

import numpy as np

from oracle_lib import REF_SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(name):
    txt = open(os.path.join(ROOT, 'mobiclipdecoder_b200', 'csrc', 'mobi_tables.h')).read()
    body = re.search(r'%s\[\d+\] = \{(.*?)\};' % name, txt, flags=re.S).group(1)
    return [int(x, 0) for x in re.findall(r'0x[0-9A-Fa-f]+|\d+', body)]


class RefBitWriter:
    def __init__(self):
        L = C.CDLL(REF_SO)
        L.mobiref2_bw_create.restype = C.c_void_p
        L.mobiref2_bw_destroy.argtypes = [C.c_void_p]
        L.mobiref2_bw_bits.argtypes = [C.c_void_p, C.c_uint, C.c_int]
        L.mobiref2_bw_uvar.argtypes = [C.c_void_p, C.c_uint]
        L.mobiref2_bw_svar.argtypes = [C.c_void_p, C.c_int]
        L.mobiref2_bw_dct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mobiref2_bw_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.L, self.h = L, L.mobiref2_bw_create()

    def bits(self, v, n):
        self.L.mobiref2_bw_bits(self.h, v, n)

    def uvar(self, v):
        self.L.mobiref2_bw_uvar(self.h, v)

    def svar(self, v):
        self.L.mobiref2_bw_svar(self.h, int(v))

    def dct(self, levels_in_scan_order):
        a = np.ascontiguousarray(levels_in_scan_order, dtype=np.int32)
        assert self.L.mobiref2_bw_dct(self.h, a.ctypes.data, a.size, 0) == 1

    def bytes(self):
        n = self.L.mobiref2_bw_bytes(self.h, None, 0)
        buf = (C.c_uint8 * max(n, 1))()
        assert self.L.mobiref2_bw_bytes(self.h, buf, n) == n
        self.L.mobiref2_bw_destroy(self.h)
        self.h = None
        return bytes(buf[:n])


def _block(rng, n, budget):
    """Quantised levels of one transform unit in scan order: a few small ones, now and then a long run or a large level so
    that EncodeDCT has to pick its escape forms; sum |level| stays under `budget` (keeps pixel + residual inside the clip table)."""
    lv = np.zeros(n, dtype=np.int32)
    k = 1 + int(rng.integers(0, 6))
    pos = sorted(set(int(x) for x in rng.integers(0, n, size=k)))
    for p in pos:
        r = rng.random()
        mag = int(rng.integers(1, 4)) if r < 0.75 else int(rng.integers(4, 32)) if r < 0.93 else int(rng.integers(32, 60))
        mag = max(1, min(mag, budget))
        budget -= mag
        lv[p] = mag if rng.random() < 0.5 else -mag
        if budget <= 1:
            break
    if not lv.any():
        lv[0] = 1
    return lv


def make_i_picture(width, height, seed, quantizer=12):
    """-> (frame bytes incl. the two pad bytes of MoLiveDemux.cs:353, [(blk, is8, sub, scan_pos, level), ...] in stream order)."""
    rng = np.random.default_rng(seed)
    cbp6_inv = {v: i for i, v in enumerate(_table('MOBI_CBP6_INTRA'))}
    cbp4_tab = _table('MOBI_CBP4_INTRA')
    cbp4_inv = {}
    for i, v in enumerate(cbp4_tab):
        if i >= 1:                       # the varint's leading zero doubles as the "not one 8x8 transform" bit: index 0 is unreachable
            cbp4_inv.setdefault(v, i)
    w = RefBitWriter()
    want = []
    w.bits(1, 1)                          # I-picture (MD:113)
    w.bits(0, 1)                          # YuvFormat (MD:224)
    w.bits(0, 1)                          # VLC table 0 (EncodeDCT writes table 0 only, ME:696)
    w.bits(quantizer, 6)                  # MD:236
    for mb in range((width // 16) * (height // 16)):
        w.bits(0, 1)                      # full-block mode (MD:244-249)
        cbp6 = int(rng.integers(0, 64))
        w.uvar(cbp6_inv[cbp6])            # MD:1761, byte_115FC4
        w.bits(3, 3)                      # luma predictor: DC (MD:1764)

        def coded(blk):
            if rng.random() < 0.6:
                w.bits(1, 1)              # one 8x8 transform (MD:2871)
                lv = _block(rng, 64, 64)
                w.dct(lv)
                want.extend((blk, 1, 0, int(p), int(lv[p])) for p in np.flatnonzero(lv))
            else:
                cbp4 = int(rng.choice([v for v in cbp4_inv if v != 0]))
                w.uvar(cbp4_inv[cbp4])    # MD:2879, byte_1164F4
                for k in range(4):
                    if (cbp4 >> k) & 1:
                        lv = _block(rng, 16, 16)
                        w.dct(lv)
                        want.extend((blk, 0, k, int(p), int(lv[p])) for p in np.flatnonzero(lv))
        for b in range(4):
            if (cbp6 >> b) & 1:
                coded(b)
        w.bits(3, 3)                      # chroma predictor: DC (MD:1866)
        for b in (4, 5):
            if (cbp6 >> b) & 1:
                coded(b)
    return w.bytes() + b'\0\0', want


# ------------------------------------------------------------------------------------------------------------------
# P-pictures.  The partition tree is written the way Analyzer.PBlock.Encode does it (Analyzer.cs:528-565): one code per
# node from the ENCODER's inverse tables HuffEncodeValTable / HuffEncodeBitTable (Analyzer.cs:472-526; symbol 9 = left /
# right halves, 8 = top / bottom halves, 0 = predicted vector on picture 1, 1..5 = picture number followed by the vector
# difference as two signed varints), coded-block patterns through the encoder's inverse maps REV_byte_116160 /
# REV_byte_1165C4 (MobiEncoder.cs:149-161, 330-374), coefficient blocks by MobiEncoder.EncodeDCT, all bits through the
# reference's BitWriter.  The tables are the frozen copies in tests/golden/tables_partition_encoder.json.
# ------------------------------------------------------------------------------------------------------------------
def _enc_tables():
    import json
    return json.load(open(os.path.join(ROOT, 'tests', 'golden', 'tables_partition_encoder.json')))


def _svar(w, v):
    w.L.mobiref2_bw_svar(w.h, int(v))


def _med3(a, b, c):
    return sorted((a, b, c))[1]


def make_p_picture(width, height, seed, n_prev, p_split=0.35):
    """-> (frame bytes, leaves [(mb, x, y, w, h, ref, mvx, mvy), ...] in stream order, coefficient records like make_i_picture).
    n_prev = pictures decoded so far (references 1..min(n_prev, 5) exist).  Vectors keep every window inside the visible picture."""
    rng = np.random.default_rng(seed)
    T = _enc_tables()
    val, bits = T['value'], T['bits']
    rev6, rev4 = T['REV_byte_116160'], T['REV_byte_1165C4']
    mbw, mbh = width // 16, height // 16
    w = RefBitWriter()
    leaves, want = [], []
    w.bits(0, 1)            # P-picture (MD:113)
    _svar(w, 0)             # quantiser delta 0 (MD:122)
    prev_row = [(0, 0)] * (mbw + 2)   # last-leaf vectors of the row above, by column (+1: the entry right of the last column stays 0)
    for mby in range(mbh):
        cur_row = [(0, 0)] * (mbw + 2)
        left = (0, 0)
        for mbx in range(mbw):
            mb = mby * mbw + mbx
            top, topright = prev_row[mbx], prev_row[mbx + 1]
            px, py = _med3(left[0], top[0], topright[0]), _med3(left[1], top[1], topright[1])   # MD:163-208
            last = [(0, 0)]

            def legal(x, y, bw, bh, mx, my):
                x0, y0 = mbx * 16 + x + (mx >> 1), mby * 16 + y + (my >> 1)
                return x0 >= 0 and y0 >= 0 and x0 + bw + (mx & 1) <= width and y0 + bh + (my & 1) <= height

            def node(x, y, bw, bh):
                iw, ih = bw.bit_length() - 2, bh.bit_length() - 2      # SizeToIdx: 2 -> 0 ... 16 -> 3
                can_lr, can_tb = bits[iw][ih][9] > 0, bits[iw][ih][8] > 0
                if (can_lr or can_tb) and rng.random() < p_split:
                    lr = can_lr and (not can_tb or rng.random() < 0.5)
                    sym = 9 if lr else 8
                    w.bits(val[iw][ih][sym], bits[iw][ih][sym])
                    if lr:
                        node(x, y, bw // 2, bh); node(x + bw // 2, y, bw // 2, bh)
                    else:
                        node(x, y, bw, bh // 2); node(x, y + bh // 2, bw, bh // 2)
                    return
                if rng.random() < 0.3 and legal(x, y, bw, bh, px, py):
                    w.bits(val[iw][ih][0], bits[iw][ih][0])
                    ref, mx, my = 1, px, py
                else:
                    ref = 1 if n_prev == 1 or rng.random() < 0.7 else int(rng.integers(1, min(n_prev, 5) + 1))
                    for _ in range(50):
                        mx, my = int(rng.integers(-9, 10)), int(rng.integers(-9, 10))
                        if legal(x, y, bw, bh, mx, my):
                            break
                    else:
                        mx = my = 0
                    w.bits(val[iw][ih][ref], bits[iw][ih][ref])
                    _svar(w, mx - px); _svar(w, my - py)
                leaves.append((mb, x, y, bw, bh, ref, mx, my))
                last[0] = (mx, my)

            node(0, 0, 16, 16)
            left = last[0]
            cur_row[mbx] = last[0]
            # inter residual (loc_1161A0 MD:1818): pattern through the encoder's inverse map, then per coded 8x8 block
            cbp6 = int(rng.integers(0, 64)) if rng.random() < 0.7 else 0
            w.uvar(rev6[cbp6])
            for b in range(6):
                if not (cbp6 >> b) & 1:
                    continue
                if rng.random() < 0.6:
                    w.bits(1, 1)
                    lv = _block(rng, 64, 48)
                    w.dct(lv)
                    want.extend((mb, b, 1, 0, int(p), int(lv[p])) for p in np.flatnonzero(lv))
                else:
                    cbp4 = int(rng.integers(1, 16))
                    w.uvar(rev4[cbp4])
                    for k in range(4):
                        if (cbp4 >> k) & 1:
                            lv = _block(rng, 16, 12)
                            w.dct(lv)
                            want.extend((mb, b, 0, k, int(p), int(lv[p])) for p in np.flatnonzero(lv))
        prev_row = cur_row
    return w.bytes() + b'\0\0', leaves, want
