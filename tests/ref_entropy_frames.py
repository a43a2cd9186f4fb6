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
