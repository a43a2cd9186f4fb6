"""TEST INFRASTRUCTURE -- a closed-loop Mobiclip encoder, just enough of the reference's to emit I- and P-picture streams
whose reconstruction is KNOWN on the encoder side: the reference decoder, the oracle and the GPU path must
all reproduce YDec / UVDec from the bytes.

What is restated here (control flow and stream syntax only), with the reference lines it follows:
    MobiEncoder.EncodeFrame                     ME:117-147   picture type, zeroed YDec/UVDec, shifting the past-picture list
    MobiEncoder.SetupQuantizationTables         ME:930-960   quantiser steps from the encoder's own tables
    MobiEncoder.EncodeIntra (stream half)       ME:505-530   I-picture header, per-macroblock mode bit
    MobiEncoder.EncodeBlockIntraFullBlockPMode  ME:532-598   coded-block patterns through REV_byte_115FC4 / 1164F4, predictor ids
    MobiEncoder.EncodePrediction (stream half)  ME:258-424   P-picture header, median vector prediction, REV_byte_116160 / 1165C4
    MacroBlock.SetupDCTs                        MB:224-509   residual -> forward transform -> (int)Math.Round(c / q) -> c * (int)q ->
                                                             inverse transform onto the prediction -> YDec / UVDec, block by block
    Analyzer.PBlock.Encode / GetCompvalsY,U,V   AN:389-470, 528-565   partition codes, reference picture id, vector differences
What is NOT restated: the Analyzer's mode SEARCH (AN:600-1200) and the bit-budget loop (ME:166-256, 432-503); modes,
partitions and vectors are drawn from a seeded generator among the choices the reference encoder itself can emit.

Everything numeric is done by the REFERENCE's encoder-side code, compiled from its files into oracle/_ref (the `prims` object:
tests/oracle_lib.Ref2 -> FrameUtil.GetPBlock, MobiEncoder.DCT64/IDCT64/DCT16/IDCT16, MacroBlock.GetCompvals8x8/4x4,
PredictIntraPlane16x16) and every bit is written by its BitWriter / EncodeDCT (tests/ref_entropy_frames.RefBitWriter).  The
decoder-side code under test is never consulted.  Needs oracle/_ref, i.e. /root/reference at build time.

Known limits, all of the reference encoder's own making: no sub-block intra mode (its writer is unfinished, ME:600-645 "TODO"),
no chroma plane predictor (the writer emits no arguments for it, ME:576), no 4x4 transforms under the 16x16 plane predictor
(MB:299 predicts those with the 4x4 plane predictor, which is not what the stream then says), directional predictors only where
the neighbour BLOCKS the encoder's predictor twins fetch exist (MB:630-1470 index [X - 8], [Y - 8] directly).
"""
import numpy as np

SCAN8 = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]   # MobiConst.DeZigZagTable8x8 (MC:623)
SCAN4 = [0, 4, 1, 2, 5, 8, 12, 9, 6, 3, 7, 10, 13, 14, 11, 15]                                                                        # MobiConst.DeZigZagTable4x4 (MC:645): column-first, unlike the 8x8 one


def _median(a, b, c):
    return sorted((a, b, c))[1]


class MiniEncoder:
    def __init__(self, width, height, quantizer, prims, make_writer, tables, seed, p_inter=0.85, p_split=0.3):
        """tables: tests/golden/tables_partition_encoder.json (frozen copies of the ENCODER's tables: Analyzer.HuffEncodeValTable /
        BitTable AN:472-526, REV_byte_* ME:149-161, 426-434, quantiser rows ME:870-928)."""
        if width % 16 or height % 16:
            raise ValueError('ME:21')
        self.W, self.H = width, height
        self.Q = min(max(int(quantizer), 0xC), 0x34)         # ME:22-23
        self.last_q = self.Q
        self.S = 256 if width <= 256 else 512 if width <= 512 else 1024   # ME:30-32
        self.P, self.new_writer, self.T = prims, make_writer, tables
        self.rng = np.random.default_rng(seed)
        self.past_y, self.past_uv = [], []                   # PastFramesY / UV, newest first (ME:138-145), at most 5
        self.first = True
        self.n_p = 0
        self.p_inter, self.p_split = p_inter, p_split
        self.YDec = self.UVDec = None
        self.stats = {'i_mbs': 0, 'p_inter_mbs': 0, 'p_intra_mbs': 0, 'units8': 0, 'units4': 0, 'uncoded_units': 0, 'leaves': 0, 'plane16': 0}
        self._setup_quant()

    # ---- ME:930-960 ------------------------------------------------------------------------------------
    def _setup_quant(self):
        T = self.T
        r6 = T['enc_byte_119004'][self.Q] + 8
        r5 = T['enc_byte_11903A'][self.Q]
        t4 = [(T['enc_byte_118F94'][(r5 << 4) + i] << r6) >> 8 for i in range(16)]
        r6 -= 2
        t8 = [(T['enc_byte_118DD4'][(r5 << 6) + i] << r6) >> 8 for i in range(64)]
        self.q4 = np.zeros(16, dtype=np.float32)
        self.q8 = np.zeros(64, dtype=np.float32)
        for i in range(16):
            self.q4[SCAN4[i]] = t4[i]                        # "Dezigzag"
        for i in range(64):
            self.q8[SCAN8[i]] = t8[i]

    # ---- one transform unit of SetupDCTs (MB:246-290 and its seven siblings) ---------------------------------
    def _unit(self, n, src, comp):
        """src: n*n source pixels (int), comp: n*n predicted pixels (uint8).  -> (levels in scan order or None when the unit is
        not coded, reconstructed pixels)."""
        q, scan = (self.q8, SCAN8) if n == 8 else (self.q4, SCAN4)
        resid = src.astype(np.int32).ravel() - comp.astype(np.int32).ravel()
        dct = self.P.fdct(n, resid)
        # (int)Math.Round(dct[i] / QTable[i]): int / float is a float32 division, Math.Round rounds half to even
        lev_nat = np.rint((dct.astype(np.float32) / q).astype(np.float64)).astype(np.int32)
        lev_scan = np.array([lev_nat[scan[p]] for p in range(n * n)], dtype=np.int32)
        nz = np.flatnonzero(lev_scan)
        last = int(nz[-1]) if nz.size else 0
        if last == 0 and lev_scan[0] == 0:                   # MB:271 "lastnonzero == 0 && DCT[0] == 0"
            self.stats['uncoded_units'] += 1
            return None, comp.copy()
        real = lev_nat * q.astype(np.int32)                  # realdct[i] * (int)QTable[i]
        rec = self.P.idct(n, real, comp)
        if rec is None:
            raise RuntimeError('the reference inverse transform threw (clip table)')
        self.stats['units8' if n == 8 else 'units4'] += 1
        return lev_scan, rec

    # ---- plane access in the reference's layout ------------------------------------------------------------
    def _put(self, plane, off, n, px):
        v = plane.reshape(-1, self.S)
        r, c = divmod(off, self.S)
        v[r:r + n, c:c + n] = np.asarray(px, dtype=np.uint8).reshape(n, n)

    # ---- blocks of one macroblock: SetupDCTs, then the coded-block pattern and the units' bits -------------------
    def _code_blocks(self, X, Y, src, pred_fn, allow4, keep_empty4):
        """pred_fn(plane_id, bx, by, n, sub) -> n*n predicted pixels, evaluated when the unit is reached (it may read YDec).
        Returns the list of per-8x8-block results: None (not coded) | ('8', levels) | ('4', [levels or None] * 4)."""
        S, out = self.S, []
        for blk in range(6):
            plane = 0 if blk < 4 else blk - 3
            if plane == 0:
                bx, by = X + (blk & 1) * 8, Y + (blk >> 1) * 8
                dec, off0, sp = self.YDec, by * S + bx, src[0][by:by + 8, bx:bx + 8]
            else:
                bx, by = X // 2, Y // 2
                dec, off0, sp = self.UVDec, by * S + bx + (S // 2 if plane == 2 else 0), src[plane][by:by + 8, bx:bx + 8]
            if allow4(blk) and self.rng.random() < 0.35:
                subs = []
                for k in range(4):
                    sx, sy = (k & 1) * 4, (k >> 1) * 4
                    comp = pred_fn(plane, bx + sx, by + sy, 4, (blk, k))
                    lv, rec = self._unit(4, sp[sy:sy + 4, sx:sx + 4], comp)
                    self._put(dec, off0 + sy * S + sx, 4, rec)
                    subs.append(lv)
                # The reference leaves the block "complex" even when none of its 4x4 units is coded (MB:292-337) and then writes
                # pattern 0: REV_byte_1164F4[0] = 2 is a valid intra code, but REV_byte_1165C4[0] = 0 is the one-bit varint '1',
                # which an inter macroblock's reader takes for the "one 8x8 transform" flag (MD:2911) -- a stream the reference
                # decoder cannot parse.  Such an inter block is dropped from the coded-block pattern here (same pixels).
                out.append(('4', subs) if keep_empty4 or any(s is not None for s in subs) else None)
            else:
                comp = pred_fn(plane, bx, by, 8, (blk, None))
                lv, rec = self._unit(8, sp, comp)
                self._put(dec, off0, 8, rec)
                out.append(('8', lv) if lv is not None else None)
        return out

    def _write_blocks(self, w, blocks, rev4):
        for b in blocks:
            if b is None:
                continue
            if b[0] == '8':
                w.bits(1, 1)                                 # "Don't use 4x4 blocks"
                w.dct(b[1])
            else:
                mask = sum(1 << k for k in range(4) if b[1][k] is not None)
                w.uvar(rev4[mask])
                for k in range(4):
                    if b[1][k] is not None:
                        w.dct(b[1][k])

    # ---- an intra macroblock in full-block mode (ME:532-598 + MB:224-509 with UseInterPrediction false) ----------
    def _intra_mb(self, w, X, Y, src):
        S, mbx, mby, mbw = self.S, X // 16, Y // 16, self.W // 16
        interior = 0 < mbx < mbw - 1 and mby > 0
        ymode = int(self.rng.choice([0, 1, 2, 3, 4, 5, 6, 7])) if interior else 3
        uvmode = int(self.rng.choice([0, 1, 3, 4, 5, 6, 7])) if interior else 3
        arg, plane16 = 0, None
        if ymode == 2:
            arg = int(self.rng.integers(-6, 7))
            plane16 = self.P.plane(16, self.YDec, Y * S + X, S, arg).reshape(16, 16)      # MB:230
            self.stats['plane16'] += 1

        def pred(plane, bx, by, n, where):
            if plane == 0:
                if ymode == 2:
                    return plane16[by - Y:by - Y + 8, bx - X:bx - X + 8].copy()              # MB:254
                return self.P.compvals(n, ymode if n == 8 else 10 + ymode, self.YDec, bx, by, S, 0)   # MB:255, 299
            return self.P.compvals(n, uvmode if n == 8 else 10 + uvmode, self.UVDec, bx, by, S, S // 2 if plane == 2 else 0)   # MB:349, 387, 430
        blocks = self._code_blocks(X, Y, src, pred, lambda blk: not (blk < 4 and ymode == 2), True)
        mask = sum(1 << i for i, b in enumerate(blocks) if b is not None)
        w.uvar(self.T['REV_byte_115FC4'][mask])              # ME:541
        w.bits(ymode, 3)                                     # ME:542
        if ymode == 2:
            w.svar(arg)                                      # ME:543
        self._write_blocks(w, blocks[:4], self.T['REV_byte_1164F4'])
        w.bits(uvmode, 3)                                    # ME:575
        self._write_blocks(w, blocks[4:], self.T['REV_byte_1164F4'])

    # ---- an inter macroblock (ME:322-404, AN:528-565) -------------------------------------------------------
    def _inter_mb(self, w, X, Y, src, px, py):
        """Returns the last leaf's vector (what lands in the prediction stack)."""
        S, val, bits = self.S, self.T['value'], self.T['bits']
        predY = np.zeros((16, 16), dtype=np.uint8)
        predC = [np.zeros((8, 8), dtype=np.uint8), np.zeros((8, 8), dtype=np.uint8)]
        last = [(0, 0)]

        def legal(x, y, bw, bh, mx, my):
            x0, y0 = X + x + (mx >> 1), Y + y + (my >> 1)
            return x0 >= 0 and y0 >= 0 and x0 + bw + (mx & 1) <= self.W and y0 + bh + (my & 1) <= self.H

        def node(x, y, bw, bh):
            iw, ih = bw.bit_length() - 2, bh.bit_length() - 2                                 # SizeToIdx AN:524
            can_lr, can_tb = bits[iw][ih][9] > 0, bits[iw][ih][8] > 0
            if (can_lr or can_tb) and self.rng.random() < self.p_split:
                lr = can_lr and (not can_tb or self.rng.random() < 0.5)
                sym = 9 if lr else 8                                                          # Horizontal = left | right (AN:531), Vertical (AN:537)
                w.bits(val[iw][ih][sym], bits[iw][ih][sym])
                if lr:
                    node(x, y, bw // 2, bh); node(x + bw // 2, y, bw // 2, bh)
                else:
                    node(x, y, bw, bh // 2); node(x, y + bh // 2, bw, bh // 2)
                return
            frame = 0 if len(self.past_y) == 1 or self.rng.random() < 0.6 else int(self.rng.integers(0, len(self.past_y)))
            if self.rng.random() < 0.3 and legal(x, y, bw, bh, px, py):
                frame, mx, my = 0, px, py
            else:
                for _ in range(60):
                    mx, my = int(self.rng.integers(-12, 13)), int(self.rng.integers(-12, 13))
                    if legal(x, y, bw, bh, mx, my):
                        break
                else:
                    mx = my = 0
            if frame == 0 and mx == px and my == py:
                w.bits(val[iw][ih][0], bits[iw][ih][0])                                       # AN:547
            else:
                w.bits(val[iw][ih][frame + 1], bits[iw][ih][frame + 1])                       # AN:552-558
                w.svar(mx - px); w.svar(my - py)                                              # AN:560-561
            # GetCompvalsY / U / V (AN:389-470)
            predY[y:y + bh, x:x + bw] = self.P.pblock(self.past_y[frame], mx, my, bw, bh, (Y + y) * S + X + x, S).reshape(bh, bw)
            co = ((Y + y) // 2) * S + (X + x) // 2
            for p in range(2):
                predC[p][y // 2:y // 2 + bh // 2, x // 2:x // 2 + bw // 2] = self.P.pblock(
                    self.past_uv[frame], mx >> 1, my >> 1, bw >> 1, bh >> 1, co + (S // 2 if p else 0), S).reshape(bh // 2, bw // 2)
            last[0] = (mx, my)
            self.stats['leaves'] += 1
        node(0, 0, 16, 16)

        def pred(plane, bx, by, n, where):
            if plane == 0:
                return predY[by - Y:by - Y + n, bx - X:bx - X + n].copy()                     # MB:250, 298
            return predC[plane - 1][by - Y // 2:by - Y // 2 + n, bx - X // 2:bx - X // 2 + n].copy()
        blocks = self._code_blocks(X, Y, src, pred, lambda blk: True, False)
        mask = sum(1 << i for i, b in enumerate(blocks) if b is not None)
        w.uvar(self.T['REV_byte_116160'][mask])                                               # ME:331
        self._write_blocks(w, blocks, self.T['REV_byte_1165C4'])
        return last[0]

    # ---- EncodeFrame (ME:117-147) ---------------------------------------------------------------------------
    def encode_frame(self, y, u, v, force_intra=False, new_quantizer=None):
        """y: H x W, u / v: H/2 x W/2 source pixels (uint8).  Returns the frame's bytes (BitWriter.ToArray(), ME:420-423) and
        leaves the reconstruction in self.YDec / self.UVDec (Stride-wide, U | V side by side as in the decoder)."""
        S, W, H = self.S, self.W, self.H
        self.YDec = np.zeros(S * H, dtype=np.uint8)
        self.UVDec = np.zeros(S * H // 2, dtype=np.uint8)
        src = (np.asarray(y), np.asarray(u), np.asarray(v))
        intra = self.first or force_intra or self.n_p >= 90                                     # ME:123-125
        if new_quantizer is not None:
            self.Q = min(max(int(new_quantizer), 12), 40)                                       # the rate loop's range (ME:236-238)
            self._setup_quant()
        w = self.new_writer()
        if intra:
            self.n_p = 0
            self.past_y, self.past_uv = [], []                                                  # ME:129-130
            w.bits(1, 1); w.bits(1, 1); w.bits(0, 1); w.bits(self.Q, 6)                         # ME:505-509
            self.last_q = self.Q
            for Y in range(0, H, 16):
                for X in range(0, W, 16):
                    w.bits(0, 1)                                                                # ME:518 full-block mode
                    self._intra_mb(w, X, Y, src)
                    self.stats['i_mbs'] += 1
        else:
            self.n_p += 1
            w.bits(0, 1)                                                                        # ME:260
            w.svar(self.Q - self.last_q)                                                        # ME:261
            self.last_q = self.Q
            mbw = W // 16
            stack = [(0, 0)] * (mbw + 2)                                                        # PredictionStack (ME:258)
            for Y in range(0, H, 16):
                off = 0
                for X in range(0, W, 16):
                    a, b, c = stack[off], stack[off + 1], stack[off + 2]                        # ME:269-275
                    off += 1
                    px, py = _median(a[0], b[0], c[0]), _median(a[1], b[1], c[1])               # ME:276-318
                    stack[off] = (0, 0)                                                         # ME:319-320
                    if self.rng.random() < self.p_inter:
                        stack[off] = self._inter_mb(w, X, Y, src, px, py)
                        self.stats['p_inter_mbs'] += 1
                    else:
                        w.bits(0xE >> 1, 5)                                                     # ME:408
                        self._intra_mb(w, X, Y, src)
                        self.stats['p_intra_mbs'] += 1
        w.bits(0, 16)                                                                           # ME:420
        data = w.bytes()
        self.first = False
        self.past_y = ([self.YDec] + self.past_y)[:5]                                           # ME:138-145
        self.past_uv = ([self.UVDec] + self.past_uv)[:5]
        return data
