"""Host entropy parser (mobi_parser_*, no GPU): position/state parity with the oracle, structural invariants of
the packed arrays it emits, and the error statuses that stand in for the reference's exceptions."""
import numpy as np
import pytest

from mobiclipdecoder_b200 import MobiParser, MobiError
from mobiclipdecoder_b200.workloads import CONFIGS, frames, make_stream
from oracle_lib import Oracle


@pytest.mark.parametrize('name,n', [('mods_256x192', 40), ('pframes_256x192', 12), ('moflex_400x240', 95), ('moc5_640x480', 32)])
def test_offsets_counts_and_state_follow_the_oracle(name, n):
    w, h, ver, _ = CONFIGS[name]
    s, o, p = make_stream(name, 3), Oracle(w, h, ver), MobiParser(w, h, ver)
    for f in range(n):
        data, key = s.next_frame()
        st = s.stats()
        ok, off, _ = o.decode(data, 0, False)
        rc, off2, pf = p.parse(data, 0)
        assert ok and rc == 0, p.last_error()
        hd = pf.hdr.contents
        assert off2 == off and hd.bytes_consumed == off
        assert (hd.flags & 1) == int(key)
        assert hd.quantizer == o.quantizer and hd.yuv_format == o.yuvformat
        assert hd.n_mb == st.n_mb == (w // 16) * (h // 16)
        assert hd.n_intra == st.n_intra_mb and hd.n_parts == st.n_leaves and hd.n_coefs == st.n_coefs
        _check_structure(pf, w, h)


def _check_structure(pf, w, h):
    hd = pf.hdr.contents
    cover_ok = True
    n_intra = 0
    part_cursor = op_cursor = coef_cursor = 0
    for m in range(hd.n_mb):
        mb = pf.mbs[m]
        kind, nsub, nco = mb.info & 3, (mb.info >> 2) & 127, (mb.info >> 9) & 511
        assert mb.first_coef == coef_cursor
        coef_cursor += nco
        if kind == 1:
            assert mb.first_sub == op_cursor and pf.intra_list[mb.intra_rank] == m and mb.intra_rank == n_intra
            op_cursor += nsub
            n_intra += 1
            last_flags = sum(1 for k in range(nco) if pf.coefs[mb.first_coef + k].blk & 0x40)
            res_ops = sum(1 for k in range(nsub) if pf.ops[mb.first_sub + k] & 32)
            assert last_flags == res_ops  # one "last" coefficient per residual-carrying op
            continue
        assert mb.first_sub == part_cursor and nsub >= 1
        part_cursor += nsub
        area = 0
        grid = np.zeros((16, 16), dtype=np.int32)
        for k in range(nsub):
            pt = pf.parts[mb.first_sub + k]
            x, y = (pt.xy & 15) * 2, (pt.xy >> 4) * 2
            pw, ph, ref = 2 << (pt.shape & 3), 2 << ((pt.shape >> 2) & 3), pt.shape >> 4
            assert 1 <= ref <= 5 and x + pw <= 16 and y + ph <= 16
            grid[y:y + ph, x:x + pw] += 1
            area += pw * ph
        assert area == 256 and (grid == 1).all()  # leaves tile the macroblock exactly
        mask = (mb.info >> 18) & 63
        seen = 0
        for k in range(nco):
            seen |= 1 << (pf.coefs[mb.first_coef + k].blk & 7)
        assert seen == mask
    assert part_cursor == hd.n_parts and op_cursor == hd.n_ops and coef_cursor == hd.n_coefs and n_intra == hd.n_intra
    assert cover_ok


def test_error_statuses():
    w, h, ver, _ = CONFIGS['moflex_400x240']
    fr = frames('moflex_400x240', 5, 3)
    p = MobiParser(w, h, ver)
    rc, off, pf = p.parse(fr[1][0], 0)            # P-picture with an empty ring: Y[1] == null (MD:413)
    assert rc == -4 and pf is None and off == 0
    rc, off, pf = p.parse(fr[0][0][:40], 0)       # truncated: IndexOutOfRange in ReadU16LE (IO:39)
    assert rc == -3 and 'past end' in p.last_error()
    rc, off, pf = p.parse(fr[0][0], 0)            # state was rolled back: the stream still parses from its I-picture
    assert rc == 0
    for data, _ in fr[1:]:
        assert p.parse(data, 0)[0] == 0
    with pytest.raises(MobiError):
        MobiParser(400, 240, 0)                   # VxDS: DecodeVXS1 is a stub in the reference (MD:63-95)
    with pytest.raises(MobiError):
        MobiParser(401, 240, 2)
    with pytest.raises(MobiError):
        MobiParser(2048, 240, 2)


def test_offset_input_is_honoured():
    """MOC5 callers pass the whole file and a moving Offset (MobiclipDecoder/Form1.cs:291-302)."""
    w, h, ver, _ = CONFIGS['moflex_400x240']
    fr = [d for d, _ in frames('moflex_400x240', 9, 4)]
    blob, at = b'', []
    for d in fr:
        blob += b'\xAA' * 8   # stand-in for the per-frame 8-byte block header
        at.append(len(blob))
        blob += d
    p, o = MobiParser(w, h, ver), Oracle(w, h, ver)
    for a in at:
        rc, off, _ = p.parse(blob, a)
        ok, off2, _ = o.decode(blob, a, False)
        assert rc == 0 and ok and off == off2 and off > a


def test_garbage_never_crashes_and_agrees_with_the_oracle_on_acceptance():
    w, h, ver, _ = CONFIGS['mods_256x192']
    key = frames('mods_256x192', 1, 1)[0][0]
    rng = np.random.default_rng(3)
    for t in range(200):
        p, o = MobiParser(w, h, ver), Oracle(w, h, ver)
        assert p.parse(key, 0)[0] == 0 and o.decode(key, 0, False)[0]
        junk = rng.integers(0, 256, size=int(rng.integers(2, 900)), dtype=np.uint8).tobytes()
        rc, off, _ = p.parse(junk, 0)
        ok, off2, _ = o.decode(junk, 0, False)
        # the parser may be stricter than the reference only where the reference would need pixels to fail
        # (clip-table overflow); it must never accept what the reference rejects
        if rc == 0:
            assert ok and off == off2, 'trial %d' % t
