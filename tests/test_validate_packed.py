"""mobi_packed_validate (host only): what mobi_submit_packed checks before caller-supplied arrays reach the kernels.
Frames the parser emits pass; each field the kernels index shared or global memory with is then corrupted in turn and
must be rejected (op fields of intra macroblocks, contiguity of the coefficient ranges)."""
import ctypes as C

import pytest

from mobiclipdecoder_b200 import MobiParser, _native
from mobiclipdecoder_b200.workloads import CONFIGS, frames

L = _native.mobicuda()


def _copy(pf):
    """Deep, writable copy of a parser-owned packed frame."""
    hd = _native.FrameHdr.from_buffer_copy(pf.hdr.contents)
    keep = {'hdr': hd}

    def arr(ptr, n, typ):
        a = (typ * max(n, 1))()
        if n:
            C.memmove(a, ptr, n * C.sizeof(typ))
        return a

    keep['mbs'] = arr(pf.mbs, hd.n_mb, _native.Mb)
    keep['parts'] = arr(pf.parts, hd.n_parts, _native.Part)
    keep['ops'] = arr(pf.ops, hd.n_ops, C.c_uint32)
    keep['coefs'] = arr(pf.coefs, hd.n_coefs, _native.Coef)
    keep['intra'] = arr(pf.intra_list, hd.n_intra, C.c_uint32)
    out = _native.PackedFrame(C.pointer(hd), C.cast(keep['mbs'], C.POINTER(_native.Mb)), C.cast(keep['parts'], C.POINTER(_native.Part)),
                              C.cast(keep['ops'], C.POINTER(C.c_uint32)), C.cast(keep['coefs'], C.POINTER(_native.Coef)),
                              C.cast(keep['intra'], C.POINTER(C.c_uint32)))
    return out, keep


def _validate(w, h, ver, pf, pictures):
    err = C.create_string_buffer(512)
    rc = L.mobi_packed_validate(w, h, int(ver), C.byref(pf), pictures, err, 512)
    return rc, err.value.decode()


def _parsed(name, seed, n):
    w, h, ver, _ = CONFIGS[name]
    p = MobiParser(w, h, ver)
    out = []
    for i, (data, key) in enumerate(frames(name, seed, n)):
        rc, _, pf = p.parse(data, 0)
        assert rc == 0
        out.append((_copy(pf), min(i, 6)))
    return w, h, ver, out


@pytest.mark.parametrize('name', ['mods_256x192', 'moflex_400x240', 'moc5_640x480'])
def test_parser_output_is_accepted(name):
    w, h, ver, fr = _parsed(name, 11, 4)
    for (pf, keep), pics in fr:
        rc, msg = _validate(w, h, ver, pf, pics)
        assert rc == 0, msg


def test_reference_beyond_the_ring_is_rejected():
    w, h, ver, fr = _parsed('moflex_400x240', 11, 2)
    (pf, keep), _ = fr[1]
    assert _validate(w, h, ver, pf, 0)[0] == -4   # P-picture, empty ring: MOBI_ERR_REFERENCE


def _first(keep, kind):
    for m in range(keep['hdr'].n_mb):
        if (keep['mbs'][m].info & 3) == kind:
            return m
    raise AssertionError('no macroblock of kind %d' % kind)


def test_intra_op_fields_are_range_checked():
    w, h, ver, fr = _parsed('moflex_400x240', 11, 1)
    (pf, keep), pics = fr[0]
    m = _first(keep, 1) + 30   # an intra macroblock away from the picture's top row
    mb = keep['mbs'][m]
    k = mb.first_sub
    good = keep['ops'][k]
    for bad, what in [((good & ~31) | 21, 'mode'), ((good & ~31) | 31, 'mode'), (good | (3 << 6), 'plane'),
                      ((good & ~(15 << 8)) | (3 << 8) | (3 << 10) | 0, 'outside'), ((good & ~0xFFF) | 20 | (1 << 8), 'outside'),
                      ((good & ~0xFFF) | 8 | (1 << 6), 'chroma'), ((good & ~0xFFF) | 18 | (2 << 6), 'chroma'),
                      ((good & ~0xFFF) | 20 | 32, 'residual')]:
        keep['ops'][k] = bad
        rc, msg = _validate(w, h, ver, pf, pics)
        assert rc == -1 and what in msg, (hex(bad), msg)
    keep['ops'][k] = good
    assert _validate(w, h, ver, pf, pics)[0] == 0
    # a residual flag on a block the macroblock does not code
    nomask = (keep['mbs'][m].info & ~(63 << 18))
    saved = keep['mbs'][m].info
    if any(keep['ops'][mb.first_sub + j] & 32 for j in range((mb.info >> 2) & 127)):
        keep['mbs'][m].info = nomask
        rc, msg = _validate(w, h, ver, pf, pics)
        assert rc == -1 and 'does not code' in msg
        keep['mbs'][m].info = saved
    # a top-reading predictor in the picture's first macroblock row: the reference would throw (MD:1893)
    m0 = 0
    k0 = keep['mbs'][m0].first_sub
    saved = keep['ops'][k0]
    keep['ops'][k0] = (saved & ~0xFFFF)   # predictor 0 (vertical), luma, block (0, 0), no residual
    rc, msg = _validate(w, h, ver, pf, pics)
    assert rc == -5 and 'above/left' in msg
    keep['ops'][k0] = saved


def test_coefficient_ranges_must_be_contiguous_and_tagged():
    w, h, ver, fr = _parsed('moflex_400x240', 11, 2)
    (pf, keep), pics = fr[1]
    hd = keep['hdr']
    m = next(i for i in range(hd.n_mb - 1) if (keep['mbs'][i].info >> 9) & 511 and (keep['mbs'][i + 1].info >> 9) & 511)
    saved = keep['mbs'][m + 1].first_coef
    keep['mbs'][m + 1].first_coef = saved - 1      # overlaps its predecessor
    rc, msg = _validate(w, h, ver, pf, pics)
    assert rc == -1 and ('contiguous' in msg or 'owner' in msg)
    keep['mbs'][m + 1].first_coef = saved
    c = keep['coefs'][keep['mbs'][m].first_coef]
    saved = c.blk
    c.blk = saved ^ (1 << 3)                       # owner tag of another macroblock
    rc, msg = _validate(w, h, ver, pf, pics)
    assert rc == -1 and 'owner' in msg
    c.blk = (saved & ~7) | 6                       # block 6 does not exist
    assert _validate(w, h, ver, pf, pics)[0] == -1
    c.blk = saved
    assert _validate(w, h, ver, pf, pics)[0] == 0


def test_partition_geometry_and_vectors():
    w, h, ver, fr = _parsed('moflex_400x240', 11, 2)
    (pf, keep), pics = fr[1]
    m = _first(keep, 0)
    p = keep['parts'][keep['mbs'][m].first_sub]
    saved = (p.xy, p.shape, p.mvx, p.mvy)
    p.mvy = -4000
    assert _validate(w, h, ver, pf, pics)[0] in (-5, -1)
    p.mvy = saved[3]
    p.shape = (saved[1] & 15) | (7 << 4)
    assert _validate(w, h, ver, pf, pics)[0] in (-4, -1)   # (-1 where the macroblock's inline copy of the leaf disagrees first)
    p.shape = saved[1]
    keep['mbs'][m].info = (keep['mbs'][m].info & ~(127 << 2))   # no partitions at all
    assert _validate(w, h, ver, pf, pics)[0] == -1


def test_luma_ops_must_precede_chroma_ops():
    """The I-picture kernel's luma and chroma warps each walk their own range of a macroblock's ops (decode order, MD:1759-1880)."""
    from hand_frames import HandFrame, Mb
    qtab = [i % 64 | 16 << 8 for i in range(64)] + [i | 16 << 8 for i in range(16)]
    good = HandFrame(16, 16, 256, qtab, 20, key=True).add(Mb('intra', ops=[(3, 0, 0, 0, 0), (3, 0, 2, 0, 0), (3, 1, 0, 0, 0), (3, 2, 0, 0, 0)]))
    bad = HandFrame(16, 16, 256, qtab, 20, key=True).add(Mb('intra', ops=[(3, 0, 0, 0, 0), (3, 1, 0, 0, 0), (3, 0, 2, 0, 0), (3, 2, 0, 0, 0)]))
    pf, keep = good.packed()
    assert _validate(16, 16, 2, pf, 0)[0] == 0
    pf, keep = bad.packed()
    rc, msg = _validate(16, 16, 2, pf, 0)
    assert rc == -1 and 'luma ops must precede' in msg
