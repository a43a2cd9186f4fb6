"""Known-answer test for the entropy half of the path against the reference's own WRITER: pictures whose bits come from
BitWriter.cs and MobiEncoder.EncodeDCT (compiled from the reference, tests/ref_entropy_frames.py) must (1) decode to
the same planes in the compiled reference decoder and in the oracle, (2) parse, in the product's host parser
(libmobicuda.so, no GPU needed), into exactly the coefficient records that were handed to EncodeDCT -- level, scan
position, block and sub-block tags, 8x8 / 4x4 -- and (3) reconstruct to the same planes on the GPU."""
import numpy as np
import pytest

from oracle_lib import Oracle, Ref, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref/libmobiref.so not built (needs /root/reference)')

CASES = [(64, 48, 1, 12, 2), (64, 48, 2, 18, 2), (256, 192, 3, 12, 2), (400, 240, 4, 14, 2), (256, 192, 8, 16, 1), (64, 48, 9, 13, 1)]   # last field: 2 Moflex3DS, 1 ModsDS


@pytest.mark.parametrize('w,h,seed,q,ver', CASES)
def test_reference_written_picture_decodes_alike_and_parses_to_what_was_written(w, h, seed, q, ver):
    from mobiclipdecoder_b200 import MobiParser
    from ref_entropy_frames import make_i_picture
    data, want = make_i_picture(w, h, seed, q)
    r, o = Ref(w, h, ver), Oracle(w, h, ver)
    ok_r, off_r, bgra_r = r.decode(data, 0)
    ok_o, off_o, bgra_o = o.decode(data, 0)
    assert ok_r and ok_o and off_r == off_o
    assert np.array_equal(r.y, o.y) and np.array_equal(r.uv, o.uv) and np.array_equal(bgra_r, bgra_o)
    par = MobiParser(w, h, ver)
    rc, off, pf = par.parse(data, 0)
    assert rc == 0 and off == off_r
    hdr = pf.hdr.contents
    assert hdr.n_coefs == len(want) and hdr.quantizer == q
    got = []
    for i in range(hdr.n_coefs):
        c = pf.coefs[i]
        got.append((c.blk & 7, c.blk >> 7, (c.pos >> 6) if not (c.blk >> 7) else 0, c.pos & 63, c.level))
    assert got == want
    # escape forms were really exercised: levels beyond the 31 the plain table reaches, and runs it has no code for
    assert max(abs(x[4]) for x in want) > 31
    par.close()


@pytest.mark.gpu
@pytest.mark.parametrize('w,h,seed,q,ver', CASES)
def test_reference_written_picture_on_the_gpu(w, h, seed, q, ver):
    from mobiclipdecoder_b200 import MobiclipDecoder
    from ref_entropy_frames import make_i_picture
    data, want = make_i_picture(w, h, seed, q)
    o, d = Oracle(w, h, ver), MobiclipDecoder(w, h, ver)
    ok, off, bgra = o.decode(data, 0)
    assert ok
    d.Data, d.Offset = data, 0
    bmp = d.DecodeFrame()
    assert bmp is not None and d.Offset == off
    assert np.array_equal(d.Y[0], o.y) and np.array_equal(d.UV[0], o.uv) and np.array_equal(bmp, bgra)
    d.close()


def _sequence(w, h, seed, n_p):
    from ref_entropy_frames import make_i_picture, make_p_picture
    out = [(make_i_picture(w, h, seed, 12)[0], None, None)]
    for k in range(n_p):
        out.append(make_p_picture(w, h, seed * 100 + k, n_prev=k + 1))
    return out


@pytest.mark.parametrize('w,h,seed', [(64, 48, 5), (256, 192, 6), (400, 240, 7)])
def test_p_pictures_written_from_the_encoders_tables(w, h, seed):
    """P-pictures whose partition codes come from the encoder's inverse tables (Analyzer.cs:472-565), patterns from its
    inverse maps, coefficients from EncodeDCT, bits from BitWriter: the compiled reference decoder and the oracle must decode
    them alike, and the product's host parser must recover every leaf (position, size, reference, vector -- i.e. also the
    median prediction MD:163-208) and every coefficient record exactly as written."""
    from mobiclipdecoder_b200 import MobiParser
    r, o, par = Ref(w, h, 2), Oracle(w, h, 2), MobiParser(w, h, 2)
    n_leaves = n_split = 0
    for i, (data, leaves, want) in enumerate(_sequence(w, h, seed, 6)):
        ok_r, off_r, bgra_r = r.decode(data, 0)
        ok_o, off_o, bgra_o = o.decode(data, 0)
        assert ok_r and ok_o and off_r == off_o, 'picture %d' % i
        assert np.array_equal(r.y, o.y) and np.array_equal(r.uv, o.uv) and np.array_equal(bgra_r, bgra_o), 'picture %d' % i
        rc, off, pf = par.parse(data, 0)
        assert rc == 0 and off == off_r
        if leaves is None:
            continue
        hdr = pf.hdr.contents
        assert hdr.n_parts == len(leaves) and hdr.n_intra == 0
        got = []
        for m in range(hdr.n_mb):
            mb = pf.mbs[m]
            for q in range((mb.info >> 2) & 127):
                p = pf.parts[mb.first_sub + q]
                got.append((m, (p.xy & 15) * 2, (p.xy >> 4) * 2, 2 << (p.shape & 3), 2 << ((p.shape >> 2) & 3), p.shape >> 4, p.mvx, p.mvy))
        assert got == leaves, 'picture %d' % i
        gotc = []
        for m in range(hdr.n_mb):
            mb = pf.mbs[m]
            for q in range((mb.info >> 9) & 511):
                c = pf.coefs[mb.first_coef + q]
                gotc.append((m, c.blk & 7, c.blk >> 7, (c.pos >> 6) if not (c.blk >> 7) else 0, c.pos & 63, c.level))
        assert gotc == want, 'picture %d' % i
        n_leaves += len(leaves)
        n_split += sum(1 for l in leaves if l[3] < 16 or l[4] < 16)
    assert n_split > n_leaves // 3 and any(l[5] > 1 for l in leaves)   # split trees and references beyond picture 1 really occur
    par.close()


@pytest.mark.gpu
@pytest.mark.parametrize('w,h,seed', [(64, 48, 5), (256, 192, 6), (400, 240, 7)])
def test_p_pictures_written_from_the_encoders_tables_on_the_gpu(w, h, seed):
    from mobiclipdecoder_b200 import MobiclipDecoder
    o, d = Oracle(w, h, 2), MobiclipDecoder(w, h, 2)
    for i, (data, leaves, want) in enumerate(_sequence(w, h, seed, 6)):
        ok, off, bgra = o.decode(data, 0)
        assert ok
        d.Data, d.Offset = data, 0
        bmp = d.DecodeFrame()
        assert bmp is not None and d.Offset == off, 'picture %d' % i
        assert np.array_equal(d.Y[0], o.y) and np.array_equal(d.UV[0], o.uv) and np.array_equal(bmp, bgra), 'picture %d' % i
    d.close()
