"""Parity tests proper: the CUDA path, called through the C ABI (libmobicuda.so via ctypes), against the CPU
oracle on the same seeded inputs.  Bit-exact on Y, U, V (strided planes incl. padding), on Offset/Quantizer, and
on the BGRA bitmap (binary32, source order; tolerance 0)."""
import numpy as np
import pytest

from mobiclipdecoder_b200 import MobiBatch, MobiclipDecoder, MobiParser
from mobiclipdecoder_b200.workloads import CONFIGS, frames, make_stream
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _diff_report(a, b, S):
    idx = np.flatnonzero(a != b)
    if idx.size == 0:
        return 'equal'
    return '%d bytes differ; first at flat %d (row %d, col %d): got %d want %d' % (idx.size, idx[0], idx[0] // S, idx[0] % S, a[idx[0]], b[idx[0]])


@pytest.mark.parametrize('name,n_frames', [('mods_256x192', 64), ('pframes_256x192', 24), ('moflex_400x240', 100), ('moc5_640x480', 34)])
def test_single_stream_bit_exact(name, n_frames):
    w, h, ver, _ = CONFIGS[name]
    fr = frames(name, 0xC0FFEE, n_frames)
    dec = MobiclipDecoder(w, h, ver)
    ora = Oracle(w, h, ver)
    for i, (data, key) in enumerate(fr):
        dec.Data, dec.Offset = data, 0
        bmp = dec.DecodeFrame()
        ok, off, want_bmp = ora.decode(data, 0)
        assert ok, 'generator produced an out-of-contract frame %d' % i
        assert bmp is not None, 'frame %d: %s' % (i, dec.last_error())
        assert dec.Offset == off
        assert dec.Quantizer == ora.quantizer and dec.YuvFormat == ora.yuvformat
        y, uv = dec.Y[0], dec.UV[0]
        assert np.array_equal(y, ora.y), 'frame %d (%s) luma: %s' % (i, 'I' if key else 'P', _diff_report(y, ora.y, dec.Stride))
        assert np.array_equal(uv, ora.uv), 'frame %d (%s) chroma: %s' % (i, 'I' if key else 'P', _diff_report(uv, ora.uv, dec.Stride))
        assert np.array_equal(bmp, want_bmp), 'frame %d bitmap differs' % i
    ty, tu, tv = dec.ReadYuv()
    oy, ou, ov = ora.crop()
    assert np.array_equal(ty, oy) and np.array_equal(tu, ou) and np.array_equal(tv, ov)
    dec.close()


def test_batch_lockstep_bit_exact():
    name, n_streams, n_frames = 'moflex_400x240', 12, 20
    w, h, ver, _ = CONFIGS[name]
    streams = [frames(name, 100 + s, n_frames, gop=7 + s % 3) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams, n_threads=4)
    for f in range(n_frames):
        offs, status = b.decode([streams[s][f][0] for s in range(n_streams)])
        assert all(st == 0 for st in status)
        got = b.read_yuv()
        for s in range(n_streams):
            ok, off, _ = oracles[s].decode(streams[s][f][0], 0, False)
            assert ok and offs[s] == off
            assert np.array_equal(got[s], oracles[s].i420()), 'stream %d frame %d' % (s, f)
    bg = b.read_bgra_all()
    for s in range(n_streams):
        assert np.array_equal(bg[s], b.read_bgra(s))
    b.close()


def test_staged_replay_matches_live_decode():
    name, n_streams, n_frames = 'pframes_256x192', 16, 12
    w, h, ver, _ = CONFIGS[name]
    streams = [frames(name, 1 + s, n_frames) for s in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams, n_threads=2)
    for f in range(n_frames):
        b.stage([streams[s][f][0] for s in range(n_streams)])
    assert b.staged_steps() == n_frames
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    for s in range(n_streams):
        for f in range(n_frames):
            assert oracles[s].decode(streams[s][f][0], 0, False)[0]
    want = np.stack([o.i420() for o in oracles])
    for _ in range(2):  # replay twice: the second pass must not depend on leftovers of the first
        b.reset()
        b.replay(0, n_frames)
        assert np.array_equal(b.read_yuv(), want)
    st = b.stats()
    assert st['frames'] == 2 * n_streams * n_frames and st['launches'] > 0
    b.close()


def test_submit_packed_path():
    name = 'pframes_256x192'
    w, h, ver, _ = CONFIGS[name]
    fr = frames(name, 77, 6)
    dec, par, ora = MobiclipDecoder(w, h, ver), MobiParser(w, h, ver), Oracle(w, h, ver)
    for data, key in fr:
        rc, off, pf = par.parse(data, 0)
        assert rc == 0
        dec.SubmitPacked(pf)
        assert ora.decode(data, 0, False)[0]
        assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv)
    dec.close()


def test_error_behaviour_matches_reference_null():
    w, h, ver, _ = CONFIGS['moflex_400x240']
    fr = frames('moflex_400x240', 5, 3)
    dec, ora = MobiclipDecoder(w, h, ver), Oracle(w, h, ver)
    # a P-frame before any picture exists: Y[1] == null -> exception -> null bitmap (MD:413, 325)
    dec.Data, dec.Offset = fr[1][0], 0
    assert dec.DecodeFrame() is None and dec.last_status == -4
    assert not ora.decode(fr[1][0], 0, False)[0]
    # truncated payload
    dec.Data, dec.Offset = fr[0][0][:40], 0
    assert dec.DecodeFrame() is None
    # the decoder is still usable and exact afterwards
    ora = Oracle(w, h, ver)
    for data, _ in fr:
        dec.Data, dec.Offset = data, 0
        assert dec.DecodeFrame() is not None
        assert ora.decode(data, 0, False)[0]
    assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv)
    dec.close()


def test_config2_full_size_1024_preparsed_p_pictures():
    """BASELINE config 2 at full size: 1024 independent (reference picture, P-picture) pairs at 256x192 from seeds
    1..1024, parsed and uploaded once (mobi_batch_stage), reconstructed by the kernels alone (mobi_batch_replay), every
    plane of every stream compared with the oracle."""
    name, n_streams = 'pframes_256x192', 1024
    w, h, ver, _ = CONFIGS[name]
    streams = [frames(name, 1 + s, 2) for s in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams)
    for f in range(2):
        b.stage([streams[s][f][0] for s in range(n_streams)])
    b.reset()
    b.replay(0, 2)
    got = b.read_yuv()
    for s in range(n_streams):
        o = Oracle(w, h, ver)
        assert o.decode(streams[s][0][0], 0, False)[0] and o.decode(streams[s][1][0], 0, False)[0]
        assert np.array_equal(got[s], o.i420()), 'stream %d (seed %d)' % (s, 1 + s)
    st = b.stats()
    assert st['frames'] == 2 * n_streams and st['inter_mbs'] == n_streams * (w // 16) * (h // 16)
    b.close()


@pytest.mark.parametrize('w,h,ver', [(16, 16, 2), (48, 32, 1), (256, 16, 1), (1024, 64, 2), (512, 48, 2), (272, 32, 1)])
def test_edge_geometries(w, h, ver):
    """Smallest picture, one-macroblock-wide / -high pictures, Width == Stride for all three strides (row wrap of the
    reference's flat addressing), and the first width past a stride boundary."""
    from mobiclipdecoder_b200 import SynthParams, SynthStream
    s = SynthStream(SynthParams(w, h, ver, 99, gop=5, p_intra_mb=0.2, p_split=0.5, p_oob_mv=0.3, p_cbp=0.5))
    dec, ora = MobiclipDecoder(w, h, ver), Oracle(w, h, ver)
    for i in range(12):
        data, key = s.next_frame()
        dec.Data, dec.Offset = data, 0
        bmp = dec.DecodeFrame()
        ok, off, want = ora.decode(data, 0)
        assert ok and bmp is not None, dec.last_error()
        assert dec.Offset == off
        assert np.array_equal(dec.Y[0], ora.y), 'frame %d luma: %s' % (i, _diff_report(dec.Y[0], ora.y, dec.Stride))
        assert np.array_equal(dec.UV[0], ora.uv), 'frame %d chroma: %s' % (i, _diff_report(dec.UV[0], ora.uv, dec.Stride))
        assert np.array_equal(bmp, want)
    dec.close()


def test_pipelined_submit_fetch_matches_oracle():
    name, n_streams, n_frames = 'moflex_400x240', 8, 10
    w, h, ver, _ = CONFIGS[name]
    streams = [frames(name, 300 + s, n_frames, gop=4) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    want = []
    for f in range(n_frames):
        row = []
        for s in range(n_streams):
            ok, _, bg = oracles[s].decode(streams[s][f][0], 0, True)
            assert ok
            row.append((oracles[s].i420(), bg))
        want.append(row)
    for fmt in (MobiBatch.OUT_I420, MobiBatch.OUT_BGRA):
        b = MobiBatch(w, h, ver, n_streams, n_threads=3)
        packed = [b.pack_inputs([streams[s][f][0] for s in range(n_streams)]) for f in range(n_frames)]
        got = []
        b.submit(packed[0], fmt=fmt)
        for f in range(1, n_frames):
            b.submit(packed[f], fmt=fmt)
            got.append(b.fetch())
        got.append(b.fetch())
        for f in range(n_frames):
            for s in range(n_streams):
                ref = want[f][s][0] if fmt == MobiBatch.OUT_I420 else want[f][s][1].ravel()
                assert np.array_equal(got[f][s], ref), 'fmt %d frame %d stream %d' % (fmt, f, s)
        with pytest.raises(Exception):
            b.fetch()   # nothing outstanding
        b.close()


def test_batch_stream_failure_is_isolated():
    """A stream whose frame cannot be parsed sits the step out (status < 0, its picture unchanged) ; the other
    streams advance bit-exactly.  (The failed stream itself is out of contract from then on: the
    reference leaves a half-written picture in its ring, MD:325, this library leaves the ring untouched.)"""
    name, n_streams, n_frames = 'moflex_400x240', 4, 9
    w, h, ver, _ = CONFIGS[name]
    streams = [frames(name, 500 + s, n_frames, gop=4) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams, n_threads=2)
    before = None
    for f in range(n_frames):
        batch_in = [streams[s][f][0] for s in range(n_streams)]
        if f == 2:
            batch_in[1] = batch_in[1][:1]           # not even one 16-bit word: ReadU16LE throws in the reference (IO:39)
            before = b.read_yuv()[1].copy()
        offs, status = b.decode(batch_in)
        assert status[0] == status[2] == status[3] == 0
        if f < 2:
            assert status[1] == 0
        if f == 2:
            assert status[1] < 0
        # afterwards stream 1 may fail again (its ring is one picture short of what the stream references) or succeed
        got = b.read_yuv()
        if f == 2:
            assert np.array_equal(got[1], before)   # newest picture of the failed stream is still the previous one
        for s in (0, 2, 3):
            assert oracles[s].decode(streams[s][f][0], 0, False)[0]
            assert np.array_equal(got[s], oracles[s].i420()), 'stream %d frame %d' % (s, f)
    b.close()


def test_config3_full_length_stream():
    """BASELINE config 3 at full length: one 400x240 Moflex3DS stream, 1024 frames, I-picture every 90; every picture's planes
    against the oracle (long P-chains: an error anywhere propagates, so this is also the drift check)."""
    name, n_frames = 'moflex_400x240', 1024
    w, h, ver, _ = CONFIGS[name]
    s = make_stream(name, 3)
    dec, ora = MobiclipDecoder(w, h, ver), Oracle(w, h, ver)
    for i in range(n_frames):
        data, key = s.next_frame()
        assert key == (i % 90 == 0)
        dec.Data, dec.Offset = data, 0
        assert dec.DecodeFrame(want_bitmap=False) is not None, 'frame %d: %s' % (i, dec.last_error())
        ok, off, _ = ora.decode(data, 0, False)
        assert ok and off == dec.Offset
        if i % 8 == 0 or key or i == n_frames - 1:     # planes every 8th picture, every key picture and the last one
            assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv), 'frame %d' % i
    dec.close()


def test_config5_eight_streams_lockstep():
    """BASELINE config 5 per GPU share: independent 400x240 streams with seeds 1..8 advancing in lock step."""
    name, n_streams, n_frames = 'moflex_400x240', 8, 100
    w, h, ver, _ = CONFIGS[name]
    gens = [make_stream(name, 1 + s) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams, n_threads=4)
    for f in range(n_frames):
        batch_in = [g.next_frame()[0] for g in gens]
        offs, status = b.decode(batch_in)
        assert all(st == 0 for st in status)
        for s in range(n_streams):
            ok, off, _ = oracles[s].decode(batch_in[s], 0, False)
            assert ok and off == offs[s]
        if f % 10 == 9 or f == 90:
            got = b.read_yuv()
            for s in range(n_streams):
                assert np.array_equal(got[s], oracles[s].i420()), 'stream %d frame %d' % (s, f)
    b.close()


def test_submit_packed_leaf_order_is_free():
    """A macroblock split once carries two leaf records; the parser emits top/left first, but the packed-array contract
    (include/mobicuda.h) does not fix the order.  Swapping the two records of every such macroblock must not change a
    single pixel (the inter kernel fetches both leaves as boxes and picks by position, not by record index)."""
    import ctypes as C
    name = 'moflex_400x240'
    w, h, ver, _ = CONFIGS[name]
    fr = frames(name, 31, 5)
    dec, par, ora = MobiclipDecoder(w, h, ver), MobiParser(w, h, ver), Oracle(w, h, ver)
    swapped = 0
    for data, key in fr:
        rc, off, pf = par.parse(data, 0)
        assert rc == 0
        hdr = pf.hdr.contents
        for m in range(hdr.n_mb):
            mb = pf.mbs[m]
            if (mb.info & 3) == 0 and ((mb.info >> 2) & 127) == 2:
                a, b = pf.parts[mb.first_sub], pf.parts[mb.first_sub + 1]
                ta = (a.xy, a.shape, a.mvx, a.mvy)
                a.xy, a.shape, a.mvx, a.mvy = b.xy, b.shape, b.mvx, b.mvy
                b.xy, b.shape, b.mvx, b.mvy = ta
                swapped += 1
        dec.SubmitPacked(pf)
        assert ora.decode(data, 0, False)[0]
        assert np.array_equal(dec.Y[0], ora.y), _diff_report(dec.Y[0].ravel(), ora.y.ravel(), 512)
        assert np.array_equal(dec.UV[0], ora.uv)
    assert swapped > 50
    dec.close()


def test_inter_kernel_variants_agree():
    """The three formulations of the inter path (mobi_kernels.cu, launch_inter): the default fused kernel, k_mc + k_res
    (MOBI_INTER_KERNEL=split) and k_inter_v3 (=v3, also with 8-macroblock chunks).  The choice is read once per process,
    so each variant runs in a child process and the digests of every decoded plane are compared."""
    import hashlib
    import os
    import subprocess
    import sys
    code = (
        "import hashlib, sys\n"
        "from mobiclipdecoder_b200 import MobiclipDecoder\n"
        "from mobiclipdecoder_b200.workloads import CONFIGS, frames\n"
        "w, h, ver, _ = CONFIGS['moc5_640x480']\n"
        "dec = MobiclipDecoder(w, h, ver)\n"
        "d = hashlib.sha256()\n"
        "for data, key in frames('moc5_640x480', 9, 12):\n"
        "    dec.Data, dec.Offset = data, 0\n"
        "    assert dec.DecodeFrame(False) is not None\n"
        "    d.update(dec.Y[0].tobytes()); d.update(dec.UV[0].tobytes())\n"
        "print(d.hexdigest())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for kern, envs in (('chunk', {}), ('split', {'MOBI_INTER_KERNEL': 'split'}), ('v3', {'MOBI_INTER_KERNEL': 'v3'}), ('v3_8', {'MOBI_INTER_KERNEL': 'v3', 'MOBI_INTER_CHUNK': '8'})):
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''), **envs)
        r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[kern] = r.stdout.strip().splitlines()[-1]
    assert out['split'] == out['v3'] == out['chunk'] == out['v3_8']


def test_bgra_arithmetic_gives_the_reference_bytes_for_every_input():
    """k_bgra evaluates a cheaper expression than the reference's float sequence (two contracted multiply-adds for R and B, a
    multiplication by RN(1/239) instead of the division for G); the library compares the two on the device for every possible
    (Y, U, V) -- 256 x 1021 x 1021 triples -- and must find no differing byte."""
    import ctypes as C
    from mobiclipdecoder_b200 import _native
    bad = (C.c_ulonglong * 2)(12345, 0)
    assert _native.mobicuda().mobicuda_selftest_bgra(0, bad) == 0
    assert bad[0] == 0, '%d triples differ, e.g. Y %d u %d v %d' % (bad[0], bad[1] >> 20, (bad[1] >> 10) & 1023, bad[1] & 1023)


@pytest.mark.parametrize('name,w,h', [('mods_256x192', 256, 192), ('moflex_400x240', 400, 240)])
def test_step_of_many_i_pictures_goes_by_ticket_and_is_bit_exact(name, w, h):
    """More I-pictures in one step than the GPU has SMs: their macroblocks join the depth-ordered ticket list of k_intra
    instead of one k_intra_key CTA per picture (mobi_runtime.cu pack_step).  160 streams, all on their I-picture in step 0
    (and again at the GOP boundary, together), P-pictures in between through the usual path."""
    _, _, ver, _ = CONFIGS[name]
    n_streams, n_frames = 160, 4
    gens = [make_stream(name, 7000 + s) for s in range(n_streams)]
    oracles = [Oracle(w, h, ver) for _ in range(n_streams)]
    b = MobiBatch(w, h, ver, n_streams, n_threads=4)
    for f in range(n_frames):
        batch_in = [g.next_frame()[0] for g in gens]
        offs, status = b.decode(batch_in)
        assert all(st == 0 for st in status)
        for s in range(n_streams):
            ok, off, _ = oracles[s].decode(batch_in[s], 0, False)
            assert ok and off == offs[s]
        if f in (0, n_frames - 1):
            got = b.read_yuv()
            for s in range(n_streams):
                assert np.array_equal(got[s], oracles[s].i420()), 'stream %d frame %d' % (s, f)
    b.close()
