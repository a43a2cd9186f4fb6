"""The native demuxers (include/mobidemux.h) against the reference's OWN container classes compiled from /root/reference
(ModsDemuxer.cs:16-117, MoLiveDemux.cs:67-414 with its chunk classes and MoLiveInBitStream, MoflexMuxer.cs as the writer):
same files in, same headers / key-frame tables / frames / per-call status codes out."""
import struct

import numpy as np
import pytest

from container_ref import _ep, _synchro_header, _variable_byte, _video_chunk, write_mods, write_moflex
from mobiclipdecoder_b200 import _native as N
from mobiclipdecoder_b200.containers import ModsDemuxer, MoLiveDemux
from mobiclipdecoder_b200.decoder import MobiError
from mobiclipdecoder_b200.workloads import CONFIGS, frames
from oracle_lib import have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref not built (no /root/reference on this box and no prebuilt copy)')


def _mods(n, seed, gop, with_audio_table=False, first_key=0):
    from ref_containers import MODS_FIELDS  # noqa: F401
    fr = frames('mods_256x192', seed, n, gop=gop)
    rng = np.random.default_rng(seed)
    audio = [int(rng.integers(0, 9)) for _ in fr]
    packets = [(d[:-2] + bytes(rng.integers(0, 256, size=2 * a + 2, dtype=np.uint8)), k) for (d, k), a in zip(fr, audio)]
    return write_mods(packets, 256, 192, audio_packets=audio, fps=int(rng.integers(1, 1 << 31)), tag_id=int(rng.integers(0, 1 << 16)))


@pytest.mark.parametrize('n,seed,gop', [(40, 21, 8), (9, 3, 4), (1, 5, 1), (33, 8, 33)])
def test_mods_demuxer_matches_the_compiled_reference(n, seed, gop):
    from ref_containers import RefMods
    blob = _mods(n, seed, gop)
    ref, dm = RefMods(blob), ModsDemuxer(blob)
    assert ref.ok
    h = ref.header()
    assert h['magic'] == int.from_bytes(bytes(dm.Header.magic), 'little')
    for name in h:
        if name != 'magic':
            assert getattr(dm.Header, name) == h[name], name
    assert dm.KeyFrames == ref.keyframes(h['keyframe_count'])
    count = 0
    while True:
        a, b = ref.read_frame(), dm.ReadFrame()
        if a is None or b is None:
            assert a is None and b is None
            break
        assert bytes(b[0]) == a[0] and b[1] == a[1] and b[2] == a[2], count
        count += 1
    assert count == n
    assert ref.read_frame() is None and dm.ReadFrame() is None


def test_mods_header_fields_and_key_frame_quirks_match_the_reference():
    """Hand-made header: every field distinct, an audio codebook section, a key-frame table that does not start at frame 0
    (JumpToKeyFrame(0) then starts CurFrame at that frame number and the reader hands out FrameCount - first frames)."""
    from ref_containers import RefMods
    rng = np.random.default_rng(12)
    payloads = [bytes(rng.integers(0, 256, size=int(rng.integers(4, 60)) * 2, dtype=np.uint8)) for _ in range(12)]
    keys = [3, 7, 10]
    nb_channel, table_off = 2, 0x30 + 2 * 0xC34
    data_off = table_off + 8 * len(keys)
    offs, at, body = {}, data_off, b''
    for i, p in enumerate(payloads[3:], start=3):
        offs[i] = at
        body += struct.pack('<I', (len(p) << 14) | (i & 0x3FFF)) + p
        at += 4 + len(p)
    hdr = b'MODS' + struct.pack('<HHIIIIHHIIIII', 0x1234, 0x5678, len(payloads), 256, 192, 0x1E000000, 3, nb_channel, 32768, 999, 0x30, table_off, len(keys))
    blob = hdr + bytes(rng.integers(0, 256, size=2 * 0xC34, dtype=np.uint8)) + b''.join(struct.pack('<II', k, offs[k]) for k in keys) + body
    ref, dm = RefMods(blob), ModsDemuxer(blob)
    assert ref.ok
    h = ref.header()
    for name in h:
        if name != 'magic':
            assert getattr(dm.Header, name) == h[name], name
    assert dm.KeyFrames == ref.keyframes(3) == [(k, offs[k]) for k in keys]
    got_ref, got = [], []
    while True:
        a, b = ref.read_frame(), dm.ReadFrame()
        if a is None:
            assert b is None
            break
        got_ref.append(a)
        got.append((bytes(b[0]), b[1], b[2]))
    assert got == got_ref and len(got) == 9 and [g[2] for g in got] == [False, False, False, False, True, False, False, True, False]


def _both_moflex(blob, max_calls=4096):
    """Run ReadPacket() on both demuxers until the reference reports one of the CLI's stop conditions; every call must return
    the same status and deliver the same frames (chunk fields and bytes) in the same order."""
    from ref_containers import RefMoLive, STREAM_FIELDS
    ref, dm = RefMoLive(blob), MoLiveDemux(blob)
    mine = []
    dm.OnCompleteFrameReceived = lambda c, d: mine.append(({f: getattr(c, f) & 0xFFFFFFFF for f in STREAM_FIELDS}, d))
    statuses, delivered = [], []
    for call in range(max_calls):
        st_ref, fr_ref = ref.read_packet()
        del mine[:]
        st = dm.ReadPacket()
        assert st_ref != 0xFFFFFFFF, 'the reference threw on call %d: not a file this test should feed it' % call
        assert st == st_ref, 'call %d: status %#x, reference %#x' % (call, st, st_ref)
        assert len(mine) == len(fr_ref), call
        for (ca, da), (cb, db) in zip(mine, fr_ref):
            assert da == db, call
            assert ca == cb, (call, ca, cb)
        statuses.append(st)
        delivered += fr_ref
        if st_ref in (73, 1, 0x80):
            break
    return statuses, delivered


@pytest.mark.parametrize('name,n,seed', [('moflex_400x240', 12, 31), ('moc5_640x480', 4, 5), ('mods_256x192', 6, 2)])
def test_moflex_demuxer_matches_the_compiled_reference_on_the_references_own_muxer_output(name, n, seed):
    from ref_containers import ref_mux_simple_video
    w, h, _, _ = CONFIGS[name]
    fr = [d[:-2] for d, _ in frames(name, seed, n)]
    blob = ref_mux_simple_video(fr, w, h)
    assert blob == write_moflex(fr, w, h)      # the Python writer the other container tests use is byte-identical to MoflexMuxer
    statuses, delivered = _both_moflex(blob)
    assert statuses[-1] == 73
    assert [d for _, d in delivered] == [f + b'\0\0' for f in fr]
    assert all((c['chunk_id'], c['width'], c['height'], c['fps_rate'], c['fps_scale']) == (1, w, h, 24, 1) for c, _ in delivered)


def _ep_any(ep, data, end_frame):
    """An end-point header for any stream index, laid out the way ReadEp reads it (MoLiveDemux.cs:299-324): unary length of
    the index, the index, EndFrame, [frame type, sign, timestamp length, 28-bit timestamp], 13-bit size - 1, padded to a byte.
    (MoflexMuxer.WriteEp's byte count is only right for end-points 0 and 1.)"""
    nrbits = max(1, ep.bit_length())
    bits = '0' * (nrbits - 1) + '1' + format(ep, '0%db' % nrbits) + ('1' if end_frame else '0')
    if end_frame:
        bits += '1' + '0' + '0' + '1' + '0' * 28
    bits += format(len(data) - 1, '013b')
    nbytes = (len(bits) + 7) // 8
    return int(bits.ljust(nbytes * 8, '0'), 2).to_bytes(nbytes, 'big') + data


def test_moflex_two_streams_layout_chunk_timeline_and_counting_match_the_reference():
    """Hand-built packets: a MoLiveStreamVideoWithLayout, an audio and a timeline stream over end-points 0-2, fixed-size
    packets, packet counting with a gap (status 0x50), a re-announced stream table, a synchro-counter change."""
    PS = 0x200
    layout = _variable_byte(3) + _variable_byte(13) + struct.pack('>BBHHHHBB', 0, 7, 30, 1, 400, 240, 5, 9) + bytes([0x24])   # layout 4 (side by side), rotation 2
    audio = _variable_byte(2) + _variable_byte(6) + struct.pack('>BB', 1, 4) + (32000 - 1).to_bytes(3, 'big') + bytes([2 - 1])
    timeline = _variable_byte(4) + _variable_byte(2) + bytes([2, 0])
    table = layout + audio + timeline + _variable_byte(0) + _variable_byte(0)
    rng = np.random.default_rng(19)
    blobs = lambda n: bytes(rng.integers(0, 256, size=n, dtype=np.uint8))   # noqa: E731
    v, a, t = [blobs(n) for n in (300, 120, 77)], [blobs(n) for n in (40, 64, 8)], [blobs(5)]

    def packet(body, counter, header=None, synchro=0):
        p = (header or b'') + bytes([0 | 2 | synchro << 2]) + struct.pack('>H', counter) + body + b'\x00'
        assert len(p) <= PS
        return p.ljust(PS, b'\xEE')
    hdr = lambda ts: _synchro_header(ts=ts, packet_size_field=PS - 1) + table   # noqa: E731
    blob = packet(_ep(0, v[0][:200], False) + _ep(1, a[0], True), 7, hdr(1))
    blob += packet(_ep(0, v[0][200:], True) + _ep_any(2, t[0], True), 8)
    blob += packet(_ep(1, a[1], True) + _ep(0, v[1], True), 9, hdr(5))
    blob += packet(_ep(0, v[2][:50], False), 11)                           # counter gap: 0x50, the packet is not consumed
    blob += packet(_ep(0, v[2][50:], True) + _ep(1, a[2], True), 12, None, synchro=3)   # synchro counter moved: partial frames are dropped
    blob += bytes(PS - 1)
    statuses, delivered = _both_moflex(blob)
    assert 0x50 in statuses
    kinds = [(c['chunk_id'], c['stream_index']) for c, _ in delivered]
    assert (3, 0) in kinds and (2, 1) in kinds and (4, 2) in kinds
    lay = next(c for c, _ in delivered if c['chunk_id'] == 3)
    assert (lay['image_layout'], lay['image_rotation'], lay['codec_id'], lay['width'], lay['height']) == (4, 2, 7, 400, 240)
    # MoLiveStreamVideoWithLayout.Read assigns byte 9 to PelRatioRate as well and never sets PelRatioScale (MoLiveStreamVideoWithLayout.cs:36-37)
    assert (lay['pel_ratio_rate'], lay['pel_ratio_scale']) == (9, 0)


def test_moflex_damaged_files_give_the_references_status_sequence():
    """Bytes flipped in end-point headers and data-block flags (not in the 14-byte synchro header, whose loss makes the
    reference index outside its packet buffer): both demuxers walk the same status sequence and deliver the same frames.
    A case where the reference throws is skipped (include/mobidemux.h documents what the native reader returns there)."""
    from ref_containers import RefMoLive, STREAM_FIELDS
    w, h, _, _ = CONFIGS['moflex_400x240']
    fr = [d[:-2] for d, _ in frames('moflex_400x240', 44, 6)]
    blob = write_moflex(fr, w, h)
    rng = np.random.default_rng(6)
    compared = 0
    for trial in range(40):
        b = bytearray(blob)
        for _ in range(3):
            b[int(rng.integers(0x30, len(b) - 0x1000))] = int(rng.integers(0, 256))
        b = bytes(b)
        ref, dm = RefMoLive(b), MoLiveDemux(b)
        mine = []
        dm.OnCompleteFrameReceived = lambda c, d: mine.append(d)
        ok = True
        for call in range(200):
            st_ref, fr_ref = ref.read_packet()
            if st_ref == 0xFFFFFFFF:
                ok = False
                break
            del mine[:]
            st = dm.ReadPacket()
            assert st == st_ref, (trial, call, hex(st), hex(st_ref))
            assert mine == [d for _, d in fr_ref], (trial, call)
            if st_ref in (73, 1, 0x80):
                break
        compared += ok
    assert compared >= 20
