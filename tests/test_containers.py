"""Mods / MOC5 framing (include/mobidemux.h): the native readers against a restatement of the reference readers on
synthetic containers, and container -> demux -> decode end to end against the oracle."""
import numpy as np
import pytest

from container_ref import read_moc5_reference, read_mods_reference, write_moc5, write_mods
from mobiclipdecoder_b200.containers import Moc5Reader, ModsDemuxer
from mobiclipdecoder_b200.decoder import MobiError
from mobiclipdecoder_b200.workloads import CONFIGS, frames
from oracle_lib import Oracle


def _mods_file(n=40, seed=21):
    fr = frames('mods_256x192', seed, n, gop=8)
    rng = np.random.default_rng(seed)
    audio = [int(rng.integers(0, 5)) for _ in fr]
    # the demuxer hands out the packet as is; give every packet a tail standing in for the audio that follows the video
    packets = [(d[:-2] + bytes(rng.integers(0, 256, size=2 * a + 2, dtype=np.uint8)), k) for (d, k), a in zip(fr, audio)]
    return write_mods(packets, 256, 192, audio_packets=audio), packets, audio


def test_mods_reader_matches_reference_reader():
    blob, packets, audio = _mods_file()
    h, keys, want = read_mods_reference(blob)
    dm = ModsDemuxer(blob)
    for name in ('tag_id', 'frame_count', 'width', 'height', 'fps', 'audio_codec', 'nb_channel', 'frequency', 'biggest_frame',
                 'audio_offset', 'keyframe_index_offset', 'keyframe_count'):
        assert getattr(dm.Header, name) == h[name], name
    assert dm.Header.magic == b'MODS'
    assert dm.KeyFrames == [tuple(k) for k in keys] and len(keys) == 5
    got = []
    while True:
        r = dm.ReadFrame()
        if r is None:
            break
        got.append(r)
    assert len(got) == len(want) == len(packets)
    for (a, na, ka), (b, nb, kb), (p, k) in zip(got, want, packets):
        assert bytes(a) == b == p and na == nb and ka == kb
    # reference quirk kept: JumpToKeyFrame(0) arms NextKeyFrame = 1, so the very first key frame is not flagged (MODS:88-106)
    assert [k for _, _, k in got] == [i > 0 and i % 8 == 0 for i in range(len(got))]
    assert dm.ReadFrame() is None   # stays at end of stream


def test_mods_reader_rejects_truncation():
    blob, _, _ = _mods_file(10)
    with pytest.raises(MobiError):
        ModsDemuxer(blob[:0x20])
    dm = ModsDemuxer(blob[:len(blob) - 100])
    with pytest.raises(MobiError):
        while dm.ReadFrame() is not None:
            pass


def test_moc5_reader_matches_reference_reader():
    fr = frames('moc5_640x480', 3, 6)
    blob = write_moc5(fr, 640, 480)
    (w, h, fps), want = read_moc5_reference(blob)
    rd = Moc5Reader(blob)
    assert (rd.Width, rd.Height, rd.info.fps_x128) == (w, h, fps) == (640, 480, 30 * 128)
    got = list(rd)
    assert got == want and len(got) == len(fr)
    for (off, bs), (p, _) in zip(got, fr):
        assert bytes(rd.data[off:off + len(p)]) == p   # Offset points at the frame payload inside the whole file


def test_mods_container_decodes_like_the_oracle_on_cpu_side():
    """Parser + demuxer on CPU: Offset after each packet's video part is where the audio tail begins (Program.cs:250)."""
    from mobiclipdecoder_b200 import MobiParser
    blob, packets, audio = _mods_file(20)
    dm, o, p = ModsDemuxer(blob), Oracle(256, 192, 1), MobiParser(256, 192, 1)
    while True:
        r = dm.ReadFrame()
        if r is None:
            break
        data, na, key = r
        ok, off, _ = o.decode(bytes(data), 0, False)
        rc, off2, _ = p.parse(data, 0)
        assert ok and rc == 0 and off == off2


@pytest.mark.gpu
def test_containers_end_to_end_on_gpu():
    from mobiclipdecoder_b200 import MobiclipDecoder
    blob, _, _ = _mods_file(24)
    dm = ModsDemuxer(blob)
    dec, ora = MobiclipDecoder(dm.Header.width, dm.Header.height, 1), Oracle(256, 192, 1)
    while True:
        r = dm.ReadFrame()
        if r is None:
            break
        dec.Data, dec.Offset = r[0], 0
        assert dec.DecodeFrame() is not None
        ok, off, _ = ora.decode(bytes(r[0]), 0, False)
        assert ok and off == dec.Offset
        assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv)
    dec.close()
    fr = frames('moc5_640x480', 3, 8)
    rd = Moc5Reader(write_moc5(fr, 640, 480))
    dec, ora = MobiclipDecoder(rd.Width, rd.Height, 2), Oracle(640, 480, 2)
    for off, bs in rd:            # Data = the whole file, Offset = block + 8 (Form1.cs:292, 301)
        dec.Data, dec.Offset = rd.data, off
        assert dec.DecodeFrame() is not None
        ok, off2, _ = ora.decode(rd.data.tobytes(), off, False)
        assert ok and off2 == dec.Offset
        assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv)
    dec.close()


# ---- Moflex ---------------------------------------------------------------------------------------------------------
from container_ref import write_moflex
from mobiclipdecoder_b200.containers import MoLiveDemux


def _moflex_file(name='moflex_400x240', n=12, seed=31):
    w, h, ver, _ = CONFIGS[name]
    fr = [d[:-2] for d, _ in frames(name, seed, n)]          # the writer takes bare payloads; the demuxer re-appends two zero bytes
    return write_moflex(fr, w, h), fr, (w, h)


def test_moflex_signature_and_round_trip():
    blob, fr, (w, h) = _moflex_file()
    assert blob[:4] == bytes([0x4C, 0x32, 0xAA, 0xAB])       # what the CLI sniffs (MobiConverter/Program.cs:45)
    dm = MoLiveDemux(blob)
    got = list(dm.frames())
    assert len(got) == len(fr)
    for (chunk, data), want in zip(got, fr):
        assert data == want + b'\x00\x00'                     # MoLiveDemux.cs:353
        assert (chunk.chunk_id, chunk.stream_index, chunk.width, chunk.height, chunk.fps_rate, chunk.fps_scale) == (1, 0, w, h, 24, 1)
    assert dm.ReadPacket() == 73                              # the zero tail is shorter than a packet: the CLI's end condition


def test_moflex_frames_split_over_many_end_points():
    blob, fr, _ = _moflex_file('moc5_640x480', 4, 5)          # ~22 KB frames -> six end-points each
    assert max(len(f) for f in fr) > 3 * (0x1000 - 0x80)
    got = [d for _, d in MoLiveDemux(blob).frames()]
    assert got == [f + b'\x00\x00' for f in fr]


def test_moflex_status_codes_and_event_order():
    blob, fr, _ = _moflex_file(n=3)
    dm = MoLiveDemux(blob)
    seen = []
    dm.OnCompleteFrameReceived = lambda chunk, data: seen.append(len(data))
    assert dm.ReadPacket() == 0 and seen == []               # first call only synchronises (MoLiveDemux.cs:75-95)
    assert dm.ReadPacket() == 0 and seen == []               # second: header announces 0x1001-byte packets > the 0x1000 buffer: retry (:131-136)
    calls = 0
    while len(seen) < 3:
        assert dm.ReadPacket() == 0
        calls += 1
    # one data block per call; a frame longer than 0xF80 bytes spans several blocks
    assert calls == sum(-(-len(f) // (0x1000 - 0x80)) for f in fr)
    assert seen == [len(f) + 2 for f in fr]
    assert MoLiveDemux(b'\x00' * 64).ReadPacket() == 0x80     # no synchro pattern
    assert MoLiveDemux(b'L2').ReadPacket() == 1               # fewer than 14 bytes


def test_moflex_garbage_never_crashes():
    rng = np.random.default_rng(4)
    blob, _, _ = _moflex_file(n=4)
    for t in range(60):
        b = bytearray(blob)
        for _ in range(8):
            b[int(rng.integers(14, len(b) - 0x1000))] = int(rng.integers(0, 256))
        dm = MoLiveDemux(bytes(b))
        for _ in range(64):
            if dm.ReadPacket() == 73:
                break


@pytest.mark.gpu
def test_moflex_container_end_to_end_on_gpu():
    from mobiclipdecoder_b200 import MobiclipDecoder
    blob, fr, (w, h) = _moflex_file(n=16)
    dec, ora, n = None, Oracle(w, h, 2), 0
    for chunk, data in MoLiveDemux(blob).frames():
        if dec is None:
            dec = MobiclipDecoder(chunk.width, chunk.height, 2)   # Program.cs:64-67
        dec.Data, dec.Offset = data, 0
        assert dec.DecodeFrame() is not None
        ok, off, _ = ora.decode(data, 0, False)
        assert ok and off == dec.Offset
        assert np.array_equal(dec.Y[0], ora.y) and np.array_equal(dec.UV[0], ora.uv)
        n += 1
    assert n == 16
    dec.close()


def test_moflex_two_streams_fixed_packets_and_counting():
    """Hand-built packets: a video and an audio stream interleaved over end-points 0 and 1, fixed-size packets
    (flag bit 0 clear: the rest of the packet after the terminator is padding, MoLiveDemux.cs:286-296), packet counting
    (flag bit 1, :245-266) and a stream table re-announced by a second synchro header (:124-129)."""
    import struct
    from container_ref import _ep, _synchro_header, _variable_byte, _video_chunk
    PS = 0x200                                           # packet size; the header field holds size - 1
    audio = _variable_byte(2) + _variable_byte(6) + struct.pack('>BB', 1, 0) + (32000 - 1).to_bytes(3, 'big') + bytes([2 - 1])
    table = _video_chunk(0, 0, 30, 1, 64, 48) + audio + _variable_byte(0) + _variable_byte(0)
    rng = np.random.default_rng(9)
    v = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in (300, 120)]
    a = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in (40, 64)]

    def packet(body, counter, header=None):
        p = (header or b'') + bytes([0 | 2 | 0 << 2]) + struct.pack('>H', counter) + body + b'\x00'
        assert len(p) <= PS
        return p.ljust(PS, b'\xEE')                      # padding is skipped, never parsed
    hdr = _synchro_header(ts=1, packet_size_field=PS - 1) + table
    blob = packet(_ep(0, v[0][:200], False) + _ep(1, a[0], True), 7, hdr)
    blob += packet(_ep(0, v[0][200:], True), 8)
    blob += packet(_ep(1, a[1], True) + _ep(0, v[1], True), 9, _synchro_header(ts=5, packet_size_field=PS - 1) + table)
    blob += bytes(PS - 1)                                 # a short tail: status 73
    dm = MoLiveDemux(blob)
    got = [(c.chunk_id, c.stream_index, c.width, c.frequency, c.channels, d) for c, d in dm.frames()]
    assert got == [(2, 1, 0, 32000, 2, a[0] + b'\0\0'), (1, 0, 64, 0, 0, v[0] + b'\0\0'),
                   (2, 1, 0, 32000, 2, a[1] + b'\0\0'), (1, 0, 64, 0, 0, v[1] + b'\0\0')]
    # packet counting: LastCounter starts at 0 (field default, MoLiveDemux.cs:27), so the first counted packet (7, expected 1)
    # is reported as a gap (0x50) without being consumed and accepted on the retry; same for a real gap (8 -> 11)
    bad = packet(_ep(0, v[0][:200], False), 7, hdr) + packet(_ep(0, v[0][200:], True), 11) + bytes(PS - 1)
    dm = MoLiveDemux(bad)
    assert [dm.ReadPacket() for _ in range(6)] == [0, 0x50, 0, 0x50, 0, 73]


def _layout_file(layout, n=8, second_video=True):
    """A Moflex file whose video stream is announced by a MoLiveStreamVideoWithLayout chunk (id 3, 13 bytes: the video
    chunk's 12 + layout | rotation << 4), plus an audio stream and a SECOND video stream the player must ignore."""
    import struct
    from container_ref import _ep, _synchro_header, _variable_byte
    w, h, ver, _ = CONFIGS['moflex_400x240']
    fr = [d[:-2] for d, _ in frames('moflex_400x240', 77, n)]
    lay = _variable_byte(3) + _variable_byte(13) + struct.pack('>BBHHHHBB', 0, 0, 30, 1, w, h, 1, 1) + bytes([layout & 15])
    audio = _variable_byte(2) + _variable_byte(6) + struct.pack('>BB', 1, 0) + (32000 - 1).to_bytes(3, 'big') + bytes([1])
    other = b''
    if second_video:
        other = _variable_byte(1) + _variable_byte(12) + struct.pack('>BBHHHHBB', 1, 0, 30, 1, 64, 48, 1, 1)
        audio = b''     # end-points 0 and 1 only (MoflexMuxer.WriteEp's byte count is right for those)
    out = bytearray(_synchro_header() + lay + audio + other + _variable_byte(0) + _variable_byte(0))
    cap = 0xE00
    for i, f in enumerate(fr):
        for at in range(0, len(f), cap):          # one data block per slice; the other stream's frame rides in some of the last ones
            last = at + cap >= len(f)
            out += b'\x01' + _ep(0, f[at:at + cap], last) + (_ep(1, b'\x55' * 10, True) if i % 3 == 1 and last else b'') + _ep(0, None, False)
    out += bytes(0x1000)
    return bytes(out), fr, (w, h)


class _OracleDecoder:
    """Stand-in with the decoder object's surface, backed by the CPU oracle: lets the player's policy be tested without a GPU."""

    def __init__(self, w, h):
        self.o, self.Data, self.Offset = Oracle(w, h, 2), None, 0

    def DecodeFrame(self):
        ok, self.Offset, bgra = self.o.decode(self.Data, self.Offset, True)
        return bgra if ok else None


@pytest.mark.parametrize('layout,is3d', [(0, True), (3, True), (4, True), (6, False)])
def test_moflex_player_acts_on_image_layout_like_the_reference_player(layout, is3d):
    """Form1.cs:508-545: every frame of the chosen stream is decoded by ONE decoder; a 3-D layout shows the 1st, 3rd, ... and
    doubles the frame period; Simple2D (6) shows all; the second video stream and the audio stream are ignored."""
    from mobiclipdecoder_b200.containers import MoflexPlayer
    blob, fr, (w, h) = _layout_file(layout)
    made = []
    pl = MoflexPlayer(lambda ww, hh: made.append((ww, hh)) or _OracleDecoder(ww, hh))
    got = list(pl.play(MoLiveDemux(blob)))
    assert made == [(w, h)] and pl.PlayingVideoStream == 0 and pl.Is3D == is3d
    assert len(got) == len(fr) and all(g['bitmap'] is not None for g in got)       # the other stream's frames never reach the decoder
    ref = _OracleDecoder(w, h)
    for g, f in zip(got, fr):
        ref.Data, ref.Offset = f + b'\0\0', 0
        assert np.array_equal(g['bitmap'], ref.DecodeFrame())
    if is3d:
        assert [g['present'] for g in got] == [i % 2 == 0 for i in range(len(fr))]
        assert [g['eye'] for g in got[:4]] == ['left', 'right', 'left', 'right']
        assert all(g['period_ms'] == pytest.approx(2000.0 / 30) for g in got)
    else:
        assert all(g['present'] for g in got) and all(g['period_ms'] == pytest.approx(1000.0 / 30) for g in got)


def test_moflex_player_plain_video_chunk_is_2d():
    from mobiclipdecoder_b200.containers import MoflexPlayer
    blob, fr, (w, h) = _moflex_file(n=5)
    pl = MoflexPlayer(_OracleDecoder)
    got = list(pl.play(MoLiveDemux(blob)))
    assert not pl.Is3D and len(got) == 5 and all(g['present'] and g['eye'] is None for g in got)


@pytest.mark.gpu
def test_moflex_player_3d_on_gpu():
    from mobiclipdecoder_b200 import MobiclipDecoder
    from mobiclipdecoder_b200.containers import MoflexPlayer
    blob, fr, (w, h) = _layout_file(4, n=10)
    pl = MoflexPlayer(lambda ww, hh: MobiclipDecoder(ww, hh, 2))
    ref = _OracleDecoder(w, h)
    shown = 0
    for g, f in zip(pl.play(MoLiveDemux(blob)), fr):
        ref.Data, ref.Offset = f + b'\0\0', 0
        assert np.array_equal(np.asarray(g['bitmap']).reshape(h, w, 4), ref.DecodeFrame())
        shown += g['present']
    assert shown == 5 and pl.Is3D
    pl.decoder.close()
