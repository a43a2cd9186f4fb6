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
