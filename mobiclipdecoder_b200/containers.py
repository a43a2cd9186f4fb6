"""Container framing in front of the decoder (include/mobidemux.h): Mods (DS) and MOC5 (Wii), mirroring
LibMobiclip.Containers.Mods.ModsDemuxer and the MOC5 loop of the reference player (Form1.cs:282-320)."""
import ctypes as C

import numpy as np

from . import _native as N
from .decoder import MobiError


class ModsDemuxer:
    """new ModsDemuxer(stream); .Header; .KeyFrames; .ReadFrame() -> (framedata, NrAudioPackets, IsKeyFrame) or None."""

    def __init__(self, data):
        self._lib = N.mobicuda()
        self._buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        rc = self._lib.mobi_mods_open(self._buf.ctypes.data_as(C.c_void_p), self._buf.size, C.byref(h))
        if rc != 0:
            raise MobiError(rc, 'not a Mods container')
        self._h = h
        self.Header = N.ModsHeader()
        self._lib.mobi_mods_get_header(self._h, C.byref(self.Header))
        self.KeyFrames = []
        for i in range(self.Header.keyframe_count):
            fn, off = C.c_uint32(), C.c_uint32()
            self._lib.mobi_mods_keyframe(self._h, i, C.byref(fn), C.byref(off))
            self.KeyFrames.append((fn.value, off.value))

    def ReadFrame(self):
        p, n, na, key = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self._lib.mobi_mods_read_frame(self._h, C.byref(p), C.byref(n), C.byref(na), C.byref(key))
        if rc == 0:
            return None
        if rc < 0:
            raise MobiError(rc, 'Mods packet runs past the end of the file')
        start = p.value - self._buf.ctypes.data
        return self._buf[start:start + n.value], na.value, bool(key.value)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mobi_mods_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MoLiveDemux:
    """LibMobiclip.Containers.Moflex.MoLiveDemux over mobi_moflex_*: ReadPacket() returns the reference's status code and
    fires OnCompleteFrameReceived(chunk, data) for every frame the packet completed (data ends with the two zero bytes the
    reference appends, MoLiveDemux.cs:353)."""

    def __init__(self, data):
        self._lib = N.mobicuda()
        self._buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        rc = self._lib.mobi_moflex_open(self._buf.ctypes.data_as(C.c_void_p), self._buf.size, C.byref(h))
        if rc != 0:
            raise MobiError(rc, 'mobi_moflex_open failed')
        self._h = h
        self.OnCompleteFrameReceived = None

    def ReadPacket(self):
        status = self._lib.mobi_moflex_read_packet(self._h)
        st, p, n = N.MoflexStream(), C.c_void_p(), C.c_uint32()
        while self._lib.mobi_moflex_next_frame(self._h, C.byref(st), C.byref(p), C.byref(n)):
            if self.OnCompleteFrameReceived is not None:
                chunk = N.MoflexStream.from_buffer_copy(st)
                self.OnCompleteFrameReceived(chunk, C.string_at(p, n.value))
        return status

    def frames(self):
        """Convenience: run ReadPacket() until it reports 73 (the CLI's loop, Program.cs:162-166) or cannot make progress
        (1: fewer than 14 bytes, 0x80: no synchro pattern); yields (chunk, data).  Other statuses (e.g. 0x50, a gap in the
        packet counter) are recoverable: the reference's caller just calls ReadPacket() again."""
        out = []
        prev, self.OnCompleteFrameReceived = self.OnCompleteFrameReceived, lambda c, d: out.append((c, d))
        try:
            for _ in range(1 << 24):
                st = self.ReadPacket()
                while out:
                    yield out.pop(0)
                if st in (73, 1, 0x80):
                    return
        finally:
            self.OnCompleteFrameReceived = prev

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mobi_moflex_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Moc5Reader:
    """Iterates (whole_file, decode_offset) pairs the way MOC5ThreadMain does (Form1.cs:291-318)."""

    def __init__(self, data):
        self._lib = N.mobicuda()
        self.data = np.frombuffer(data, dtype=np.uint8)
        self.info = N.Moc5Info()
        rc = self._lib.mobi_moc5_open(self.data.ctypes.data_as(C.c_void_p), self.data.size, C.byref(self.info))
        if rc != 0:
            raise MobiError(rc, 'not a MOC5 container')
        self.Width, self.Height = self.info.width, self.info.height

    def __iter__(self):
        cur, off, bs = C.c_uint32(self.info.first_block), C.c_uint32(), C.c_uint32()
        while True:
            rc = self._lib.mobi_moc5_next(self.data.ctypes.data_as(C.c_void_p), self.data.size, C.byref(cur), C.byref(off), C.byref(bs))
            if rc == 0:
                return
            if rc < 0:
                raise MobiError(rc, 'MOC5 block header outside the file')
            yield off.value, bs.value


class MoflexPlayer:
    """What the reference player does with the frames MoLiveDemux hands it (MobiclipDecoder/Form1.cs:508-545,
    d_OnCompleteFrameReceived): the first video stream seen becomes THE video stream; one decoder object decodes every one of
    its frames; a MoLiveStreamVideoWithLayout chunk whose ImageLayout is not Simple2D (6) marks the stream as stereoscopic --
    the two eyes' pictures alternate in the stream (and in the decoder's ring, so every frame must be decoded), the first,
    third, ... are shown and the frame period doubles (Form1.cs:516-528).  Audio chunks are not handled here.

    make_decoder(width, height) returns an object with Data / Offset / DecodeFrame() (mobiclipdecoder_b200.MobiclipDecoder,
    or a stand-in).  on_frame() returns None for chunks the player ignores, else a dict: bitmap (what DecodeFrame returned),
    present (shown or only decoded), period_ms (time until the next presented frame), eye ('left' / 'right' / None)."""
    SIMPLE_2D = 6   # MoLiveStreamVideoWithLayout.VideoLayout.Simple2D (MoLiveStreamVideoWithLayout.cs:10-19)

    def __init__(self, make_decoder):
        self._make = make_decoder
        self.decoder = None
        self.PlayingVideoStream = -1
        self.Is3D = False
        self.left = False

    def on_frame(self, chunk, data):
        if chunk.chunk_id not in (1, 3) or not (self.PlayingVideoStream == -1 or chunk.stream_index == self.PlayingVideoStream):
            return None
        self.left = not self.left
        if self.decoder is None:
            self.decoder = self._make(chunk.width, chunk.height)
            self.PlayingVideoStream = chunk.stream_index
            self.Is3D = chunk.chunk_id == 3 and chunk.image_layout != self.SIMPLE_2D
        self.decoder.Data, self.decoder.Offset = data, 0
        bitmap = self.decoder.DecodeFrame()
        present = (not self.Is3D) or self.left
        period = (2000.0 if self.Is3D else 1000.0) / (chunk.fps_rate / chunk.fps_scale)
        return {'bitmap': bitmap, 'present': present, 'period_ms': period, 'eye': None if not self.Is3D else ('left' if self.left else 'right')}

    def play(self, demux):
        """Drive a MoLiveDemux to its end (the CLI's stop conditions); yields on_frame()'s results for the video stream."""
        for chunk, data in demux.frames():
            r = self.on_frame(chunk, data)
            if r is not None:
                yield r
