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
