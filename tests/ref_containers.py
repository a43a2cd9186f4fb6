"""TEST INFRASTRUCTURE: the reference's own container classes, compiled from /root/reference by oracle/build_ref.py
(ModsDemuxer.cs, MoLiveDemux.cs + chunk classes + MoLiveInBitStream.cs, MoflexMuxer.cs), behind oracle/ref_capi.cpp."""
import ctypes as C

from oracle_lib import REF_SO

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
        L.mobiref_mods_open.restype = vp
        L.mobiref_mods_open.argtypes = [C.c_char_p, C.c_size_t]
        L.mobiref_mods_close.argtypes = [vp]
        L.mobiref_mods_header.argtypes = [vp, u32p]
        L.mobiref_mods_keyframe.argtypes = [vp, C.c_uint32, u32p, u32p]
        L.mobiref_mods_read_frame.argtypes = [vp, C.POINTER(vp), u32p, u32p, C.POINTER(C.c_int)]
        L.mobiref_moflex_open.restype = vp
        L.mobiref_moflex_open.argtypes = [C.c_char_p, C.c_size_t]
        L.mobiref_moflex_close.argtypes = [vp]
        L.mobiref_moflex_read_packet.restype = C.c_uint32
        L.mobiref_moflex_read_packet.argtypes = [vp]
        L.mobiref_moflex_position.restype = C.c_longlong
        L.mobiref_moflex_position.argtypes = [vp]
        L.mobiref_moflex_next_frame.argtypes = [vp, u32p, C.POINTER(vp), u32p]
        L.mobiref_mux_create.restype = vp
        for n in ('destroy', 'synchro_header', 'end_chunks', 'data_block'):
            getattr(L, 'mobiref_mux_' + n).argtypes = [vp]
        L.mobiref_mux_video_chunk.argtypes = [vp] + [C.c_uint32] * 6 + [C.c_int, C.c_uint32]
        L.mobiref_mux_ep.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.mobiref_mux_pad.argtypes = [vp, C.c_int]
        L.mobiref_mux_bytes.restype = C.c_size_t
        L.mobiref_mux_bytes.argtypes = [vp, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


MODS_FIELDS = ['magic', 'tag_id', 'tag_id_size_dword', 'frame_count', 'width', 'height', 'fps', 'audio_codec', 'nb_channel', 'frequency',
               'biggest_frame', 'audio_offset', 'keyframe_index_offset', 'keyframe_count']
STREAM_FIELDS = ['stream_index', 'chunk_id', 'codec_id', 'fps_rate', 'fps_scale', 'width', 'height', 'pel_ratio_rate', 'pel_ratio_scale',
                 'image_layout', 'image_rotation', 'frequency', 'channels', 'associated_stream_index']


class RefMods:
    """new ModsDemuxer(stream) of the reference.  .ok is False when its constructor threw."""

    def __init__(self, data):
        self.L = lib()
        self.h = self.L.mobiref_mods_open(bytes(data), len(data))
        self.ok = bool(self.h)

    def header(self):
        v = (C.c_uint32 * 14)()
        self.L.mobiref_mods_header(self.h, v)
        return dict(zip(MODS_FIELDS, list(v)))

    def keyframes(self, n):
        out = []
        for i in range(n):
            a, b = C.c_uint32(), C.c_uint32()
            assert self.L.mobiref_mods_keyframe(self.h, i, C.byref(a), C.byref(b)) == 0
            out.append((a.value, b.value))
        return out

    def read_frame(self):
        """(bytes, NrAudioPackets, IsKeyFrame); None when ReadFrame returned null; 'threw' when it threw."""
        p, n, na, key = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self.L.mobiref_mods_read_frame(self.h, C.byref(p), C.byref(n), C.byref(na), C.byref(key))
        if rc == 0:
            return None
        if rc < 0:
            return 'threw'
        return C.string_at(p, n.value), na.value, bool(key.value)

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.mobiref_mods_close(self.h)


class RefMoLive:
    def __init__(self, data):
        self.L = lib()
        self.h = self.L.mobiref_moflex_open(bytes(data), len(data))

    def read_packet(self):
        """(status or 0xFFFFFFFF when the reference threw, [(chunk fields dict, data bytes), ...] delivered by this call)."""
        st = self.L.mobiref_moflex_read_packet(self.h)
        out = []
        v, p, n = (C.c_uint32 * 14)(), C.c_void_p(), C.c_uint32()
        while self.L.mobiref_moflex_next_frame(self.h, v, C.byref(p), C.byref(n)):
            out.append((dict(zip(STREAM_FIELDS, list(v))), C.string_at(p, n.value)))
        return st, out

    def position(self):
        return self.L.mobiref_moflex_position(self.h)

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.mobiref_moflex_close(self.h)


def ref_mux_simple_video(frames, width, height, fps_rate=24, fps_scale=1):
    """What MoflexSimpleVideoMuxer does around the encoder (MoflexSimpleVideoMuxer.cs:13-70), driving the reference's
    MoflexMuxer: synchro header, one MoLiveStreamVideo chunk, the end-of-chunks marker, then per frame data blocks of at most
    0x1000 - 0x80 bytes per end-point, and FinalizeMoflex's 0x1000 zero bytes."""
    L = lib()
    m = L.mobiref_mux_create()
    L.mobiref_mux_synchro_header(m)
    L.mobiref_mux_video_chunk(m, fps_rate, fps_scale, width, height, 1, 1, 0, 0)
    L.mobiref_mux_end_chunks(m)
    cap = 0x1000 - 0x80
    for data in frames:
        pos, left = 0, len(data)
        if left <= cap:
            L.mobiref_mux_data_block(m); L.mobiref_mux_ep(m, 0, data, left, 1); L.mobiref_mux_ep(m, 0, None, 0, 0)
            continue
        while left >= cap:
            L.mobiref_mux_data_block(m); L.mobiref_mux_ep(m, 0, data[pos:pos + cap], cap, 1 if left == cap else 0); L.mobiref_mux_ep(m, 0, None, 0, 0)
            pos += cap
            left -= cap
        if left > 0:
            L.mobiref_mux_data_block(m); L.mobiref_mux_ep(m, 0, data[pos:], left, 1); L.mobiref_mux_ep(m, 0, None, 0, 0)
    L.mobiref_mux_pad(m, 0x1000)
    n = L.mobiref_mux_bytes(m, None, 0)
    buf = C.create_string_buffer(n)
    L.mobiref_mux_bytes(m, buf, n)
    L.mobiref_mux_destroy(m)
    return buf.raw
