"""Host-side mirror of the reference decoder object over libmobicuda.so.

Reference surface being mirrored (LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs, "MD"):
    new MobiclipDecoder(uint Width, uint Height, MobiclipVersion Version)   MD:41-54
    d.Data = byte[]; d.Offset = int; Bitmap b = d.DecodeFrame();           MD:15-16, 56
    d.Y[0], d.UV[0], d.Stride, d.Quantizer, d.YuvFormat, d.Width, d.Height  MD:17-30
Same member names, same argument meaning, same error behaviour (DecodeFrame returns None where the reference
returns null, MD:325-328).  All pixel work happens in the CUDA library; nothing here touches pixels.
"""
import ctypes as C
import enum

import numpy as np

from . import _native as N


class MobiclipVersion(enum.IntEnum):  # MD:32-37
    VxDS = 0
    ModsDS = 1
    Moflex3DS = 2


class MobiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('mobicuda error %d: %s' % (code, msg))
        self.code = code


def _stride_for(width):  # MD:50-52
    return 256 if width <= 256 else 512 if width <= 512 else 1024


def _as_u8(buf):
    """A ctypes view of caller memory without copying (bytes, bytearray, numpy)."""
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags['C_CONTIGUOUS']:
            raise TypeError('frame data must be a contiguous uint8 array')
        return buf.ctypes.data_as(C.c_void_p), buf.size
    if isinstance(buf, (bytes, bytearray, memoryview)):
        a = np.frombuffer(buf, dtype=np.uint8)
        return a.ctypes.data_as(C.c_void_p), a.size
    raise TypeError('frame data must be bytes-like')


class _Ring:
    """d.Y[0] / d.UV[0]: index 0 reads the newest picture back from the GPU (strided, byte-identical to the
    reference's arrays).  Older ring entries stay on the device; the reference exposes them only as a side effect."""

    def __init__(self, dec, chroma):
        self._dec = dec
        self._chroma = chroma

    def __getitem__(self, i):
        if i != 0:
            raise IndexError('only the newest picture ([0]) is readable through the boundary')
        y, uv = self._dec._read_strided()
        return uv if self._chroma else y


class MobiclipDecoder:
    def __init__(self, Width, Height, Version, device=0):
        self._lib = N.mobicuda()
        self.Width, self.Height, self.Version = int(Width), int(Height), MobiclipVersion(Version)
        self.Stride = _stride_for(self.Width)
        self.Data = None
        self.Offset = 0
        h = C.c_void_p()
        rc = self._lib.mobi_create(self.Width, self.Height, int(self.Version), device, C.byref(h))
        if rc != 0:
            raise MobiError(rc, 'mobi_create(%d, %d, %s) failed' % (self.Width, self.Height, self.Version.name))
        self._h = h
        self.Y = _Ring(self, False)
        self.UV = _Ring(self, True)
        self._cache = None

    # -- the per-frame call ---------------------------------------------------------------------
    def DecodeFrame(self, want_bitmap=True):
        """Decodes the frame at Data[Offset:]; advances Offset like the reference.  Returns the 32bpp bitmap as an
        (H, W, 4) uint8 array in memory order B,G,R,A (MD:320), or None where the reference returns null."""
        ptr, n = _as_u8(self.Data)
        off = C.c_int(self.Offset)
        rc = self._lib.mobi_decode_frame(self._h, ptr, n, C.byref(off))
        self._cache = None
        if rc != 0:
            self.last_status = rc
            return None
        self.last_status = 0
        self.Offset = off.value
        if not want_bitmap:
            return True
        out = np.empty((self.Height, self.Width, 4), dtype=np.uint8)
        rc = self._lib.mobi_read_bgra(self._h, out.ctypes.data_as(C.c_void_p), self.Width * 4)
        if rc != 0:
            raise MobiError(rc, self.last_error())
        return out

    def SubmitPacked(self, packed):
        """Pre-parsed path (mobi_submit_packed): `packed` is a _native.PackedFrame."""
        rc = self._lib.mobi_submit_packed(self._h, C.byref(packed))
        self._cache = None
        if rc != 0:
            raise MobiError(rc, self.last_error())

    # -- public fields of the reference object -----------------------------------------------------
    @property
    def Quantizer(self):
        return self._state()[0]

    @property
    def YuvFormat(self):
        return self._state()[1]

    def _state(self):
        q, f, s = C.c_uint32(), C.c_uint32(), C.c_int()
        self._lib.mobi_get_state(self._h, C.byref(q), C.byref(f), C.byref(s))
        return q.value, f.value, s.value

    def _read_strided(self):
        if self._cache is None:
            y = np.empty(self.Stride * self.Height, dtype=np.uint8)
            uv = np.empty(self.Stride * self.Height // 2, dtype=np.uint8)
            rc = self._lib.mobi_read_planes_strided(self._h, y.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p))
            if rc != 0:
                raise MobiError(rc, self.last_error())
            self._cache = (y, uv)
        return self._cache

    def ReadYuv(self):
        """Cropped planar I420: (Y[H,W], U[H/2,W/2], V[H/2,W/2])."""
        y = np.empty((self.Height, self.Width), dtype=np.uint8)
        u = np.empty((self.Height // 2, self.Width // 2), dtype=np.uint8)
        v = np.empty_like(u)
        rc = self._lib.mobi_read_yuv(self._h, y.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise MobiError(rc, self.last_error())
        return y, u, v

    def last_error(self):
        return self._lib.mobi_last_error(self._h).decode()

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mobi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MobiParser:
    """Host-only entropy parse (mobi_parser_*): frame bytes -> packed arrays.  Needs no GPU."""

    def __init__(self, Width, Height, Version):
        self._lib = N.mobicuda()
        h = C.c_void_p()
        rc = self._lib.mobi_parser_create(int(Width), int(Height), int(Version), C.byref(h))
        if rc != 0:
            raise MobiError(rc, 'mobi_parser_create failed')
        self._h = h

    def parse(self, data, offset=0):
        """Returns (status, new_offset, PackedFrame view or None).  The view is valid until the next call."""
        ptr, n = _as_u8(data)
        off = C.c_int(offset)
        pf = N.PackedFrame()
        rc = self._lib.mobi_parser_parse(self._h, ptr, n, C.byref(off), C.byref(pf))
        return rc, off.value, (pf if rc == 0 else None)

    def last_error(self):
        return self._lib.mobi_parser_last_error(self._h).decode()

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mobi_parser_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PackedInputs:
    """One step's frames in one host buffer plus the argument tables of mobi_batch_decode (see MobiBatch.pack_inputs)."""
    __slots__ = ('blob', 'ptrs', 'lens', 'offs')

    def __init__(self, blob, ptrs, lens, offs):
        self.blob, self.ptrs, self.lens, self.offs = blob, ptrs, lens, offs


class MobiBatch:
    """N independent streams of one geometry advancing in lock step on one GPU (mobi_batch_*)."""

    def __init__(self, Width, Height, Version, n_streams, device=0, n_threads=0):
        self._lib = N.mobicuda()
        self.Width, self.Height, self.Version = int(Width), int(Height), MobiclipVersion(Version)
        self.Stride = _stride_for(self.Width)
        self.n_streams = int(n_streams)
        h = C.c_void_p()
        rc = self._lib.mobi_batch_create(self.Width, self.Height, int(self.Version), device, self.n_streams, n_threads, C.byref(h))
        if rc != 0:
            raise MobiError(rc, 'mobi_batch_create failed')
        self._h = h
        self._ptrs = (C.c_void_p * self.n_streams)()
        self._lens = (C.c_int * self.n_streams)()
        self._offs = (C.c_int * self.n_streams)()
        self._status = (C.c_int * self.n_streams)()
        self._keep = None

    def _check(self, rc):
        if rc != 0:
            raise MobiError(rc, self._lib.mobi_batch_last_error(self._h).decode())

    def pack_inputs(self, frames, offsets=None):
        """Marshal one step's arguments once: the frames are laid end to end in one host buffer and the pointer /
        length / offset tables the C call takes are built from it.  Pass the result to decode / submit / stage in
        place of the list to keep per-call Python overhead out of the way (a C# or C caller passes such tables as is)."""
        if len(frames) != self.n_streams:
            raise ValueError('need one frame per stream')
        lens = np.fromiter((len(f) for f in frames), dtype=np.int32, count=self.n_streams)
        starts = np.zeros(self.n_streams, dtype=np.int64)
        np.cumsum(lens[:-1], out=starts[1:])
        blob = np.frombuffer(b''.join(bytes(f) if not isinstance(f, (bytes, bytearray)) else f for f in frames), dtype=np.uint8)
        ptrs = (starts + blob.ctypes.data).astype(np.uint64)
        offs = np.zeros(self.n_streams, dtype=np.int32) if offsets is None else np.asarray(offsets, dtype=np.int32).copy()
        return _PackedInputs(blob, ptrs, lens, offs)

    def _marshal(self, frames, offsets):
        if isinstance(frames, _PackedInputs):
            C.memmove(self._ptrs, frames.ptrs.ctypes.data, 8 * self.n_streams)
            C.memmove(self._lens, frames.lens.ctypes.data, 4 * self.n_streams)
            C.memmove(self._offs, frames.offs.ctypes.data, 4 * self.n_streams)
            self._keep = frames
            return
        if len(frames) != self.n_streams:
            raise ValueError('need one frame per stream')
        keep = []
        for i, f in enumerate(frames):
            p, n = _as_u8(f)
            keep.append(f)
            self._ptrs[i] = p
            self._lens[i] = n
            self._offs[i] = 0 if offsets is None else offsets[i]
        self._keep = keep

    def decode(self, frames, offsets=None):
        """One frame per stream.  Returns (offsets after, per-stream status)."""
        self._marshal(frames, offsets)
        rc = self._lib.mobi_batch_decode(self._h, self._ptrs, self._lens, self._offs, self._status)
        if rc != 0 and self.n_streams == 1:
            return list(self._offs), list(self._status)
        self._check(rc)
        return list(self._offs), list(self._status)

    OUT_I420, OUT_BGRA = 1, 2

    def submit(self, frames, offsets=None, fmt=1):
        """Pipelined decode (mobi_batch_submit): returns at once; at most two results may be outstanding."""
        self._marshal(frames, offsets)
        self._check(self._lib.mobi_batch_submit(self._h, self._ptrs, self._lens, self._offs, self._status, fmt))
        return self._offs, self._status

    def fetch(self, out=None, copy=True):
        """Oldest outstanding result as uint8 [n_streams, bytes per stream].  copy=False returns a view of the
        library's pinned buffer, valid until the second submit() from now."""
        view, nbytes = C.c_void_p(), C.c_size_t()
        if copy and out is None:
            self._check(self._lib.mobi_batch_fetch(self._h, None, C.byref(view), C.byref(nbytes)))
            a = np.ctypeslib.as_array(C.cast(view, C.POINTER(C.c_uint8)), shape=(nbytes.value,))
            return a.reshape(self.n_streams, -1).copy()
        if copy:
            self._check(self._lib.mobi_batch_fetch(self._h, out.ctypes.data_as(C.c_void_p), None, None))
            return out
        self._check(self._lib.mobi_batch_fetch(self._h, None, C.byref(view), C.byref(nbytes)))
        a = np.ctypeslib.as_array(C.cast(view, C.POINTER(C.c_uint8)), shape=(nbytes.value,))
        return a.reshape(self.n_streams, -1)

    def stage(self, frames, offsets=None):
        self._marshal(frames, offsets)
        self._check(self._lib.mobi_batch_stage(self._h, self._ptrs, self._lens, self._offs))
        return list(self._offs)

    def replay(self, first, count, fmt=0):
        """Reconstruct staged steps [first, first + count); fmt = MobiBatch.OUT_BGRA also converts every step's new pictures
        on the device (no copies)."""
        self._check(self._lib.mobi_batch_replay_convert(self._h, first, count, fmt) if fmt else self._lib.mobi_batch_replay(self._h, first, count))

    def staged_steps(self):
        return self._lib.mobi_batch_staged_steps(self._h)

    def clear_staged(self):
        self._lib.mobi_batch_clear_staged(self._h)

    def reset(self):
        self._check(self._lib.mobi_batch_reset(self._h))

    def reset_streams(self):
        self._check(self._lib.mobi_batch_reset_streams(self._h))

    def sync(self):
        self._check(self._lib.mobi_batch_sync(self._h))

    def cuda_stream(self):
        return self._lib.mobi_batch_cuda_stream(self._h)

    def read_yuv(self, out=None):
        """Tight I420 of every stream's newest picture: uint8 [n_streams, W*H*3/2]."""
        if out is None:
            out = np.empty((self.n_streams, self.Width * self.Height * 3 // 2), dtype=np.uint8)
        self._check(self._lib.mobi_batch_read_yuv(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def read_bgra_all(self, out=None, device_only=False):
        if device_only:
            self._check(self._lib.mobi_batch_read_bgra_all(self._h, None))
            return None
        if out is None:
            out = np.empty((self.n_streams, self.Height, self.Width, 4), dtype=np.uint8)
        self._check(self._lib.mobi_batch_read_bgra_all(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def read_planes_strided(self, stream):
        y = np.empty(self.Stride * self.Height, dtype=np.uint8)
        uv = np.empty(self.Stride * self.Height // 2, dtype=np.uint8)
        self._check(self._lib.mobi_batch_read_planes_strided(self._h, stream, y.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p)))
        return y, uv

    def read_bgra(self, stream):
        out = np.empty((self.Height, self.Width, 4), dtype=np.uint8)
        self._check(self._lib.mobi_batch_read_bgra(self._h, stream, out.ctypes.data_as(C.c_void_p), self.Width * 4))
        return out

    def stats(self):
        st = N.BatchStats()
        self._check(self._lib.mobi_batch_get_stats(self._h, C.byref(st)))
        return {n: getattr(st, n) for n, _ in N.BatchStats._fields_}

    def clear_stats(self):
        self._lib.mobi_batch_clear_stats(self._h)

    def phase_times(self):
        """Host wall time since clear_stats(), ms: {'parse', 'pack', 'enqueue', 'fetch_wait'} (mobi_batch_get_phase_times)."""
        ms = (C.c_double * 4)()
        self._check(self._lib.mobi_batch_get_phase_times(self._h, ms))
        return {'parse': ms[0], 'pack': ms[1], 'enqueue': ms[2], 'fetch_wait': ms[3]}

    def set_kernel_timing(self, on):
        self._check(self._lib.mobi_batch_set_kernel_timing(self._h, 1 if on else 0))

    def kernel_times(self):
        """Per-kernel device time since the last call (synchronises):
        {'inter_ms', 'inter_launches' (k_mc), 'res_ms', 'res_launches' (k_res), 'intra_ms', 'intra_launches', 'key_ms', 'key_launches'}."""
        ms, n = (C.c_double * 5)(), (C.c_uint64 * 5)()
        self._check(self._lib.mobi_batch_get_kernel_times(self._h, ms, n))
        return {'inter_ms': ms[0], 'inter_launches': n[0], 'intra_ms': ms[1], 'intra_launches': n[1], 'key_ms': ms[2], 'key_launches': n[2],
                'res_ms': ms[3], 'res_launches': n[3], 'bgra_ms': ms[4], 'bgra_launches': n[4]}

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mobi_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MobiMultiBatch:
    """All GPUs of one node from ONE process -- the shape the reference's host has (one C# process, one decoder object per
    stream, MobiclipDecoder.cs:13-61): a MobiBatch per device, each driven from its own host thread (ctypes drops the GIL
    for the duration of a native call, and the library selects its device on every call), global stream g living on
    device g % n_devices (sharding.streams_of_rank).  Nothing crosses devices: streams are independent."""

    def __init__(self, Width, Height, Version, n_streams, devices, n_threads=0):
        from concurrent.futures import ThreadPoolExecutor
        from . import sharding
        self.devices = list(devices)
        if not self.devices or n_streams < len(self.devices):
            raise ValueError('need at least one stream per device')
        world = len(self.devices)
        self.n_streams = int(n_streams)
        self.owned = [sharding.streams_of_rank(self.n_streams, r, world) for r in range(world)]
        per_dev_threads = n_threads if n_threads > 0 else max(1, len(sharding.cores_of_rank(0, world)))
        self.parts = [MobiBatch(Width, Height, Version, len(self.owned[r]), device=d, n_threads=per_dev_threads) for r, d in enumerate(self.devices)]
        self.Width, self.Height = self.parts[0].Width, self.parts[0].Height
        self._pool = [ThreadPoolExecutor(max_workers=1) for _ in self.devices]   # one thread per device, calls stay in order

    def _each(self, fn):
        futs = [self._pool[r].submit(fn, r, b) for r, b in enumerate(self.parts)]
        return [f.result() for f in futs]

    def _split(self, per_stream):
        return [[per_stream[g] for g in own] for own in self.owned]

    def _merge(self, per_part):
        out = [None] * self.n_streams
        for own, vals in zip(self.owned, per_part):
            for g, v in zip(own, vals):
                out[g] = v
        return out

    def decode(self, frames, offsets=None):
        """One frame per global stream.  Returns (offsets after, status), both in global stream order."""
        fr, of = self._split(frames), (self._split(offsets) if offsets is not None else [None] * len(self.parts))
        res = self._each(lambda r, b: b.decode(fr[r], of[r]))
        return self._merge([o for o, _ in res]), self._merge([s for _, s in res])

    def submit(self, frames, fmt=MobiBatch.OUT_I420):
        fr = self._split(frames)
        self._each(lambda r, b: b.submit(fr[r], fmt=fmt) and None)

    def fetch(self):
        """Oldest outstanding result of every device, as uint8 [n_streams, bytes per stream] in global stream order."""
        res = self._each(lambda r, b: b.fetch())
        out = np.empty((self.n_streams, res[0].shape[1]), dtype=np.uint8)
        for own, a in zip(self.owned, res):
            out[own] = a
        return out

    def read_yuv(self):
        res = self._each(lambda r, b: b.read_yuv())
        out = np.empty((self.n_streams, res[0].shape[1]), dtype=np.uint8)
        for own, a in zip(self.owned, res):
            out[own] = a
        return out

    def read_bgra_all(self):
        res = self._each(lambda r, b: b.read_bgra_all())
        out = np.empty((self.n_streams, self.Height, self.Width, 4), dtype=np.uint8)
        for own, a in zip(self.owned, res):
            out[own] = a
        return out

    def sync(self):
        self._each(lambda r, b: b.sync())

    def close(self):
        for b in self.parts:
            b.close()
        for p in self._pool:
            p.shutdown(wait=True)
        self.parts = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
