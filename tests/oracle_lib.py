"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/mobi_oracle.c) and, where it was built, to the
compiled transliteration of the reference's own source (oracle/_ref, see oracle/build_ref.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, 'oracle', '_build', 'libmobioracle.so')
REF_SO = os.path.join(ROOT, 'oracle', '_ref', 'libmobiref.so')


def stride_for(w):
    return 256 if w <= 256 else 512 if w <= 512 else 1024


class _CpuDecoder:
    """Common shape of both CPU decoders: decode(data, offset) -> (ok, new_offset, bgra or None); .y/.uv planes."""

    def __init__(self, w, h, version):
        self.W, self.H, self.version = w, h, int(version)
        self.S = stride_for(w)

    def crop(self):
        y = self.y.reshape(self.H, self.S)[:, :self.W]
        c = self.uv.reshape(self.H // 2, self.S)
        return np.ascontiguousarray(y), np.ascontiguousarray(c[:, :self.W // 2]), np.ascontiguousarray(c[:, self.S // 2:self.S // 2 + self.W // 2])

    def i420(self):
        y, u, v = self.crop()
        return np.concatenate([y.ravel(), u.ravel(), v.ravel()])


class Oracle(_CpuDecoder):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(ORACLE_SO)
            L.mobi_oracle_create.restype = C.c_void_p
            L.mobi_oracle_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
            L.mobi_oracle_destroy.argtypes = [C.c_void_p]
            L.mobi_oracle_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
            L.mobi_oracle_y.restype = C.c_void_p
            L.mobi_oracle_y.argtypes = [C.c_void_p]
            L.mobi_oracle_uv.restype = C.c_void_p
            L.mobi_oracle_uv.argtypes = [C.c_void_p]
            L.mobi_oracle_quantizer.restype = C.c_uint32
            L.mobi_oracle_quantizer.argtypes = [C.c_void_p]
            L.mobi_oracle_yuvformat.restype = C.c_uint32
            L.mobi_oracle_yuvformat.argtypes = [C.c_void_p]
            L.mobi_oracle_set_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.mobi_oracle_predict_intra.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint32]
            L.mobi_oracle_plane16.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
            L.mobi_oracle_copy_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
            L.mobi_oracle_idct.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, w, h, version):
        super().__init__(w, h, version)
        self.L = self.lib()
        self.h = self.L.mobi_oracle_create(w, h, int(version))

    def decode(self, data, offset=0, want_bgra=True):
        buf = np.frombuffer(data, dtype=np.uint8)
        off = C.c_int(offset)
        bgra = np.empty((self.H, self.W, 4), dtype=np.uint8) if want_bgra else None
        ok = self.L.mobi_oracle_decode(self.h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(off), bgra.ctypes.data_as(C.c_void_p) if want_bgra else None)
        return bool(ok), off.value, (bgra if ok else None)

    @property
    def y(self):
        return np.ctypeslib.as_array(C.cast(self.L.mobi_oracle_y(self.h), C.POINTER(C.c_uint8)), shape=(self.S * self.H,)).copy()

    @property
    def uv(self):
        return np.ctypeslib.as_array(C.cast(self.L.mobi_oracle_uv(self.h), C.POINTER(C.c_uint8)), shape=(self.S * self.H // 2,)).copy()

    @property
    def quantizer(self):
        return self.L.mobi_oracle_quantizer(self.h)

    @property
    def yuvformat(self):
        return self.L.mobi_oracle_yuvformat(self.h)

    # primitive hooks
    def set_planes(self, y, uv):
        self.L.mobi_oracle_set_planes(self.h, y.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p))

    def predict_intra(self, mode, plane, offset, window):
        return self.L.mobi_oracle_predict_intra(self.h, mode, plane, offset, window)

    def plane16(self, offset, window):
        return self.L.mobi_oracle_plane16(self.h, offset, window)

    def copy_block(self, plane, src, dx, dy, w, h, offset):
        return self.L.mobi_oracle_copy_block(self.h, plane, src.ctypes.data_as(C.c_void_p), dx, dy, w, h, offset)

    def idct(self, plane, n, coef, endpos, offset):
        coef = np.ascontiguousarray(coef, dtype=np.int32)
        return self.L.mobi_oracle_idct(self.h, plane, n, coef.ctypes.data_as(C.c_void_p), endpos, offset)

    def __del__(self):
        try:
            self.L.mobi_oracle_destroy(self.h)
        except Exception:
            pass


def have_ref():
    return os.path.exists(REF_SO)


class Ref(_CpuDecoder):
    """The reference's own decoder source, compiled (oracle/_ref).  Present wherever build_ref.py ran."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.mobiref_create.restype = C.c_void_p
            L.mobiref_create.argtypes = [C.c_uint, C.c_uint, C.c_int]
            L.mobiref_destroy.argtypes = [C.c_void_p]
            L.mobiref_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
            L.mobiref_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.mobiref_quantizer.restype = C.c_uint
            L.mobiref_quantizer.argtypes = [C.c_void_p]
            L.mobiref_yuvformat.restype = C.c_uint
            L.mobiref_yuvformat.argtypes = [C.c_void_p]
            L.mobiref_set_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.mobiref_predict_intra.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_uint]
            L.mobiref_plane16.argtypes = [C.c_void_p, C.c_int, C.c_uint]
            L.mobiref_copy_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int]
            L.mobiref_idct.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, w, h, version):
        super().__init__(w, h, version)
        self.L = self.lib()
        self.h = self.L.mobiref_create(w, h, int(version))

    def decode(self, data, offset=0, want_bgra=True):
        buf = np.frombuffer(data, dtype=np.uint8)
        off = C.c_int(offset)
        bgra = np.empty((self.H, self.W, 4), dtype=np.uint8) if want_bgra else None
        ok = self.L.mobiref_decode(self.h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(off), bgra.ctypes.data_as(C.c_void_p) if want_bgra else None)
        return bool(ok), off.value, (bgra if ok else None)

    def _planes(self):
        y = np.zeros(self.S * self.H, dtype=np.uint8)
        uv = np.zeros(self.S * self.H // 2, dtype=np.uint8)
        self.L.mobiref_planes(self.h, y.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p))
        return y, uv

    @property
    def y(self):
        return self._planes()[0]

    @property
    def uv(self):
        return self._planes()[1]

    @property
    def quantizer(self):
        return self.L.mobiref_quantizer(self.h)

    @property
    def yuvformat(self):
        return self.L.mobiref_yuvformat(self.h)

    def set_planes(self, y, uv):
        self.L.mobiref_set_planes(self.h, y.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p))

    def predict_intra(self, mode, plane, offset, window):
        return self.L.mobiref_predict_intra(self.h, mode, plane, offset, window)

    def plane16(self, offset, window):
        return self.L.mobiref_plane16(self.h, offset, window)

    def copy_block(self, plane, src, dx, dy, w, h, offset):
        return self.L.mobiref_copy_block(self.h, plane, src.ctypes.data_as(C.c_void_p), dx, dy, w, h, offset)

    def idct(self, plane, n, coef, endpos, offset):
        coef = np.ascontiguousarray(coef, dtype=np.int32)
        return self.L.mobiref_idct(self.h, plane, n, coef.ctypes.data_as(C.c_void_p), endpos, offset)

    def __del__(self):
        try:
            self.L.mobiref_destroy(self.h)
        except Exception:
            pass


class Ref2:
    """The reference's SECOND copies of the reconstruction primitives, compiled from its own files
    like the decoder (oracle/build_ref.py): FrameUtil.GetPBlock, MobiEncoder.IDCT64 / IDCT16 / DCT64 / DCT16,
    MacroBlock.GetCompvals8x8 / 4x4 and PredictIntraPlane16x16 / 8x8 / 4x4.  Stateless."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.mobiref2_pblock.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p]
            L.mobiref2_idct.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.mobiref2_fdct.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
            L.mobiref2_compvals.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
            L.mobiref2_plane.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
            cls._lib = L
        return cls._lib

    @classmethod
    def pblock(cls, src, dx, dy, w, h, offset, stride):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        out = np.zeros(w * h, dtype=np.uint8)
        ok = cls.lib().mobiref2_pblock(src.ctypes.data, src.size, dx, dy, w, h, offset, stride, out.ctypes.data)
        return out.reshape(h, w) if ok else None

    @classmethod
    def idct(cls, n, coef, pred):
        coef = np.ascontiguousarray(coef, dtype=np.int32)
        pred = np.ascontiguousarray(pred, dtype=np.uint8)
        out = np.zeros(n * n, dtype=np.uint8)
        ok = cls.lib().mobiref2_idct(n, coef.ctypes.data, pred.ctypes.data, out.ctypes.data)
        return out.reshape(n, n) if ok else None

    @classmethod
    def fdct(cls, n, px):
        px = np.ascontiguousarray(px, dtype=np.int32)
        out = np.zeros(n * n, dtype=np.int32)
        ok = cls.lib().mobiref2_fdct(n, px.ctypes.data, out.ctypes.data)
        return out if ok else None

    @classmethod
    def compvals(cls, n, mode, data, x, y, stride, offset=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(n * n, dtype=np.uint8)
        k = cls.lib().mobiref2_compvals(n, mode, data.ctypes.data, data.size, x, y, stride, offset, out.ctypes.data)
        return out.reshape(n, n) if k == n * n else None

    @classmethod
    def plane(cls, n, data, offset, stride, param):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(n * n, dtype=np.uint8)
        k = cls.lib().mobiref2_plane(n, data.ctypes.data, data.size, offset, stride, param, out.ctypes.data)
        return out.reshape(n, n) if k == n * n else None
