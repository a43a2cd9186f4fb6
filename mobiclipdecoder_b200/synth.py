"""Seeded synthetic Mobiclip streams (include/mobisynth.h).  The reference ships no media; tests and bench.py
draw their inputs from here.  Host only."""
import ctypes as C

from . import _native as N


class SynthParams:
    """Thin attribute view over mobi_synth_params; defaults are BASELINE configs 1/3."""

    def __init__(self, width, height, version, seed, **overrides):
        self.c = N.SynthParamsC()
        N.mobisynth().mobi_synth_default_params(C.byref(self.c), width, height, int(version), seed)
        for k, v in overrides.items():
            if not hasattr(self.c, k):
                raise AttributeError('mobi_synth_params has no field %r' % k)
            setattr(self.c, k, v)


class SynthStream:
    def __init__(self, params):
        self._lib = N.mobisynth()
        self.params = params
        self._h = self._lib.mobi_synth_create(C.byref(params.c))
        if not self._h:
            raise ValueError('mobi_synth_create rejected the parameters')
        self._buf = (C.c_uint8 * (4 << 20))()

    def next_frame(self, pad=2):
        """Returns (payload bytes, is_key).  `pad` zero bytes are appended the way the Moflex demuxer does
        (MoLiveDemux.cs:353) so that the decoder's look-ahead never overruns."""
        key = C.c_int(0)
        n = self._lib.mobi_synth_next_frame(self._h, self._buf, len(self._buf), C.byref(key))
        if n < 0:
            raise RuntimeError('synthetic frame larger than the staging buffer')
        return C.string_at(self._buf, n) + b'\0' * pad, bool(key.value)

    def stats(self):
        st = N.SynthStats()
        self._lib.mobi_synth_last_stats(self._h, C.byref(st))
        return st

    def close(self):
        if self._h:
            self._lib.mobi_synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
