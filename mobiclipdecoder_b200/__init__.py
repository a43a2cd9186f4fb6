"""mobiclipdecoder_b200 -- B200-native Mobiclip frame reconstruction.

The product is ``lib/libmobicuda.so`` (C ABI in ``include/mobicuda.h``); this package is the thin Python host
side above it: :class:`MobiclipDecoder` mirrors the reference's
``LibMobiclip.Codec.Mobiclip.MobiclipDecoder`` object (MobiclipDecoder.cs:13-61) call for call, and
:class:`MobiBatch` exposes the lock-step multi-stream path.  There is no CPU fallback: if the CUDA library is
missing or no GPU is present, construction raises.
"""
from .decoder import MobiclipDecoder, MobiclipVersion, MobiBatch, MobiMultiBatch, MobiParser, MobiError  # noqa: F401
from .synth import SynthParams, SynthStream  # noqa: F401
