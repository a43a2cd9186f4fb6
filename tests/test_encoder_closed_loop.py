"""Encoder closed loop: oracle/mini_encoder.py emits I- and P-picture streams from source pictures with
the reference's ENCODER-side code (forward / inverse transforms, predictor twins, GetPBlock, BitWriter, EncodeDCT, the
encoder's own tables) and keeps its reconstruction YDec / UVDec (MacroBlock.SetupDCTs, MB:224-509).  The reference decoder
compiled from its source, the oracle and the GPU path must each reproduce YDec / UVDec from the bytes alone."""
import json
import os
import sys

import numpy as np
import pytest

from oracle_lib import Oracle, Ref, Ref2, have_ref
from ref_entropy_frames import RefBitWriter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from mini_encoder import MiniEncoder  # noqa: E402

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref/libmobiref.so not built (needs /root/reference)')
TABLES = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'tables_partition_encoder.json')))


def _source(rng, w, h, t):
    """A moving smooth picture with some texture: residuals quantise to a few coefficients per block, vectors find matches."""
    yy, xx = np.mgrid[0:h, 0:w]
    y = (96 + 60 * np.sin((xx + 3 * t) / 9.0) + 50 * np.cos((yy - 2 * t) / 7.0) + rng.integers(-6, 7, size=(h, w))).clip(0, 255).astype(np.uint8)
    cy, cx = np.mgrid[0:h // 2, 0:w // 2]
    u = (128 + 40 * np.sin((cx + t) / 5.0) + rng.integers(-3, 4, size=(h // 2, w // 2))).clip(0, 255).astype(np.uint8)
    v = (128 + 40 * np.cos((cy + 2 * t) / 6.0) + rng.integers(-3, 4, size=(h // 2, w // 2))).clip(0, 255).astype(np.uint8)
    return y, u, v


def _sequence(w, h, seed, n, quantizers):
    rng = np.random.default_rng(seed)
    enc = MiniEncoder(w, h, quantizers[0], Ref2, RefBitWriter, TABLES, seed)
    out = []
    for t in range(n):
        y, u, v = _source(rng, w, h, t)
        q = quantizers[t % len(quantizers)]
        data = enc.encode_frame(y, u, v, force_intra=(t == 5), new_quantizer=q)
        out.append((data + b'\0\0', enc.YDec.copy(), enc.UVDec.copy()))      # + the two pad bytes of MoLiveDemux.cs:353
    return out, enc.stats


CASES = [(64, 48, 1, 9, [20]), (96, 64, 2, 8, [14, 14, 18, 26]), (256, 32, 3, 6, [30]), (48, 48, 4, 7, [12, 40])]


@pytest.mark.parametrize('w,h,seed,n,qs', CASES)
def test_reference_decoder_and_oracle_reproduce_the_encoders_reconstruction(w, h, seed, n, qs):
    seq, stats = _sequence(w, h, seed, n, qs)
    ref, ora = Ref(w, h, 2), Oracle(w, h, 2)
    for t, (data, ydec, uvdec) in enumerate(seq):
        ok_r, off_r, _ = ref.decode(data, 0, False)
        ok_o, off_o, _ = ora.decode(data, 0, False)
        assert ok_r and ok_o, 'frame %d' % t
        assert off_r == off_o
        assert np.array_equal(ref.y, ydec) and np.array_equal(ref.uv, uvdec), 'frame %d: the reference decoder differs from the encoder reconstruction' % t
        assert np.array_equal(ora.y, ydec) and np.array_equal(ora.uv, uvdec), 'frame %d: the oracle differs from the encoder reconstruction' % t
    # the sequences exercise what they claim to
    assert stats['i_mbs'] and stats['p_inter_mbs'] and stats['units8'] and stats['units4'] and stats['leaves'] > stats['p_inter_mbs']
    if w >= 64 and h >= 48:
        assert stats['p_intra_mbs'] and stats['uncoded_units']


def test_source_pictures_survive_the_loop():
    """Sanity of the closed loop itself: at a fine quantiser the reconstruction is close to the source."""
    rng = np.random.default_rng(9)
    enc = MiniEncoder(64, 48, 12, Ref2, RefBitWriter, TABLES, 9)
    y, u, v = _source(rng, 64, 48, 0)
    enc.encode_frame(y, u, v)
    rec = enc.YDec.reshape(48, 256)[:, :64].astype(int)
    assert np.abs(rec - y.astype(int)).mean() < 3.0


@pytest.mark.gpu
@pytest.mark.parametrize('w,h,seed,n,qs', CASES)
def test_gpu_reproduces_the_encoders_reconstruction(w, h, seed, n, qs):
    from mobiclipdecoder_b200 import MobiclipDecoder
    seq, _ = _sequence(w, h, seed, n, qs)
    dec = MobiclipDecoder(w, h, 2)
    for t, (data, ydec, uvdec) in enumerate(seq):
        dec.Data, dec.Offset = data, 0
        assert dec.DecodeFrame(False) is not None, 'frame %d' % t
        assert np.array_equal(dec.Y[0], ydec) and np.array_equal(dec.UV[0], uvdec), 'frame %d' % t
    dec.close()
