"""Host-only companions of tests/test_gpu_primitives.py: the hand-built frames those tests submit must pass
mobi_packed_validate and stay inside what the oracle's primitives accept (no read outside the plane arrays), and the
stream generator must emit every bin the stream-level GPU tests rely on."""
import numpy as np
import pytest

import test_gpu_primitives as P


@pytest.mark.parametrize('w,h,ver', P.GEOMETRIES)
def test_hand_built_motion_compensation_frames_are_in_contract(w, h, ver):
    P.test_motion_compensation_every_shape_phase_reference(w, h, ver, gpu=False)


@pytest.mark.parametrize('w,h,ver', P.GEOMETRIES)
def test_hand_built_transform_class_frames_are_in_contract(w, h, ver):
    P.test_inverse_transform_size_classes(w, h, ver, gpu=False)


@pytest.mark.parametrize('w,h,ver', P.GEOMETRIES)
def test_hand_built_intra_frames_are_in_contract(w, h, ver):
    P.test_intra_predictors_and_decode_order_hazards(w, h, ver, gpu=False)


def test_generator_covers_every_bin_of_every_gpu_tested_config():
    """The stream-level GPU tests are only as good as what the generator emits: every partition shape, half-pel phase,
    reference index and intra predictor must occur in each configuration those tests decode (mobi_synth_stats histograms)."""
    from mobiclipdecoder_b200.workloads import CONFIGS, make_stream
    for name, n in [('mods_256x192', 64), ('moflex_400x240', 100), ('moc5_640x480', 34)]:
        s = make_stream(name, 0xC0FFEE)
        shape, phase, mode, ref = np.zeros(16, int), np.zeros(4, int), np.zeros(20, int), np.zeros(6, int)
        for _ in range(n):
            s.next_frame()
            st = s.stats()
            shape += np.array(st.shape_hist); phase += np.array(st.phase_hist); mode += np.array(st.mode_hist); ref += np.array(st.ref_hist)
        s.close()
        assert (shape > 0).all(), '%s: partition shapes never emitted: %s' % (name, np.flatnonzero(shape == 0))
        assert (phase > 0).all(), '%s: half-pel phases never emitted: %s' % (name, np.flatnonzero(phase == 0))
        assert (ref[1:6] > 0).all(), '%s: reference indices never emitted: %s' % (name, 1 + np.flatnonzero(ref[1:6] == 0))
        used = [m for m in range(20) if m not in (9, 19)]
        assert all(mode[m] > 0 for m in used), '%s: intra predictors never emitted: %s' % (name, [m for m in used if mode[m] == 0])
