import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFERENCE_DIR = os.environ.get('MOBI_REFERENCE_DIR', '/root/reference')


def pytest_configure(config):
    """Native libraries are built (or found prebuilt) HERE, before collection: the reference-pinned test modules decide
    at import time whether oracle/_ref exists, so building it in a fixture would be too late on a clean tree (the whole
    oracle pin used to be skipped silently on the first run).  On the GPU box /root/reference is absent and the prebuilt
    oracle/_ref/libmobiref.so that travelled with the snapshot is used as is."""
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    from mobiclipdecoder_b200 import _build
    _build.build_all()


def _cuda_devices():
    try:
        rt = ctypes.CDLL('libcudart.so')
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return _cuda_devices() > 0


def pytest_collection_modifyitems(config, items):
    # gpu tests are skipped, not failed, where there is no device (plain `pytest tests` on a CPU box stays green)
    if not _have_gpu():
        skip = pytest.mark.skip(reason='no CUDA device')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip)


def pytest_sessionstart(session):
    # Where the reference sources exist the compiled reference MUST exist too: a missing oracle/_ref there is a failure of
    # the pin, not a reason to skip it.
    ref_so = os.path.join(ROOT, 'oracle', '_ref', 'libmobiref.so')
    if os.path.isdir(REFERENCE_DIR) and not os.path.exists(ref_so):
        pytest.exit('oracle/_ref/libmobiref.so was not built although %s exists: the oracle would go unpinned' % REFERENCE_DIR, returncode=1)


@pytest.fixture(scope='session', autouse=True)
def _native_libs():
    from mobiclipdecoder_b200 import _build
    return _build.build_all()
