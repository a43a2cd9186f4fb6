import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _native_libs():
    """Build (or find prebuilt) native libraries once per session.  On the GPU box /root/reference is absent and
    the prebuilt oracle/_ref/libmobiref.so that travelled with the snapshot is used as is."""
    from mobiclipdecoder_b200 import _build
    _build.build_all()
