import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def built_lib():
    """Builds (if stale) and loads the in-tree CUDA library."""
    from ifcb_classifier_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope='session')
def cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
