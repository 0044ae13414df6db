import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')
    config.addinivalue_line('markers', 'slow: long-running')
    # the native pieces are built once per session; building is not using
    from kwiiyatta_b200 import build as kw_build
    from oracle import build as oracle_build
    kw_build.build()
    oracle_build.build()


@pytest.fixture(scope='session')
def cuda():
    import torch
    assert torch.cuda.is_available(), 'gpu-marked test started without a CUDA device'
    return torch
