import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def pkg():
    """the product package (u-vip-slam_b200/), imported under the name uvip_slam_b200"""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope='session')
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope='session')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'cv2_golden.npz'))
