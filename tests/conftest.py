import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """make sure the in-tree shared libraries match the sources (no-op when they are up to date)"""
    from bokego_b200 import build as b
    b.build()
    from oracle import cpu as ocpu
    ocpu.build()


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def positions():
    return _load("positions.npz")


@pytest.fixture(scope="session")
def rules():
    return _load("rules.npz")


@pytest.fixture(scope="session")
def nets_golden():
    return _load("nets.npz")


@pytest.fixture(scope="session")
def playouts():
    return _load("playouts.npz")


@pytest.fixture(scope="session")
def sd17():
    return _load("weights_policy_17.npz")


@pytest.fixture(scope="session")
def sd19():
    return _load("weights_policy_19.npz")


@pytest.fixture(scope="session")
def sd_value(sd19):
    """stand-in ValueNet: policy_19 trunk + seeded head, as built in make_golden.py"""
    from oracle import nets as onets
    sd = dict(sd19)
    sd.update({k: v.numpy() for k, v in onets.standin_value_head(1234).items()})
    return sd
