import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"        # only present in the build container; never required by -m gpu tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def Q():
    import qo100net
    if not os.path.exists(qo100net.LIB_PATH):
        qo100net.build()
    qo100net.lib()
    return qo100net


@pytest.fixture(scope="session")
def R():
    from oracle import refbind
    refbind.lib()
    return refbind


@pytest.fixture(scope="session")
def W(Q):
    from qo100net import workloads
    return workloads


@pytest.fixture(scope="session")
def golden_dat():
    return np.load(os.path.join(GOLDEN, "pa_lpf_dat.npz"))


@pytest.fixture(scope="session")
def golden_nets():
    return json.load(open(os.path.join(GOLDEN, "networks.json")))


@pytest.fixture(scope="session")
def golden_b():
    return json.load(open(os.path.join(GOLDEN, "appendix_b.json")))


@pytest.fixture(scope="session")
def ctx(Q):
    """A device context; the CUDA path must be the one that runs -- no fallback."""
    c = Q.Context(device=0)
    yield c
    c.close()


def to_ref(R, net):
    return R.make_elems(net.elements)


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))
