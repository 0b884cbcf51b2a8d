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
def golden_s2p():
    return np.load(os.path.join(GOLDEN, "touchstone.npz"))


def np_spfile(f, fd, sd, polar=True):
    """Independent numpy restatement of Qucs SPfile 'linear' interpolation (third implementation next to the
    product's and the oracle's): sd is [n,4] complex (S11 S21 S12 S22) at fd; end segments extrapolated."""
    f = np.asarray(f, dtype=float)
    k = np.clip(np.searchsorted(fd, f, side="right") - 1, 0, len(fd) - 2)
    t = ((f - fd[k]) / (fd[k + 1] - fd[k]))[:, None]
    a, b = sd[k], sd[k + 1]
    if not polar:
        return a + t * (b - a)
    dp = np.angle(b) - np.angle(a)
    dp = np.where(dp > np.pi, dp - 2 * np.pi, np.where(dp < -np.pi, dp + 2 * np.pi, dp))
    return (np.abs(a) + t * (np.abs(b) - np.abs(a))) * np.exp(1j * (np.angle(a) + t * dp))


def np_s_to_abcd(s, z0):
    s11, s21, s12, s22 = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
    d = 2 * s21
    return np.stack([((1 + s11) * (1 - s22) + s12 * s21) / d, z0 * ((1 + s11) * (1 + s22) - s12 * s21) / d,
                     ((1 - s11) * (1 - s22) - s12 * s21) / d / z0, ((1 - s11) * (1 + s22) + s12 * s21) / d], axis=1)


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
