import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


class _TuningEnv:
    """The GESPMM_* environment is read once by the library (include/gespmm.h: gespmm_reload_env); tests that change
    it mid-process go through this helper, which re-reads it after every change and restores it afterwards."""

    def __init__(self, monkeypatch):
        from gespmm_b200 import capi
        self._mp, self._capi = monkeypatch, capi

    def setenv(self, name, value):
        self._mp.setenv(name, value)
        self._capi.reload_env()

    def delenv(self, name, raising=True):
        self._mp.delenv(name, raising=raising)
        self._capi.reload_env()


@pytest.fixture
def gespmm_env(pkg, monkeypatch):
    env = _TuningEnv(monkeypatch)
    yield env
    monkeypatch.undo()
    env._capi.reload_env()


@pytest.fixture(scope="session")
def oracle():
    o = entry.load_oracle()
    o.lib()
    return o


@pytest.fixture(scope="session")
def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def expected_mtx():
    with open(os.path.join(GOLDEN, "expected_mtx.json")) as f:
        return json.load(f)


def load_golden_csr(name):
    z = np.load(os.path.join(GOLDEN, name + "_csr.npz"))
    return z["rowptr"], z["colind"], tuple(int(x) for x in z["shape"])


@pytest.fixture(scope="session")
def golden_csr():
    return load_golden_csr
