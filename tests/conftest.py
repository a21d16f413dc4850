import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

GOLDEN = os.path.join(REPO, "tests", "golden")
CASES = os.path.join(GOLDEN, "cases")
THERM = os.path.join(CASES, "therm_air.dat")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The native artefacts are built by __graft_entry__.build(); build the CPU-side ones on demand
    so that `pytest -m 'not gpu'` works from a clean checkout."""
    from par2d_b200 import _capi
    import subprocess
    if not os.path.exists(_capi.HOST_LIB):
        subprocess.run(["make", "../libpar2d_host.so"], cwd=os.path.join(REPO, "par2d_b200", "csrc"), check=True)
    if not os.path.exists(_capi.ORACLE_LIB):
        subprocess.run(["make", "libpar2d_oracle.so"], cwd=os.path.join(REPO, "oracle"), check=True)


@pytest.fixture(scope="session")
def oracle():
    from par2d_b200 import _capi
    return _capi.load_oracle()


@pytest.fixture(scope="session")
def gpu():
    from par2d_b200 import _capi
    return _capi.load_gpu()   # no fallback: a missing CUDA library fails the GPU tests loudly


def case_path(name):
    return os.path.join(CASES, name)


def load_golden(name, variant):
    return np.load(os.path.join(GOLDEN, f"golden_{name}_{variant}.npz"))


def state_err(a, b):
    """max over variables of max|a-b| / max|b|  (BASELINE.md section 5)."""
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    worst = 0.0
    for k in range(4):
        scale = max(np.abs(b[:, k]).max(), 1e-300)
        worst = max(worst, np.abs(a[:, k] - b[:, k]).max() / scale)
    return worst


def make_cfg(lib, iaxi=1, igam=1.4, mode=0):
    from par2d_b200 import default_config
    cfg = default_config(lib)
    cfg.iaxi, cfg.igam, cfg.mode = iaxi, igam, mode
    return cfg


def random_states(rng, n):
    """rho in [0.005,2], u,v in [-2000,2000], T in [210,3000]."""
    u = np.empty((n, 4))
    u[:, 0] = rng.uniform(0.005, 2.0, n)
    u[:, 1] = rng.uniform(-2000, 2000, n)
    u[:, 2] = rng.uniform(-2000, 2000, n)
    u[:, 3] = rng.uniform(210, 3000, n)
    return u


def random_faces(rng, n):
    ang = rng.uniform(0, 2 * np.pi, n)
    f = np.empty((n, 4))
    f[:, 0], f[:, 1] = np.cos(ang), np.sin(ang)
    f[:, 2] = rng.uniform(1e-4, 1.0, n)
    f[:, 3] = rng.uniform(1e-4, 1.0, n)
    return f
