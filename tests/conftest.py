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


def overlay_case(base, overlay, dst):
    """Case directory = tests/golden/cases/<base> with the files of tests/golden/cases/<overlay> laid over it."""
    import shutil
    shutil.copytree(case_path(base), dst)
    for root, _, files in os.walk(case_path(overlay)):
        for f in files:
            if f == "README":
                continue
            rel = os.path.relpath(os.path.join(root, f), case_path(overlay))
            shutil.copy(os.path.join(root, f), os.path.join(dst, rel))
    return dst


def load_golden(name, variant):
    return np.load(os.path.join(GOLDEN, f"golden_{name}_{variant}.npz"))


def state_err(a, b):
    """Worst per-variable difference relative to the variable's scale (BASELINE.md section 5,
    SURVEY 8c tolerance calibration): rho and T against max|var|, the two velocity components
    against the velocity scale max sqrt(u^2+v^2) - v is rounding noise around zero in most cells of the
    forward-step and ramp cases, so max|v| of a block is not a scale.  a, b: arrays (..., 4) or lists
    of per-block arrays (the scale is then taken over all blocks)."""
    if isinstance(a, (list, tuple)):
        a = np.concatenate([x.reshape(-1, 4) for x in a])
        b = np.concatenate([x.reshape(-1, 4) for x in b])
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    vel = max(np.sqrt(b[:, 1] ** 2 + b[:, 2] ** 2).max(), 1e-300)
    scales = [max(np.abs(b[:, 0]).max(), 1e-300), vel, vel, max(np.abs(b[:, 3]).max(), 1e-300)]
    return max(np.abs(a[:, k] - b[:, k]).max() / scales[k] for k in range(4))


def flux_scales(cfg, face, uL, uR):
    """Characteristic magnitude of the four LDFSS flux components of each face: the Mach-1 mass flux
    len*yface*ahalf*(rhoL+rhoR) times (1, speed, speed, total enthalpy) plus the pressure force.  A split
    flux near |M| = 1 or M = 0 is a small difference of O(1) terms - (M -+ 1)^2 in EulerFlux.cpp:152-153 -
    so rounding differences of 1e-16 in M are 1e-12 of THAT face's flux but 1e-16 of this scale, which is
    what enters a residual."""
    T = np.maximum(np.stack([uL[:, 3], uR[:, 3]]), 201.0)
    a = np.sqrt(abs(cfg.igam) * cfg.Rs * T)
    rho = np.stack([uL[:, 0], uR[:, 0]])
    speed = np.maximum(np.sqrt(np.stack([uL[:, 1] ** 2 + uL[:, 2] ** 2, uR[:, 1] ** 2 + uR[:, 2] ** 2])), a).max(axis=0)
    area = face[:, 2] * (face[:, 3] if cfg.iaxi == 1 else 1.0)
    mass = area * a.mean(axis=0) * rho.sum(axis=0)
    p = (rho * cfg.Rs * T).sum(axis=0)
    h0 = (3.5 * cfg.Rs * T).max(axis=0) + 0.5 * speed ** 2
    return np.stack([mass, mass * speed + area * p, mass * speed + area * p, mass * h0], axis=1), face[:, 2] * p


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
