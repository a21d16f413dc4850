"""Shared helpers of the second-order (ACCUR = 1, DG(P1)) tests."""
import numpy as np

from conftest import load_golden, make_cfg, case_path
from par2d_b200 import Case, Solver

DG_RUNS = [("Case_ffstep8", "dgp1_axi_cpg"), ("Case_rktnoz9", "dgp1_axi_tpg")]


def dg_cfg(lib, iaxi=1, igam=1.4, mode=0):
    cfg = make_cfg(lib, iaxi, igam, mode)
    cfg.accur = 1
    return cfg


def run_golden(lib, name, variant, mode=0):
    g = load_golden(name, variant)
    case = Case.load(case_path(name))
    s = Solver(case, dg_cfg(lib, int(g["iaxi"]), float(g["igam"]), mode), lib=lib)
    info = s.step(int(g["steps"]))
    return g, case, s, info


def perturbed_slopes(s, case, seed, amp=0.02):
    """Random, physically small slopes (a few per cent of the mean) and a perturbed mean on every block."""
    rng = np.random.default_rng(seed)
    for b in range(case.nblocks):
        u = s.get_state(b)
        u *= 1.0 + amp * rng.uniform(-1, 1, u.shape)
        u[..., 2] += 30.0 * rng.uniform(-1, 1, u.shape[:2])
        s.set_state(b, u)
        scale = np.abs(u).max(axis=(0, 1))
        s.set_slopes(b, amp * scale * rng.uniform(-1, 1, u.shape), amp * scale * rng.uniform(-1, 1, u.shape))
