"""Shared body of the restart / converged-state tests: the reference's own committed
converged solution (FV/Outputs/solution.b1..8.tec, fixture made by tests/golden/make_converged_fixture.py)
seeds a run; the residual our step evaluates on it must sit where the reference's committed residual
history FV/Outputs/res.tec ended."""
import os

import numpy as np

from conftest import GOLDEN, case_path, make_cfg
from par2d_b200 import Case, Solver


def load_converged():
    return np.load(os.path.join(GOLDEN, "ref_committed_solution_ffstep8.npz"))


def converged_state_relres(lib, mode=None):
    """relres (main.cpp:494: sum_b sum_k ||res_k|| / sum_b sum_k res0_k, res0 from the uniform initial
    state) of the reference's converged state, and the same one step later."""
    g = load_converged()
    case = Case.load(case_path("Case_ffstep8"))
    cfg = make_cfg(lib)
    if mode is not None:
        cfg.mode = mode
    s = Solver(case, cfg, lib=lib)
    first = s.step(1)                      # iteration 0 from the uniform initial state: fixes res0 (main.cpp:365-374)
    assert first[0].relres == 1.0
    for b in range(case.nblocks):
        s.set_state(b, g[f"unk{b}"])
    info = s.step(2)
    return g, [i.relres for i in info], s
