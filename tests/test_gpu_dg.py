"""Second-order branch (cfg.accur = 1, the reference's DG(P1) path) on the GPU, through the C ABI:
the CUDA kernels of par2d_b200/csrc/dg.cuh against the oracle (bit-exact against the reference binary, see
tests/test_oracle_dg_cpu.py) and against the golden fixtures generated from the reference run with ACCUR = 1.

Bars: PAR2D_MODE_FAITHFUL - every cell's dudt, duxdt, duydt and the state after N steps BIT-IDENTICAL on the
forward-step case (no pow() on its faces), residual history identical as printed (%le); PAR2D_MODE_FAST (FMA
contraction) - state max|d|/max|var| <= 1e-11, residual history <= 1e-9 relative."""
import numpy as np
import pytest

import dg
from conftest import case_path, state_err
from par2d_b200 import MODE_FAITHFUL, MODE_FAST, Case, Solver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,igam", [("Case_ffstep8", 1.4), ("Case_rktnoz9", -1.4), ("Case_spikefine24", 1.4)])
@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
def test_dg_right_hand_sides_match_oracle(gpu, oracle, name, igam, mode):
    case = Case.load(case_path(name))
    sg = Solver(case, dg.dg_cfg(gpu, 1, igam, mode), lib=gpu)
    so = Solver(case, dg.dg_cfg(oracle, 1, igam), lib=oracle)
    for s in (sg, so):
        dg.perturbed_slopes(s, case, seed=5)
    for b in range(case.nblocks):
        rg, ro = sg.get_residual_dg(b), so.get_residual_dg(b)
        for nm, x, y in zip(("dudt", "duxdt", "duydt"), rg, ro):
            if mode == MODE_FAITHFUL and name == "Case_ffstep8":
                assert np.array_equal(x, y), f"{name} block {b} {nm}: {np.abs(x - y).max()}"
            else:
                scale = np.abs(y).reshape(-1, 4).max(axis=0)
                assert (np.abs(x - y).reshape(-1, 4).max(axis=0) / scale).max() < 1e-12, f"{name} block {b} {nm}"


@pytest.mark.parametrize("name,variant", dg.DG_RUNS)
@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
def test_dg_nstep_vs_reference_golden(gpu, name, variant, mode):
    g, case, s, info = dg.run_golden(gpu, name, variant, mode)
    assert len(info) == int(g["steps"]) and all(i.status == 0 for i in info)
    assert ["%7.4e" % i.dt for i in info] == ["%7.4e" % v for v in g["dt"]]
    rel = max(abs(i.relres / v - 1.0) for i, v in zip(info, g["relres"]))
    assert rel < 1e-6                      # the golden history carries 7 digits (%le)
    if mode == MODE_FAITHFUL:
        assert ["%le" % i.relres for i in info] == ["%le" % v for v in g["relres"]]
    nod = [s.nodal_dgp1(b) for b in range(case.nblocks)]
    ref = [g[f"n{b}"] for b in range(case.nblocks)]
    if mode == MODE_FAITHFUL and name == "Case_ffstep8":
        for b in range(case.nblocks):
            assert np.array_equal(nod[b], ref[b]), f"block {b}"
    else:
        assert state_err(nod, ref) < 1e-11


def test_dg_300_steps_vs_oracle_and_slopes(gpu, oracle):
    case = Case.load(case_path("Case_ffstep8"))
    sg = Solver(case, dg.dg_cfg(gpu), lib=gpu)
    so = Solver(case, dg.dg_cfg(oracle), lib=oracle)
    ig, io = sg.step(300), so.step(300)
    assert max(abs(a.relres / b.relres - 1.0) for a, b in zip(ig, io)) < 1e-9
    assert state_err([sg.get_state(b) for b in range(8)], [so.get_state(b) for b in range(8)]) < 1e-11
    for b in range(8):
        (ax, ay), (bx, by) = sg.get_slopes(b), so.get_slopes(b)
        u = so.get_state(b).reshape(-1, 4)
        vel = np.sqrt(u[:, 1] ** 2 + u[:, 2] ** 2).max()     # v is ~0 in the forward step: the velocity scale, as state_err does
        scale = np.array([np.abs(u[:, 0]).max(), vel, vel, np.abs(u[:, 3]).max()])
        assert (np.abs(ax - bx).reshape(-1, 4).max(axis=0) / scale).max() < 1e-11
        assert (np.abs(ay - by).reshape(-1, 4).max(axis=0) / scale).max() < 1e-11
        assert np.abs(bx).max() > 0 and np.abs(by).max() > 0


def test_dg_needs_its_config(gpu):
    from par2d_b200._capi import Par2dError
    case = Case.load(case_path("Case_ffstep1"))
    cfg = dg.dg_cfg(gpu)
    cfg.accur = 0
    s = Solver(case, cfg, lib=gpu)
    with pytest.raises(Par2dError):
        s.get_slopes(0)
    cfg.accur = 2
    with pytest.raises(Par2dError):
        Solver(case, cfg, lib=gpu)


@pytest.mark.parametrize("ncx,ncy", [(2, 2), (31, 3), (5, 40), (33, 9)])
@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
def test_dg_small_block_shapes_vs_oracle(gpu, oracle, tmp_path, ncx, ncy, mode):
    """The second-order branch on blocks of a few cells (every cell next to a boundary, the limiter's flat-index
    tests wrapping across the short rows) against the oracle."""
    import os
    nx, ny = ncx + 1, ncy + 1
    x = np.tile(np.linspace(0.0, 0.01 * ncx, nx), (ny, 1))
    y = np.tile(np.linspace(0.02, 0.02 + 0.008 * ncy, ny)[:, None], (1, nx)) + 0.001 * np.sin(40.0 * x)
    ib = np.array([4] * ncx + [3] * ncy + [4] * ncx + [1] * ncy, dtype=np.int32)
    d = tmp_path / f"c{ncx}x{ncy}"
    os.makedirs(d / "subgrids")
    with open(d / "grid.conn", "w") as f:
        f.write("1\n0: (1,-1) (3,-1) (4,-1) (4,-1)\n")
    with open(d / "subgrids" / "grid.block0.dat", "w") as f:
        f.write("npart: 1 1\nipart: 0 0\nngrid: %d %d\n" % (nx, ny))
        for j in range(ny):
            for i in range(nx):
                f.write("%.17e %.17e\n" % (x[j, i], y[j, i]))
        f.write("BC\n" + "".join("%d\n" % c for c in ib))
    case = Case.load(str(d))
    sg = Solver(case, dg.dg_cfg(gpu, 1, 1.4, mode), lib=gpu)
    so = Solver(case, dg.dg_cfg(oracle, 1, 1.4), lib=oracle)
    for s in (sg, so):
        dg.perturbed_slopes(s, case, seed=9, amp=0.01)
    for a, b in zip(sg.get_residual_dg(0), so.get_residual_dg(0)):
        scale = np.abs(b).reshape(-1, 4).max(axis=0) + 1e-300
        assert (np.abs(a - b).reshape(-1, 4).max(axis=0) / scale).max() < 1e-11
    ig, io = sg.step(15), so.step(15)
    assert [i.status for i in ig] == [i.status for i in io]
    assert state_err(sg.get_state(0), so.get_state(0)) < 1e-11
    assert max(abs(a.relres / b.relres - 1.0) for a, b in zip(ig, io)) < 1e-9
