"""The fused step, the persistent launch and the single-process multi-context driver (round 2).

Every execution strategy of the first-order path must give the same bits:
  * persistent launch (a whole batch of a small mesh in one cooperative kernel) vs one launch per step vs the
    separate ghost + march launches they replace;
  * N contexts of one process connected through peer-mapped arenas (par2d_peer_connect_local; the contexts share
    GPU 0 when the box has one GPU) vs one context: state, residual, dt and the GLOBAL residual history;
  * the tolerance stop decided on the device from the sums of all contexts (main.cpp:506-531) vs a run of exactly
    that many steps."""
import os

import numpy as np
import pytest

from conftest import case_path, load_golden, state_err
from par2d_b200 import Case, Solver, default_config, MODE_FAST, MODE_FAITHFUL
from par2d_b200.local import LocalMultiSolver

pytestmark = pytest.mark.gpu


def _cfg(gpu, mode=MODE_FAST, igam=1.4, tol=None):
    cfg = default_config(gpu)
    cfg.device, cfg.mode, cfg.igam = 0, mode, igam
    if tol is not None:
        cfg.tol = tol
    return cfg


def _run(gpu, name, mode, igam, batches, stop=False, tol=None):
    case = Case.load(case_path(name))
    s = Solver(case, _cfg(gpu, mode, igam, tol), lib=gpu)
    info = []
    for n in batches:
        info += s.step(n, stop_on_tol=stop)
    out = ([s.get_state(b) for b in range(case.nblocks)], [(i.iter, i.dt, i.relres, i.ressumsum, i.status) for i in info], s.iteration)
    s.close()
    return out


@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
@pytest.mark.parametrize("name,igam", [("Case_ffstep1", 1.4), ("Case_ffstep8", 1.4), ("Case_rktnoz9", -1.4), ("Case_spikefine24", 1.4)])
def test_persistent_launch_bitwise(gpu, monkeypatch, name, igam, mode):
    """40 steps of a shipped case as batches of 17 + 1 + 22 in the three execution strategies."""
    runs = {}
    for label, env in (("persistent", {}), ("per_step", {"PAR2D_NO_PERSIST": "1"}), ("separate", {"PAR2D_NO_OVERLAP": "1"})):
        for k in ("PAR2D_NO_PERSIST", "PAR2D_NO_OVERLAP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        runs[label] = _run(gpu, name, mode, igam, [17, 1, 22])
    ref = runs["separate"]
    assert len(ref[1]) == 40 and all(r[4] == 0 for r in ref[1])
    for label in ("persistent", "per_step"):
        got = runs[label]
        assert got[1] == ref[1], label
        for a, b in zip(got[0], ref[0]):
            assert np.array_equal(a, b), label


@pytest.mark.parametrize("name,igam", [("Case_ffstep1", 1.4), ("Case_rktnoz9", -1.4)])
def test_deferred_bookkeeping_bitwise(gpu, monkeypatch, name, igam):
    """Without a stop test the persistent launch keeps every step's strip sums and the history records are made after
    the launch (three short kernels), in launches of at most 1024 steps: 1100 + 5 + 30 steps give the records and the
    state of the same batches with the bookkeeping CTAs inside the launch (PAR2D_NO_DEFER=1), bit for bit."""
    monkeypatch.delenv("PAR2D_NO_PERSIST", raising=False)
    monkeypatch.delenv("PAR2D_NO_OVERLAP", raising=False)
    monkeypatch.delenv("PAR2D_NO_DEFER", raising=False)
    deferred = _run(gpu, name, MODE_FAST, igam, [1100, 5, 30])
    monkeypatch.setenv("PAR2D_NO_DEFER", "1")
    inside = _run(gpu, name, MODE_FAST, igam, [1100, 5, 30])
    assert len(deferred[1]) == 1135 and deferred[2] == 1135
    assert deferred[1] == inside[1]
    for a, b in zip(deferred[0], inside[0]):
        assert np.array_equal(a, b)


def test_persistent_launch_against_the_reference_golden(gpu):
    """The persistent launch in the bit-faithful mode reproduces the unmodified reference binary's state after 200
    steps of the 8-block forward step, bit for bit, in ONE kernel launch."""
    g = load_golden("Case_ffstep8", "axi_cpg")
    states, hist, _ = _run(gpu, "Case_ffstep8", MODE_FAITHFUL, 1.4, [int(g["steps"])])
    for b, s in enumerate(states):
        assert np.array_equal(s, g[f"b{b}"]), b
    assert np.abs(np.array([h[2] for h in hist]) / g["relres"] - 1).max() < 2e-6


@pytest.mark.parametrize("persist", ["1", "0"])
def test_stop_on_tol_every_strategy(gpu, monkeypatch, persist):
    """relres of the 8-block forward step: 1, .99975, .99959, ... -> tol .9996 stops after iteration 2; the state is
    the one after exactly 3 steps, in the middle of a batch and in every execution strategy."""
    monkeypatch.setenv("PAR2D_NO_PERSIST", "0" if persist == "1" else "1")
    stopped = _run(gpu, "Case_ffstep8", MODE_FAST, 1.4, [50], stop=True, tol=0.9996)
    exact = _run(gpu, "Case_ffstep8", MODE_FAST, 1.4, [3])
    assert len(stopped[1]) == 3 and stopped[2] == 3 and stopped[1][-1][2] < 0.9996
    assert stopped[1] == exact[1]
    for a, b in zip(stopped[0], exact[0]):
        assert np.array_equal(a, b)


def test_per_step_status_names_the_failing_iteration(gpu):
    """A cell that blows up: the records before the failure carry status 0, the failing one and all later ones not."""
    case = Case.load(case_path("Case_ffstep8"))
    s = Solver(case, _cfg(gpu), lib=gpu)
    u = s.get_state(0)
    u[3, 3, 0] = 1e-12
    u[3, 3, 1] = 9000.0
    s.set_state(0, u)
    info = s.step(200, check=False)
    status = [i.status for i in info]
    assert status[-1] != 0
    first = next(k for k, st in enumerate(status) if st != 0)
    assert all(st == 0 for st in status[:first]) and all(st != 0 for st in status[first:])
    s.close()


def test_batched_state_transfer(gpu):
    case = Case.load(case_path("Case_spikefine24"))
    s = Solver(case, _cfg(gpu), lib=gpu)
    rng = np.random.default_rng(7)
    want = {}
    for b in range(case.nblocks):
        nx, ny = case.shape(b)
        u = np.empty((ny - 1, nx - 1, 4))
        u[..., 0] = rng.uniform(0.01, 1.0, u.shape[:2]); u[..., 1] = rng.uniform(-500, 1500, u.shape[:2])
        u[..., 2] = rng.uniform(-300, 300, u.shape[:2]); u[..., 3] = rng.uniform(250, 900, u.shape[:2])
        want[b] = u
    s.set_states(want)
    got = s.get_states()
    for b in range(case.nblocks):
        assert np.array_equal(got[b], want[b]) and np.array_equal(s.get_state(b), want[b])
    s.step(3)
    got = s.get_states([5, 0, 23])
    for b in (5, 0, 23):
        assert np.array_equal(got[b], s.get_state(b))
    s.close()


# ------------------------------------------------------------------ N contexts in one process

def _single(gpu, case, mode, batches, igam=1.4, stop=False, tol=None):
    s = Solver(case, _cfg(gpu, mode, igam, tol), lib=gpu)
    dt, rel = [], []
    for n in batches:
        info = s.step(n, stop_on_tol=stop)
        dt += [i.dt for i in info]
        rel += [i.relres for i in info]
        if stop and rel and rel[-1] < tol:
            break                      # what the reference's loop does on the gathered relres (main.cpp:506-511)
    out = ([s.get_state(b) for b in range(case.nblocks)], np.array(dt), np.array(rel), [s.get_residual(b) for b in range(case.nblocks)], s.iteration)
    s.close()
    return out


def _multi(gpu, case, mode, batches, nranks, igam=1.4, stop=False, tol=None):
    m = LocalMultiSolver(case, _cfg(gpu, mode, igam, tol), nranks=nranks, lib=gpu)
    dt, rel = [], []
    for n in batches:
        d, r = m.step(n, stop_on_tol=stop)
        dt += list(d)
        rel += list(r)
        if stop and rel and rel[-1] < tol:
            break
    if stop:
        m.step(0)          # refresh the halos of the state the run stopped in (par2d_gpu.h, par2d_step)
    out = ([m.get_state(b) for b in range(case.nblocks)], np.array(dt), np.array(rel), [m.get_residual(b) for b in range(case.nblocks)], m.iteration)
    m.close()
    return out


@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
@pytest.mark.parametrize("what,nranks", [("Case_spikefine24", 2), ("Case_spikefine24", 5), ("Case_rktnoz9", 3), ("Case_ffstep8", 8), ("ramp", 2), ("ramp", 4)])
def test_contexts_of_one_process_match_one_context_bitwise(gpu, what, nranks, mode):
    """Blocks sharded over N peer-connected contexts (sharing the box's GPUs): the fused step's puts into the
    neighbours' arenas, the flags, the pulled ghost cells and the dt table must leave no trace in the numbers."""
    make = (lambda: Case.synthetic_ramp(256, 192, 4, 4)) if what == "ramp" else (lambda: Case.load(case_path(what)))
    igam = -1.4 if what == "Case_rktnoz9" else 1.4
    one = _single(gpu, make(), mode, [7, 1, 6], igam)
    many = _multi(gpu, make(), mode, [7, 1, 6], nranks, igam)
    assert np.array_equal(one[1], many[1]) and np.array_equal(one[2], many[2]) and len(one[1]) == 14
    for a, b in zip(one[0], many[0]):
        assert np.array_equal(a, b)
    for a, b in zip(one[3], many[3]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", [MODE_FAST, MODE_FAITHFUL])
def test_large_mesh_strategy_across_contexts_bitwise(gpu, monkeypatch, mode):
    """Large meshes across ranks run the tuned launch-per-step kernel with the strips that border another rank and
    the put on a side stream beside the interior strips (capi.cu::queue_step_overlapped).  Forced here on a small
    mesh (PAR2D_SMALL_MESH=0), three contexts sharing the GPU: bit-identical to one context."""
    monkeypatch.setenv("PAR2D_SMALL_MESH", "0")
    make = lambda: Case.synthetic_ramp(256, 192, 4, 4)
    one = _single(gpu, make(), mode, [5, 1, 4])
    many = _multi(gpu, make(), mode, [5, 1, 4], 3)
    assert np.array_equal(one[1], many[1]) and np.array_equal(one[2], many[2]) and len(one[1]) == 10
    for a, b in zip(one[0], many[0]):
        assert np.array_equal(a, b)
    for a, b in zip(one[3], many[3]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("nranks", [2, 3])
def test_stop_across_contexts(gpu, nranks):
    """tol .9996 on the 8-block forward step: every context stops after iteration 2 (the decision is made on the
    device from the published sums of all contexts), in the middle of a batch, at its very end (the tail test) and
    one before; the state equals the single-context run's."""
    case = lambda: Case.load(case_path("Case_ffstep8"))
    one = _single(gpu, case(), MODE_FAST, [50], stop=True, tol=0.9996)
    assert one[4] == 3
    for batches in ([50], [3], [4], [2, 5], [1, 1, 1, 1]):
        many = _multi(gpu, case(), MODE_FAST, batches, nranks, stop=True, tol=0.9996)
        assert many[4] == 3, batches
        assert np.array_equal(one[1], many[1][:3]) and np.array_equal(one[2], many[2][:3])
        for a, b in zip(one[0], many[0]):
            assert np.array_equal(a, b), batches
        for a, b in zip(one[3], many[3]):
            assert np.array_equal(a, b), batches


def test_contexts_vs_reference_golden(gpu):
    """24 blocks on 3 contexts, 200 steps, against the unmodified reference binary's states."""
    g = load_golden("Case_spikefine24", "axi_cpg")
    many = _multi(gpu, Case.load(case_path("Case_spikefine24")), MODE_FAST, [int(g["steps"])], 3)
    worst = max(state_err(many[0][b], g[f"b{b}"]) for b in range(24))
    assert worst <= 1e-11, worst
    assert np.abs(many[2] / g["relres"] - 1).max() < 2e-6
