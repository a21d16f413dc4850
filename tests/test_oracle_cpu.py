"""Pins the oracle (oracle/par2d_oracle.c, the CPU restatement) BITWISE against
  (1) the committed golden fixtures generated from the unmodified reference
      (tests/golden/make_golden.py), on every shipped configuration and variant,
  (2) the reference's own committed residual history FV/Outputs/res.tec, and
  (3) when oracle/_ref exists (this container, or shipped prebuilt to the GPU box), the
      reference's functions called directly through oracle/_ref/libpar2d_ref.so on
      randomised inputs (default_rng(1234)).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, THERM, case_path, load_golden, make_cfg, random_faces, random_states
import pointwise as pw
from par2d_b200 import Case, Solver

VARIANTS = [
    ("Case_ffstep1", "axi_cpg"), ("Case_ffstep8", "axi_cpg"), ("Case_ffstep8", "planar_cpg"),
    ("Case_rktnoz9", "axi_cpg"), ("Case_rktnoz9", "axi_tpg"), ("Case_spikefine24", "axi_cpg"),
]


def fmt_le(v):
    return "%le" % v


@pytest.mark.parametrize("name,variant", VARIANTS)
def test_port_matches_reference_golden_bitwise(oracle, name, variant):
    g = load_golden(name, variant)
    steps = int(g["steps"])
    case = Case.load(case_path(name))
    s = Solver(case, make_cfg(oracle, int(g["iaxi"]), float(g["igam"])), lib=oracle)
    info = s.step(steps)
    assert len(info) == steps and all(i.status == 0 for i in info)
    for b in range(case.nblocks):
        assert np.array_equal(s.get_state(b), g[f"b{b}"]), f"block {b} state differs from the reference"
    # history: the reference prints relres with %le and dt with %7.4e
    assert [fmt_le(i.relres) for i in info] == [fmt_le(v) for v in g["relres"]]
    assert ["%7.4e" % i.dt for i in info] == ["%7.4e" % v for v in g["dt"]]


def test_port_matches_references_committed_res_tec(oracle):
    """FV/Outputs/res.tec is the only golden the reference ships (8-block forward step, shipped macros)."""
    with open(os.path.join(GOLDEN, "ref_committed_res_ffstep8.tec")) as f:
        lines = f.read().splitlines()[1:]
    n = 400
    s = Solver(Case.load(case_path("Case_ffstep8")), make_cfg(oracle), lib=oracle)
    info = s.step(n)
    got = ["%d,\t%le" % (i.iter, i.relres) for i in info]
    assert got == lines[:n]


def test_block_history_sums_to_relres(oracle):
    s = Solver(Case.load(case_path("Case_spikefine24")), make_cfg(oracle), lib=oracle)
    info = s.step(5)
    h = s.block_history(5)
    for k, i in enumerate(info):
        num = den = 0.0
        for b in range(h.shape[1]):
            num += h[k, b, 0]
            den += h[k, b, 1]
        assert num / den == i.relres


def test_stop_on_tol_and_error_codes(oracle):
    case = Case.load(case_path("Case_ffstep8"))
    cfg = make_cfg(oracle)
    cfg.tol = 0.9999
    s = Solver(case, cfg, lib=oracle)
    info = s.step(50, stop_on_tol=True)
    assert len(info) == 2 and info[-1].relres < cfg.tol     # iteration 1 is the first below 0.9999
    # non-positive density -> PAR2D_ERR_DENSITY (main.cpp:401)
    s2 = Solver(case, make_cfg(oracle), lib=oracle)
    u = s2.get_state(0)
    u[3, 3, 0] = 1e-12
    u[3, 3, 1] = 9000.0
    s2.set_state(0, u)
    info = s2.step(200, check=False)
    assert info[-1].status != 0


# ---- function-level pins against the reference's own functions -------------------------------

def _ref_lib():
    path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libpar2d_ref.so")
    path = os.path.abspath(path)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return C.CDLL(path)


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    lib = _ref_lib()
    run = tmp_path_factory.mktemp("refrun")
    os.makedirs(run / "build")
    import shutil
    shutil.copy(THERM, run / "therm_air.dat")
    lib.ref_init.argtypes = [C.c_char_p, C.c_int, C.c_double]
    lib.ref_Rs.restype = C.c_double
    lib._build_dir = str(run / "build").encode()
    return lib


DP = C.POINTER(C.c_double)


def _dp(a):
    return a.ctypes.data_as(DP)


@pytest.mark.parametrize("iaxi,igam", [(1, 1.4), (0, 1.4), (1, -1.4)])
def test_pointwise_port_vs_reference_functions_bitwise(oracle, ref, iaxi, igam):
    assert ref.ref_init(ref._build_dir, iaxi, C.c_double(igam)) == 0
    cfg = make_cfg(oracle, iaxi, igam)
    assert cfg.Rs == ref.ref_Rs()
    rng = np.random.default_rng(1234)
    n = 4000
    uL, uR, face = random_states(rng, n), random_states(rng, n), random_faces(rng, n)
    # make a third of the faces subsonic on both sides and some with tiny pressure jumps
    uR[: n // 3] = uL[: n // 3] * rng.uniform(0.9, 1.1, (n // 3, 4))
    uL[: n // 3, 1:3] *= 0.1
    uR[: n // 3, 1:3] *= 0.1

    # UpdateState
    unk, der = pw.update_state(oracle, cfg, uL)
    rder = np.zeros(8)
    for i in range(0, n, 7):
        u = uL[i].copy()
        ref.ref_update_state(_dp(u), _dp(rder))
        assert np.array_equal(u, unk[i]) and np.array_equal(rder, der[i])

    # LDFSS
    flux, parr = pw.ldfss(oracle, cfg, face, uL, uR)
    rf, rp = np.zeros(4), np.zeros(1)
    ref.ref_ldfss.argtypes = [C.c_double] * 4 + [DP] * 4
    for i in range(n):
        a, b = uL[i].copy(), uR[i].copy()
        ref.ref_ldfss(face[i, 0], face[i, 1], face[i, 2], face[i, 3], _dp(a), _dp(b), _dp(rf), _dp(rp))
        assert np.array_equal(rf, flux[i]) and rp[0] == parr[i], i

    # boundary_state for every BC code incl. the remapped -1 and 0
    ref.ref_boundary_state.argtypes = [C.c_int, C.c_double, C.c_double, DP, DP, DP]
    btype = rng.choice([-1, 0, 1, 2, 3, 4, 5], n).astype(np.int32)
    ui = uL.copy()
    ui[::2, 1:3] *= 0.05   # plenty of subsonic in/outflow for code 1
    ghost = pw.boundary_state(oracle, cfg, btype, face[:, :2], ui)
    uFS = np.array(list(cfg.uFS))
    rg = np.zeros(4)
    for i in range(n):
        a = ui[i].copy()
        rg[:] = 0.0
        ref.ref_boundary_state(int(btype[i]), face[i, 0], face[i, 1], _dp(uFS), _dp(a), _dp(rg))
        rg[3] = max(rg[3], 201.0)      # the ghost's UpdateState clamp, SpatialDiscretization.cpp:39
        assert np.array_equal(rg, ghost[i]), (i, btype[i])

    # Jacobian solve
    ref.ref_jacobian_solve.argtypes = [C.c_double, DP, DP, DP]
    res = rng.normal(0, 1e3, (n, 4))
    dty = np.stack([rng.uniform(1e-9, 1e-6, n), rng.uniform(1e-4, 1.0, n)], axis=1)
    dv = pw.solve(oracle, cfg, dty, uL, res)
    rdv = np.zeros(4)
    for i in range(0, n, 3):
        a = uL[i].copy()
        ref.ref_jacobian_solve(dty[i, 0], _dp(a), _dp(res[i].copy()), _dp(rdv))
        expect = rdv * (1.0 / dty[i, 1]) if iaxi == 1 else rdv
        assert np.array_equal(expect, dv[i]), i


def test_metrics_port_vs_reference_bitwise(oracle, ref):
    case = Case.load(case_path("Case_rktnoz9"))
    s = Solver(case, make_cfg(oracle), lib=oracle)
    for b in (0, 4, 6):
        nx, ny = case.shape(b)
        x, y = case.nodes(b)
        ge, gf, yf, xf = np.zeros(3 * (nx - 1) * (ny - 1)), np.zeros(6 * nx * ny), np.zeros(2 * nx * ny), np.zeros(2 * nx * ny)
        ref.ref_metrics(nx, ny, _dp(x), _dp(y), _dp(ge), _dp(gf), _dp(yf), _dp(xf))
        oge, ogf, oyf = np.zeros_like(ge), np.zeros_like(gf), np.zeros_like(yf)
        oracle.check(oracle.get_metrics(s.handle, b, _dp(oge), _dp(ogf), _dp(oyf)), "get_metrics")
        assert np.array_equal(ge, oge)
        # entries of the last row/column that the reference never writes are uninitialised there
        gf, ogf = gf.reshape(ny, nx, 6), ogf.reshape(ny, nx, 6)
        assert np.array_equal(gf[:, : nx - 1, 0:3], ogf[:, : nx - 1, 0:3]) and np.array_equal(gf[: ny - 1, :, 3:6], ogf[: ny - 1, :, 3:6])
        yf, oyf = yf.reshape(ny, nx, 2), oyf.reshape(ny, nx, 2)
        assert np.array_equal(yf[:, : nx - 1, 0], oyf[:, : nx - 1, 0]) and np.array_equal(yf[: ny - 1, :, 1], oyf[: ny - 1, :, 1])


def test_synthetic_ramp_port_vs_live_reference_bitwise(oracle, tmp_path):
    """The benchmark's synthetic mesh, small (64x64 cells in 2x2 blocks), through real case files."""
    import run_ref
    if not run_ref.available():
        pytest.skip("oracle/_ref not built")
    case = Case.synthetic_ramp(64, 64, 2, 2)
    case.write(str(tmp_path / "Case"))
    loaded = Case.load(str(tmp_path / "Case"))
    shapes = [loaded.shape(b) for b in range(loaded.nblocks)]
    out = run_ref.run(str(tmp_path / "Case"), THERM, shapes, 25)
    s = Solver(loaded, make_cfg(oracle), lib=oracle)
    info = s.step(25)
    for b in range(4):
        assert np.array_equal(s.get_state(b), out["states"][b])
    assert [fmt_le(i.relres) for i in info] == [fmt_le(v) for v in out["relres"]]


def test_references_converged_state_sits_at_its_committed_residual(oracle):
    """The reference's own converged solution files (497 227 iterations, FV/Outputs/solution.b*.tec) read back
    by the restart reader: its residual, evaluated by the port, continues the committed res.tec
    (..., 497226: 1.000012e-06, 497227: 9.999939e-07, falling 1.8e-11 per iteration)."""
    import restart
    g, relres, _ = restart.converged_state_relres(oracle)
    assert list(g["res_tail_iter"]) == [497225, 497226, 497227]
    slope = g["res_tail_relres"][2] - g["res_tail_relres"][1]          # per iteration, ~ -1.8e-11
    expect = g["res_tail_relres"][2] + slope                           # the state was written after update 497227
    assert abs(relres[0] - expect) < 2e-11, (relres, expect)
    assert abs((relres[1] - relres[0]) - slope) < 0.2 * abs(slope)     # and keeps converging at the same rate


# ---- BC 5 (Bernoulli inflow, BoundaryConditions.cpp:135-146) through the residual ----------------------------

def bc5_block_case(tmp_path):
    """Block 11 of the reference's nested FV/Case_spikefine24/Case variant (BC code 5 on its right side) as a
    stand-alone one-block case, with a seeded state that is subsonic and partly reversed next to that side, so that
    both branches of the Bernoulli inflow reach the flux.  Returns (case, ibound, state)."""
    from conftest import overlay_case
    from test_oracle_dg_cpu import _single_block_case
    full = Case.load(overlay_case("Case_spikefine24", "Case_spikefine24_bc5", str(tmp_path / "bc5_full")))
    x, y = full.nodes(11)
    ib = np.ascontiguousarray(full.ibound(11), dtype=np.int32)
    ny, nx = x.shape
    assert set(ib[nx - 1:nx - 1 + ny - 1]) == {5}
    ib[ib < 0] = 3                                   # stand-alone: interior faces become plain extrapolation
    single = _single_block_case(tmp_path / "bc5_single", x, y, ib)
    rng = np.random.default_rng(2025)
    u = np.empty((ny - 1, nx - 1, 4))
    u[..., 0] = rng.uniform(0.008, 0.02, u.shape[:2])
    u[..., 1] = rng.uniform(-150.0, 500.0, u.shape[:2])
    u[..., 2] = rng.uniform(-120.0, 120.0, u.shape[:2])
    u[..., 3] = rng.uniform(250.0, 420.0, u.shape[:2])
    return single, ib, u


def test_bc5_residual_port_vs_reference_bitwise(oracle, ref, tmp_path):
    single, ib, u = bc5_block_case(tmp_path)
    s = Solver(single, make_cfg(oracle), lib=oracle)
    s.set_state(0, u)
    mine = s.get_residual(0)
    assert ref.ref_init(ref._build_dir, 1, C.c_double(1.4)) == 0
    IP = C.POINTER(C.c_int32)
    ref.ref_calc_dudt.argtypes = [C.c_int, C.c_int, DP, DP, IP, DP, DP, DP]
    xs, ys = single.nodes(0)
    ny, nx = xs.shape
    cfg = make_cfg(oracle)
    uFS = np.array(list(cfg.uFS))
    u2, out = u.copy(), np.zeros_like(u)
    ref.ref_calc_dudt(nx, ny, _dp(np.ascontiguousarray(xs)), _dp(np.ascontiguousarray(ys)), ib.ctypes.data_as(IP), _dp(uFS), _dp(u2), _dp(out))
    assert np.array_equal(mine, out)
    # the ghost values matter here: with code 3 (extrapolation) on that side the residual of the last column differs
    ib3 = ib.copy()
    ib3[ib3 == 5] = 3
    out3 = np.zeros_like(u)
    ref.ref_calc_dudt(nx, ny, _dp(np.ascontiguousarray(xs)), _dp(np.ascontiguousarray(ys)), ib3.ctypes.data_as(IP), _dp(uFS), _dp(u.copy()), _dp(out3))
    assert np.abs(out[:, -1] - out3[:, -1]).max() > 1e-3 * np.abs(out[:, -1]).max()
    fixture = os.path.join(GOLDEN, "ref_dudt_bc5_block11.npz")
    if os.environ.get("PAR2D_REGEN_GOLDEN") or not os.path.exists(fixture):
        np.savez_compressed(fixture, dudt=out)
    assert np.array_equal(np.load(fixture)["dudt"], out)
