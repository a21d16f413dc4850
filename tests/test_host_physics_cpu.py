"""CPU-side check of the device formulas: par2d_b200/csrc/physics.cuh compiled for the host
(tests/host_physics, test infrastructure) against the oracle on the same seeded inputs.
  hp1_* = PAR2D_MODE_FAITHFUL formulas, contraction off  -> must be BIT-IDENTICAL to the oracle
  hp0_* = PAR2D_MODE_FAST formulas (lean LDFSS, closed-form solve, series reciprocals on seeds emulated
          at the measured 2^-20 accuracy of MUFU.RCP64H/RSQ64H) -> rounding-level agreement
  hp2_* = hp0 compiled like the fast mode's translation unit (PAR2D_LEAN_BC: far-field ghosts from one logarithm)
The GPU tests (test_gpu_parity.py) repeat the same comparisons on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, flux_scales, make_cfg, random_faces, random_states
import pointwise as pw

HP_DIR = os.path.join(REPO, "tests", "host_physics")


class _HostPhysics:
    """Same call surface as _capi._Lib for tests/pointwise.py, bound to one arithmetic mode."""

    def __init__(self, lib, mode):
        from par2d_b200._capi import Config
        dp, ip, cp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(Config)
        sig = {"test_ldfss": [cp, C.c_int64, dp, dp, dp, dp, dp], "test_boundary_state": [cp, C.c_int64, ip, dp, dp, dp],
               "test_update_state": [cp, C.c_int64, dp, dp], "test_solve": [cp, C.c_int64, dp, dp, dp, dp],
               "test_volume": [C.c_int64, dp, dp], "test_fastmath": [C.c_int64, dp, dp]}
        for name, args in sig.items():
            fn = getattr(lib, f"hp{mode}_{name}")
            fn.argtypes, fn.restype = args, C.c_int
            setattr(self, name, fn)

    def check(self, rc, where):
        assert rc == 0, where


@pytest.fixture(scope="module")
def hplib():
    subprocess.run(["make", "-s"], cwd=HP_DIR, check=True)
    return C.CDLL(os.path.join(HP_DIR, "libhost_physics.so"))


@pytest.fixture(scope="module")
def hp(hplib):
    return _HostPhysics(hplib, 0), _HostPhysics(hplib, 1)


def _inputs(n=20000, seed=1234):
    rng = np.random.default_rng(seed)
    uL, uR, face = random_states(rng, n), random_states(rng, n), random_faces(rng, n)
    uR[: n // 3] = uL[: n // 3] * rng.uniform(0.9, 1.1, (n // 3, 4))
    uL[: n // 3, 1:3] *= 0.1
    uR[: n // 3, 1:3] *= 0.1
    return rng, uL, uR, face


@pytest.mark.parametrize("iaxi,igam", [(1, 1.4), (0, 1.4), (1, -1.4)])
def test_faithful_formulas_bit_identical_to_oracle(hp, oracle, iaxi, igam):
    _, slow = hp
    rng, uL, uR, face = _inputs()
    cfg = make_cfg(oracle, iaxi, igam)
    a, b = pw.update_state(slow, cfg, uL), pw.update_state(oracle, cfg, uL)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    a, b = pw.ldfss(slow, cfg, face, uL, uR), pw.ldfss(oracle, cfg, face, uL, uR)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    btype = rng.choice([-1, 0, 1, 2, 3, 4, 5], len(uL)).astype(np.int32)
    ui = uL.copy()
    ui[::2, 1:3] *= 0.05
    assert np.array_equal(pw.boundary_state(slow, cfg, btype, face[:, :2], ui), pw.boundary_state(oracle, cfg, btype, face[:, :2], ui))
    res = rng.normal(0, 1e3, (len(uL), 4))
    dty = np.stack([rng.uniform(1e-9, 1e-6, len(uL)), rng.uniform(1e-4, 1.0, len(uL))], axis=1)
    assert np.array_equal(pw.solve(slow, cfg, dty, uL, res), pw.solve(oracle, cfg, dty, uL, res))


@pytest.mark.parametrize("iaxi,igam", [(1, 1.4), (0, 1.4), (1, -1.4)])
def test_fast_formulas_agree_with_oracle_to_rounding(hp, oracle, iaxi, igam):
    fast, _ = hp
    rng, uL, uR, face = _inputs()
    cfg = make_cfg(oracle, iaxi, igam)
    (ug, dg), (uo, do) = pw.update_state(fast, cfg, uL), pw.update_state(oracle, cfg, uL)
    assert np.array_equal(ug, uo)
    assert np.abs(dg / do - 1)[:, [0, 1, 5, 6, 7]].max() < 1e-14
    assert (np.abs(dg - do).max(axis=0) / np.abs(do).max(axis=0)).max() < 1e-14
    (fg, pg), (fo, po) = pw.ldfss(fast, cfg, face, uL, uR), pw.ldfss(oracle, cfg, face, uL, uR)
    scale, pscale = flux_scales(cfg, face, uL, uR)
    assert (np.abs(fg - fo) / scale).max() < 2e-15
    assert (np.abs(pg - po) / pscale).max() < 2e-15
    res = rng.normal(0, 1e3, (len(uL), 4))
    dty = np.stack([rng.uniform(1e-9, 1e-6, len(uL)), rng.uniform(1e-4, 1.0, len(uL))], axis=1)
    vg, vo = pw.solve(fast, cfg, dty, uL, res), pw.solve(oracle, cfg, dty, uL, res)
    assert (np.abs(vg - vo) / np.abs(uL)).max() < 1e-12


def test_fast_reciprocals_are_rounding_limited(hp):
    fast, _ = hp
    rng = np.random.default_rng(5)
    x = 10.0 ** rng.uniform(-12, 12, 200000) * rng.uniform(1, 2, 200000)
    out = np.zeros((len(x), 3))
    dp = C.POINTER(C.c_double)
    fast.test_fastmath(len(x), x.ctypes.data_as(dp), out.ctypes.data_as(dp))
    xl = x.astype(np.longdouble)
    ref = np.stack([1 / xl, 1 / np.sqrt(xl), np.sqrt(xl)], axis=1)
    err = np.abs((out.astype(np.longdouble) - ref) / ref).max(axis=0)
    assert (err < 2.5e-16).all(), err


def test_cell_volume_is_the_reference_rounding(hp, oracle):
    """MeshModule.cpp:34-40 cancels to ~1e-9 on fine meshes; both modes must reproduce ITS rounding."""
    from par2d_b200 import Case
    from conftest import case_path
    case = Case.load(case_path("Case_spikefine24"))
    d = case.descs()[5]
    nx, ny = d.nx, d.ny
    x = np.ctypeslib.as_array(d.x, (ny, nx)); y = np.ctypeslib.as_array(d.y, (ny, nx))
    q = np.stack([x[:-1, :-1], y[:-1, :-1], x[:-1, 1:], y[:-1, 1:], x[1:, 1:], y[1:, 1:], x[1:, :-1], y[1:, :-1]], axis=-1).reshape(-1, 8).copy()
    x1, y1, x2, y2, x3, y3, x4, y4 = q.T
    ref = np.abs(0.5 * (x1 * y2 + x2 * y3 + x3 * y4 + x4 * y1 - x2 * y1 - x3 * y2 - x4 * y3 - x1 * y4))   # numpy: same order, no FMA
    dp = C.POINTER(C.c_double)
    for lib in hp:
        vol = np.zeros(len(q))
        lib.test_volume(len(q), q.ctypes.data_as(dp), vol.ctypes.data_as(dp))
        assert np.array_equal(vol, ref)


@pytest.mark.parametrize("iaxi,igam", [(1, 1.4), (0, 1.4), (1, -1.4)])
def test_early_upwind_vote_is_the_general_flux_to_rounding(hplib, oracle, iaxi, igam):
    """physics.cuh ldfss_early / ldfss_upwind_left: on the device a face whose warp is supersonic from the left on all
    32 faces takes the fully-upwind flux, evaluated without the rsqrt; every other face the general lean flux.  Here,
    per face and on the host: the test holds exactly where both face Mach numbers are >= 1 (up to rounding at M = 1),
    and where it holds the two fluxes agree to rounding - so which faces share a warp cannot matter beyond 1e-15."""
    from par2d_b200._capi import Config
    rng = np.random.default_rng(77)
    n = 40000
    face = random_faces(rng, n)
    uL, uR = random_states(rng, n), random_states(rng, n)
    # half of the faces: both sides flowing along +n at Mach 0.6..4 (the case the vote is for), states close to each other
    k = n // 2
    a = np.sqrt(abs(igam) * 287.0 * uL[:k, 3])
    mach = rng.uniform(0.6, 4.0, k)
    tang = rng.uniform(-300, 300, k)
    uL[:k, 1] = mach * a * face[:k, 0] - tang * face[:k, 1]
    uL[:k, 2] = mach * a * face[:k, 1] + tang * face[:k, 0]
    uR[:k] = uL[:k] * rng.uniform(0.8, 1.25, (k, 4))
    cfg = make_cfg(oracle, iaxi, igam)
    out = np.zeros((n, 11))
    dp = C.POINTER(C.c_double)
    fn = hplib.hp0_test_ldfss_upwind
    fn.argtypes, fn.restype = [C.POINTER(Config), C.c_int64, dp, dp, dp, dp], C.c_int
    assert fn(C.byref(cfg), n, *(np.ascontiguousarray(x).ctypes.data_as(dp) for x in (face, uL, uR)), out.ctypes.data_as(dp)) == 0
    up = out[:, 10] == 1.0
    # the test against the face Mach numbers of EulerFlux.cpp:134-137
    T = np.maximum(np.stack([uL[:, 3], uR[:, 3]]), 201.0)
    ahalf = 0.5 * np.sqrt(abs(igam) * cfg.Rs * T).sum(axis=0)
    xml = (uL[:, 1] * face[:, 0] + uL[:, 2] * face[:, 1]) / ahalf
    xmr = (uR[:, 1] * face[:, 0] + uR[:, 2] * face[:, 1]) / ahalf
    want = (xml >= 1) & (xmr >= 1)
    edge = (np.abs(xml - 1) < 1e-12) | (np.abs(xmr - 1) < 1e-12)
    assert np.array_equal(up[~edge], want[~edge])
    assert 0.2 * n < up.sum() < 0.5 * n        # the case is exercised
    scale, pscale = flux_scales(cfg, face, uL, uR)
    assert (np.abs(out[up, 0:4] - out[up, 5:9]) / scale[up]).max() < 2e-15
    assert (np.abs(out[up, 4] - out[up, 9]) / pscale[up]).max() < 2e-15
    # ... and the general flux from the early quantities is the lean flux of test_ldfss, bit for bit
    fo, po = pw.ldfss(_HostPhysics(hplib, 0), cfg, face, uL, uR)
    assert np.array_equal(out[:, 5:9], fo) and np.array_equal(out[:, 9], po)


@pytest.mark.parametrize("igam", [1.4, -1.4, 1.3])
def test_lean_far_field_ghost_is_the_reference_ghost_to_rounding(hplib, oracle, igam):
    """physics.cuh farfield_lean (the fast mode's translation unit): one logarithm + two exponentials and series
    reciprocals instead of the reference's two pow() and divisions (BoundaryConditions.cpp:150-190)."""
    lean = _HostPhysics(hplib, 2)
    rng = np.random.default_rng(99)
    n = 60000
    face = random_faces(rng, n)
    ui = random_states(rng, n)
    ui[::2, 1:3] *= 0.05                       # subsonic normal Mach numbers of both signs on half of the faces
    btype = np.where(rng.uniform(size=n) < 0.8, 1, rng.choice([-1, 0, 2, 3, 4, 5], n)).astype(np.int32)
    cfg = make_cfg(oracle, 1, igam)
    gl = pw.boundary_state(lean, cfg, btype, face[:, :2], ui)
    go = pw.boundary_state(oracle, cfg, btype, face[:, :2], ui)
    assert np.isfinite(go).all()
    # which of the four far-field cases a face takes may differ at |Mn| = 1 and Mn = 0 exactly; nowhere else
    a = np.sqrt(abs(igam) * cfg.Rs * np.maximum(ui[:, 3], 201.0))
    mn = (ui[:, 1] * face[:, 0] + ui[:, 2] * face[:, 1]) / a
    ok = (np.abs(np.abs(mn) - 1) > 1e-12) & (np.abs(mn) > 1e-300)
    assert ok.sum() > 0.99 * n
    sub = ok & (btype == 1) & (np.abs(mn) < 1)
    assert sub.sum() > 0.3 * n                 # the subsonic far field is exercised
    assert np.abs(gl[ok][:, [0, 3]] / go[ok][:, [0, 3]] - 1).max() < 1e-13          # density, temperature: element-wise
    speed = np.maximum(np.hypot(ui[:, 1], ui[:, 2]), a)                               # velocity components cancel: per-face scale
    assert (np.abs(gl[ok][:, 1:3] - go[ok][:, 1:3]) / speed[ok, None]).max() < 1e-14
    # every other boundary code is untouched: the same bits as the fast formulas without PAR2D_LEAN_BC
    gf = pw.boundary_state(_HostPhysics(hplib, 0), cfg, btype, face[:, :2], ui)
    assert np.array_equal(gl[btype != 1], gf[btype != 1])
