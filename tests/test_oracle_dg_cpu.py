"""Pins the second-order branch of the oracle (oracle/par2d_oracle_dg.inc: the reference's ACCUR = 1,
DG(P1) path) BITWISE against
  (1) golden fixtures generated from the unmodified reference binary run with ACCUR = 1
      (tests/golden/make_golden.py RUNS_DG): residual history, dt, and the node-averaged reconstruction
      print_state_DGP1 writes (FileIO.cpp:117-217, %.16e) after N steps - 8-block forward step (halo of two
      Gauss points per face) and the 9-block nozzle with curve-fit thermodynamics (BCs 1-4);
  (2) the reference's calc_dudt called directly with random means AND slopes through
      oracle/_ref/libpar2d_ref.so (when present)."""
import ctypes as C
import os

import numpy as np
import pytest

import dg
from conftest import GOLDEN, THERM, case_path
from par2d_b200 import Case, Solver


@pytest.mark.parametrize("name,variant", dg.DG_RUNS)
def test_dg_port_matches_reference_golden_bitwise(oracle, name, variant):
    g, case, s, info = dg.run_golden(oracle, name, variant)
    assert len(info) == int(g["steps"]) and all(i.status == 0 for i in info)
    assert ["%le" % i.relres for i in info] == ["%le" % v for v in g["relres"]]
    assert ["%7.4e" % i.dt for i in info] == ["%7.4e" % v for v in g["dt"]]
    for b in range(case.nblocks):
        assert np.array_equal(s.nodal_dgp1(b), g[f"n{b}"]), f"block {b}: nodal reconstruction differs from the reference"
    # the slopes are alive (not limited to zero everywhere) - otherwise the test would only see first order
    ux, uy = s.get_slopes(0)
    assert np.abs(ux).max() > 0 and np.abs(uy).max() > 0


def test_dg_first_order_limit(oracle):
    """With accur = 1 the scheme is NOT the first-order one with zero slopes: the volume integral and the
    Gauss-point reconstruction act from step 0 on; but accur = 0 must stay untouched by the new field."""
    case = Case.load(case_path("Case_ffstep8"))
    a = Solver(case, dg.dg_cfg(oracle), lib=oracle).step(3)
    cfg0 = dg.dg_cfg(oracle)
    cfg0.accur = 0
    b = Solver(case, cfg0, lib=oracle).step(3)
    assert a[0].relres == b[0].relres == 1.0 and a[1].relres != b[1].relres


def test_dg_residual_vs_reference_function_bitwise(oracle, tmp_path):
    """calc_dudt (SpatialDiscretization.cpp:159 with ACCUR = 1) of the reference itself on a single block with
    random means and slopes: dudt, duxdt, duydt bit for bit."""
    path = os.path.abspath(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libpar2d_ref.so"))
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = C.CDLL(path)
    os.makedirs(tmp_path / "build")
    import shutil
    shutil.copy(THERM, tmp_path / "therm_air.dat")
    ref.ref_init.argtypes = [C.c_char_p, C.c_int, C.c_double]
    DP, IP = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    ref.ref_calc_dudt_dg.argtypes = [C.c_int, C.c_int, DP, DP, IP, DP] + [DP] * 6
    for iaxi, igam in ((1, 1.4), (0, 1.4), (1, -1.4)):
        assert ref.ref_init(str(tmp_path / "build").encode(), iaxi, C.c_double(igam)) == 0
        case = Case.load(case_path("Case_rktnoz9"))
        for blk in (0, 4):   # block 4 is 19 x 10 cells with three BC kinds on its perimeter
            x, y = case.nodes(blk)
            ny, nx = x.shape
            ib = np.ascontiguousarray(case.ibound(blk), dtype=np.int32)
            ib[ib < 0] = 3   # as a stand-alone block every face is a boundary face
            single = _single_block_case(tmp_path / f"c{iaxi}{igam}{blk}", x, y, ib)
            cfg = dg.dg_cfg(oracle, iaxi, igam)
            s = Solver(single, cfg, lib=oracle)
            dg.perturbed_slopes(s, single, seed=11 + blk)
            unk = s.get_state(0)
            ux, uy = s.get_slopes(0)
            mine = s.get_residual_dg(0)
            out = [np.zeros_like(unk) for _ in range(3)]
            uFS = np.array(list(cfg.uFS))
            xs, ys = single.nodes(0)          # what the oracle holds (the 1e-4 shift need not round-trip exactly)
            u2, x2, y2 = unk.copy(), np.ascontiguousarray(xs), np.ascontiguousarray(ys)
            ref.ref_calc_dudt_dg(nx, ny, x2.ctypes.data_as(DP), y2.ctypes.data_as(DP), ib.ctypes.data_as(IP),
                                 uFS.ctypes.data_as(DP), u2.ctypes.data_as(DP), ux.ctypes.data_as(DP), uy.ctypes.data_as(DP),
                                 *[o.ctypes.data_as(DP) for o in out])
            for name, a, b in zip(("dudt", "duxdt", "duydt"), mine, out):
                assert np.array_equal(a, b), f"{name} differs (iaxi={iaxi} igam={igam} block={blk}): {np.abs(a - b).max()}"


def _single_block_case(directory, x, y, ibound):
    """Writes a one-block case directory in the reference format (FileIO.cpp:13-51) and loads it."""
    os.makedirs(directory / "subgrids")
    ny, nx = x.shape
    with open(directory / "grid.conn", "w") as f:
        f.write("1\n0: (3,-1) (3,-1) (3,-1) (3,-1)\n")
    with open(directory / "subgrids" / "grid.block0.dat", "w") as f:
        f.write("npart: 1 1\nipart: 0 0\nngrid: %d %d\n" % (nx, ny))
        for j in range(ny):
            for i in range(nx):
                f.write("%.17e %.17e\n" % (x[j, i], y[j, i] - 1e-4))   # the loader adds 1e-4 to y (FileIO.cpp:41)
        f.write("BC\n")
        for c in ibound:
            f.write("%d\n" % c)
    return Case.load(str(directory))


def test_dgp1_writer_reproduces_the_references_nodal_file(oracle, tmp_path):
    """par2d_write_solution_dgp1 (print_state_DGP1, FileIO.cpp:117-217): the values it writes for the state the
    port reaches after the golden's N steps are the reference file's, digit for digit (%.16e)."""
    import ctypes as C
    from par2d_b200 import _capi
    g, case, s, info = dg.run_golden(oracle, "Case_rktnoz9", "dgp1_axi_tpg")
    host = _capi.load_host()
    dp = C.POINTER(C.c_double)
    for b in (0, 4, 8):
        nx, ny = case.shape(b)
        x, y = case.nodes(b)
        unk = s.get_state(b)
        ux, uy = s.get_slopes(b)
        path = str(tmp_path / ("solution.b%03d.it%05d.tec" % (b + 1, len(info))))
        host.check(host.par2d_write_solution_dgp1(path.encode(), b"Final State", C.byref(s.cfg), nx, ny, x.ctypes.data_as(dp),
                                                  y.ctypes.data_as(dp), unk.ctypes.data_as(dp), ux.ctypes.data_as(dp),
                                                  uy.ctypes.data_as(dp), len(info), sum(i.dt for i in info)), "write_solution_dgp1")
        with open(path) as f:
            head = [next(f) for _ in range(3)]
            data = np.loadtxt(f, delimiter=",")
        assert head[2].startswith("ZONE I=%d, J=%d, DATAPACKING=POINT, STRANDID=%d, SOLUTIONTIME=" % (nx, ny, len(info)))
        assert np.array_equal(data[:, 2:6].reshape(ny, nx, 4), g[f"n{b}"])
        assert np.array_equal(data[:, 0].reshape(ny, nx), x) and np.array_equal(data[:, 1].reshape(ny, nx), y)
