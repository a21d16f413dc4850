"""The C++ host driver par2d_fv (par2d_b200/csrc/host/main.cpp: the reference's main.cpp re-expressed over the C ABI)
run as a user would: case directory in, res.tec / solution.b<N>.tec out, byte formats of the reference."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REPO, THERM, case_path, load_golden
from par2d_b200.host import read_solution
from par2d_b200 import Case

pytestmark = pytest.mark.gpu

BIN = os.path.join(REPO, "par2d_b200", "par2d_fv")


def _run(out, *args):
    os.makedirs(out, exist_ok=True)
    p = subprocess.run([BIN, "--out", str(out), "--therm", THERM] + list(args), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def test_driver_reproduces_the_references_committed_history(tmp_path):
    """300 iterations of the 8-block forward step in the bit-faithful mode: res.tec equals the first 301 lines of the
    reference's own committed FV/Outputs/res.tec, the stdout lines have the reference's format, and the solution
    files equal the golden states written by the reference binary after 200 steps (run 2)."""
    out = _run(tmp_path / "a", "--case", case_path("Case_ffstep8"), "--niter", "300", "--faithful", "--saveiter", "0")
    with open(os.path.join(GOLDEN, "ref_committed_res_ffstep8.tec")) as f:
        want = f.read().splitlines()[:301]
    with open(tmp_path / "a" / "res.tec") as f:
        got = f.read().splitlines()
    assert got == want
    assert "Iter:      1\tdt:2.7754e-08 \t\t RelativeTotalResisual:  9.99746e-01 " in out
    g = load_golden("Case_ffstep8", "axi_cpg")
    _run(tmp_path / "b", "--case", case_path("Case_ffstep8"), "--niter", str(int(g["steps"])), "--faithful", "--saveiter", "0")
    case = Case.load(case_path("Case_ffstep8"))
    for b in range(case.nblocks):
        nx, ny = case.shape(b)
        unk, _, _ = read_solution(str(tmp_path / "b" / f"solution.b{b + 1}.tec"), nx, ny)
        assert np.array_equal(unk, g[f"b{b}"]), f"block {b}"


def test_driver_restart_and_second_order(tmp_path):
    """--restart continues from written solution files exactly where a longer run would be (the state is the whole
    restart information in the first-order scheme except res0, so the residual history is compared through its ratio);
    --accur 1 writes print_state_DGP1's files."""
    case = case_path("Case_spikefine24")
    _run(tmp_path / "full", "--case", case, "--niter", "40", "--faithful", "--saveiter", "0")
    _run(tmp_path / "half", "--case", case, "--niter", "20", "--faithful", "--saveiter", "0")
    _run(tmp_path / "rest", "--case", case, "--niter", "20", "--faithful", "--saveiter", "0", "--restart", str(tmp_path / "half"))
    c = Case.load(case)
    for b in range(c.nblocks):
        nx, ny = c.shape(b)
        a, _, _ = read_solution(str(tmp_path / "full" / f"solution.b{b + 1}.tec"), nx, ny)
        r, _, _ = read_solution(str(tmp_path / "rest" / f"solution.b{b + 1}.tec"), nx, ny)
        assert np.array_equal(a, r), f"block {b}: restart + 20 differs from 40 straight"
    out = _run(tmp_path / "dg", "--case", case_path("Case_ffstep8"), "--niter", "30", "--accur", "1", "--saveiter", "0")
    assert os.path.exists(tmp_path / "dg" / "solution.b001.it00030.tec") and "RelativeTotalResisual" in out
    with open(tmp_path / "dg" / "solution.b003.it00030.tec") as f:
        assert f.readlines()[2].startswith("ZONE I=101, J=71, DATAPACKING=POINT, STRANDID=30, SOLUTIONTIME=")
