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
    # print_state_DGP1's layout: <out>/time/solution.b%03d.it%05d.tec, initial state labelled 0 (FileIO.cpp:121-128)
    assert os.path.exists(tmp_path / "dg" / "time" / "solution.b001.it00000.tec")
    assert os.path.exists(tmp_path / "dg" / "time" / "solution.b001.it00030.tec") and "RelativeTotalResisual" in out
    with open(tmp_path / "dg" / "time" / "solution.b003.it00030.tec") as f:
        assert f.readlines()[2].startswith("ZONE I=101, J=71, DATAPACKING=POINT, STRANDID=30, SOLUTIONTIME=")


def test_driver_periodic_saves(tmp_path):
    """saveiter > 0 with a batch size that divides it (the shipped defaults are 100 / 1000): the reference saves after
    every iteration m > 0 with m % saveiter == 0 (main.cpp:484), i.e. the state after m + 1 steps, and labels the
    second-order branch's files with m (they are all kept: FileIO.cpp:127)."""
    case = case_path("Case_ffstep8")
    out = _run(tmp_path / "d", "--case", case, "--niter", "25", "--accur", "1", "--saveiter", "10", "--batch", "5", "--printiter", "7")
    assert out.count("RelativeTotalResisual") == 4          # iterations 0, 7, 14, 21
    names = sorted(os.listdir(tmp_path / "d" / "time"))
    # labels: 0 (initial), 10 and 20 (periodic), 25 (final: the loop variable's exit value, main.cpp:541-545)
    assert [n for n in names if n.startswith("solution.b001.")] == [
        "solution.b001.it00000.tec", "solution.b001.it00010.tec", "solution.b001.it00020.tec", "solution.b001.it00025.tec"]
    # the file labelled 10 holds the state after 11 steps
    _run(tmp_path / "e", "--case", case, "--niter", "11", "--accur", "1", "--saveiter", "0", "--batch", "11")
    with open(tmp_path / "d" / "time" / "solution.b002.it00010.tec") as f:
        per = f.readlines()[3:]
    with open(tmp_path / "e" / "time" / "solution.b002.it00011.tec") as f:
        fin = f.readlines()[3:]
    assert per == fin
    # first-order files are overwritten in place (FileIO.cpp:84): trimming the batches at the save points must not
    # change where the run ends up
    _run(tmp_path / "p", "--case", case, "--niter", "23", "--faithful", "--saveiter", "10", "--batch", "5")
    _run(tmp_path / "q", "--case", case, "--niter", "23", "--faithful", "--saveiter", "0", "--batch", "23")
    c = Case.load(case)
    nx, ny = c.shape(2)
    a, _, _ = read_solution(str(tmp_path / "p" / "solution.b3.tec"), nx, ny)
    b, _, _ = read_solution(str(tmp_path / "q" / "solution.b3.tec"), nx, ny)
    assert np.array_equal(a, b)
    with open(tmp_path / "p" / "res.tec") as f, open(tmp_path / "q" / "res.tec") as g:
        assert f.read() == g.read()


def test_driver_multi_context(tmp_path):
    """par2d_fv --gpus N: one process, N contexts (sharing the GPUs present), blocks partitioned, peer-mapped halo
    exchange, residual gathered in global block order, tolerance stop decided across contexts: res.tec and every
    solution file byte-identical to the one-context run."""
    case = case_path("Case_spikefine24")
    _run(tmp_path / "one", "--case", case, "--niter", "60", "--saveiter", "0", "--batch", "25")
    for n in (2, 3):
        out = _run(tmp_path / f"m{n}", "--case", case, "--niter", "60", "--saveiter", "0", "--batch", "25", "--gpus", str(n))
        assert f"24 blocks on {n} contexts" in out
        for name in ["res.tec"] + [f"solution.b{b + 1}.tec" for b in range(24)]:
            with open(tmp_path / "one" / name, "rb") as f, open(tmp_path / f"m{n}" / name, "rb") as g:
                assert f.read() == g.read(), (n, name)
    # tolerance break in the middle of a batch on 2 contexts == on one
    case = case_path("Case_ffstep8")
    _run(tmp_path / "t1", "--case", case, "--niter", "40", "--saveiter", "0", "--tol", "0.9996", "--batch", "16")
    _run(tmp_path / "t2", "--case", case, "--niter", "40", "--saveiter", "0", "--tol", "0.9996", "--batch", "16", "--gpus", "2")
    with open(tmp_path / "t1" / "res.tec") as f:
        lines = f.read().splitlines()
    assert len(lines) == 1 + 3          # header + iterations 0, 1, 2
    for name in ["res.tec"] + [f"solution.b{b + 1}.tec" for b in range(8)]:
        with open(tmp_path / "t1" / name, "rb") as f, open(tmp_path / "t2" / name, "rb") as g:
            assert f.read() == g.read(), name
