"""oracle/run_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Runs the UNMODIFIED reference binary built by oracle/build_ref.sh
(oracle/_ref/MPI2DFV) on a case directory for a fixed number of iterations and
returns its outputs: the residual history (../Outputs/res.tec, main.cpp:495),
the printed dt per iteration (main.cpp:499) and the final cell states parsed
from ../Outputs/solution.b<N>.tec (FileIO.cpp:84-115, %.16e so the doubles
round-trip exactly).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "MPI2DFV")
REF_LIB = os.path.join(HERE, "_ref", "libpar2d_ref.so")


def available() -> bool:
    return os.path.exists(REF_BIN)


def _stage(case_dir: str, therm_path: str, nblocks: int, zero_based_expected: bool = True) -> str:
    """Build the directory layout the reference hard-codes: <run>/build (cwd),
    <run>/Case, <run>/Outputs, <run>/therm_air.dat.  The binary opens
    grid.block<rank>.dat (0-based, FileIO.cpp:17); 1-based case directories
    (Case_ffstep1/8) are renamed and their neighbour ids shifted accordingly."""
    run = tempfile.mkdtemp(prefix="par2d_ref_")
    os.makedirs(os.path.join(run, "build"))
    os.makedirs(os.path.join(run, "Outputs"))
    shutil.copy(therm_path, os.path.join(run, "therm_air.dat"))
    dst = os.path.join(run, "Case")
    os.makedirs(os.path.join(dst, "subgrids"))
    sub = os.path.join(case_dir, "subgrids")
    one_based = not os.path.exists(os.path.join(sub, "grid.block0.dat")) and \
        os.path.exists(os.path.join(sub, f"grid.block{nblocks}.dat"))
    for b in range(nblocks):
        shutil.copy(os.path.join(sub, f"grid.block{b + (1 if one_based else 0)}.dat"),
                    os.path.join(dst, "subgrids", f"grid.block{b}.dat"))
    with open(os.path.join(case_dir, "grid.conn")) as f:
        text = f.read()
    if one_based:
        def shift(m):
            bc, nb = int(m.group(1)), int(m.group(2))
            return f"({bc},{nb - 1 if nb >= 0 else nb})"
        text = re.sub(r"\((-?\d+),(-?\d+)\)", shift, text)
    with open(os.path.join(dst, "grid.conn"), "w") as f:
        f.write(text)
    return run


def read_solution(path: str, nx: int, ny: int) -> np.ndarray:
    data = np.loadtxt(path, delimiter=",", skiprows=3, dtype=np.float64)
    assert data.shape == ((nx - 1) * (ny - 1), 9), (data.shape, nx, ny)
    return data[:, 2:6].reshape(ny - 1, nx - 1, 4).copy()


def read_nodal_dgp1(path: str, nx: int, ny: int) -> np.ndarray:
    """print_state_DGP1 file (FileIO.cpp:117-217): rho,u,v,T reconstructed at the nx*ny NODES."""
    data = np.loadtxt(path, delimiter=",", skiprows=3, dtype=np.float64)
    assert data.shape == (nx * ny, 9), (data.shape, nx, ny)
    return data[:, 2:6].reshape(ny, nx, 4).copy()


def run(case_dir: str, therm_path: str, shapes, niter: int, iaxi: int = 1, igam: float = 1.4,
        keep: bool = False, timeout: float = 3600.0, accur: int = 0):
    """Returns dict(relres=[...], dt=[...], states=[(ny-1,nx-1,4) per block], wall=seconds).
    accur=1 (the second-order branch): states are replaced by `nodal`, the node-averaged reconstruction
    print_state_DGP1 writes to ../Outputs/time/solution.b%03d.it%05d.tec (ny, nx, 4 per block)."""
    nblocks = len(shapes)
    run_dir = _stage(case_dir, therm_path, nblocks)
    env = dict(os.environ, SHIM_NP=str(nblocks), ORACLE_NITER=str(niter), ORACLE_IAXI=str(iaxi),
               ORACLE_IGAM=repr(float(igam)), ORACLE_ACCUR=str(int(accur)))
    t0 = time.perf_counter()
    proc = subprocess.run([REF_BIN], cwd=os.path.join(run_dir, "build"), env=env, capture_output=True,
                          text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    if proc.returncode != 0:
        raise RuntimeError(f"reference run failed rc={proc.returncode}: {proc.stdout[-2000:]} {proc.stderr[-2000:]}")
    # only rank 0 prints the iteration line, but the ranks share one pipe and their start-up chatter
    # can land in the middle of it: accept complete %7.4e fields only
    dts = [float(m.group(1)) for m in re.finditer(r"^Iter:\s*\d+\s+dt:\s*(\d\.\d{4}e[+-]\d{2,3})(?=\s)", proc.stdout, re.M)]
    relres = []
    with open(os.path.join(run_dir, "Outputs", "res.tec")) as f:
        next(f)
        for line in f:
            relres.append(float(line.split(",")[1]))
    nodal = None
    if accur == 1:
        states = None
        nodal = [read_nodal_dgp1(os.path.join(run_dir, "Outputs", "time", "solution.b%03d.it%05d.tec" % (b + 1, niter)), nx, ny)
                 for b, (nx, ny) in enumerate(shapes)]
    else:
        states = [read_solution(os.path.join(run_dir, "Outputs", f"solution.b{b + 1}.tec"), nx, ny)
                  for b, (nx, ny) in enumerate(shapes)]
    failed = "Failed Assert" in proc.stdout
    out = dict(relres=np.array(relres), dt=np.array(dts), states=states, nodal=nodal, wall=wall, failed=failed,
               stdout_tail=proc.stdout[-500:])
    if keep:
        out["run_dir"] = run_dir
    else:
        shutil.rmtree(run_dir, ignore_errors=True)
    return out
