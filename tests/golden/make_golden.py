"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.

Runs the UNMODIFIED reference (oracle/_ref/MPI2DFV, built by oracle/build_ref.sh from
/root/reference/FV/*.cpp against the MPI shim) on the case fixtures under
tests/golden/cases/ (verbatim copies of the reference's FV/Case_* data directories) and
stores, per (case, variant):

  golden_<case>_<variant>.npz
      relres   the residual history written to res.tec (main.cpp:495, %le -> 7 digits)
      dt       the time step printed per iteration (main.cpp:499, %7.4e -> 5 digits)
      b<k>     final state of block k, (ny-1, nx-1, 4) float64, parsed from
               solution.b<k+1>.tec (FileIO.cpp:111, %.16e -> exact round trip);
               stored only for the variants listed in FULL_STATE to keep the repo small
      sum<k>   per-block, per-variable sums of the final state (all variants)

ref_committed_res_ffstep8.tec is NOT generated here: it is the first 2001 lines of the
reference's own committed FV/Outputs/res.tec (the only golden the reference ships).

Only runs where /root/reference and oracle/_ref exist (the build container).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import run_ref  # noqa: E402
from par2d_b200 import Case  # noqa: E402

CASES = os.path.join(HERE, "cases")
THERM = os.path.join(CASES, "therm_air.dat")

# (case dir, variant name, iaxi, igam, steps)
RUNS = [
    ("Case_ffstep1", "axi_cpg", 1, 1.4, 60),
    ("Case_ffstep8", "axi_cpg", 1, 1.4, 200),
    ("Case_ffstep8", "planar_cpg", 0, 1.4, 100),
    ("Case_rktnoz9", "axi_cpg", 1, 1.4, 200),
    ("Case_rktnoz9", "axi_tpg", 1, -1.4, 200),
    ("Case_spikefine24", "axi_cpg", 1, 1.4, 200),
    # the reference's nested variant FV/Case_spikefine24/Case: BC 5 (Bernoulli inflow, BoundaryConditions.cpp:135-146)
    # on the right side of blocks 11 and 15; assembled from the base case + the overlay directory
    ("Case_spikefine24_bc5", "axi_cpg", 1, 1.4, 200),
]
FULL_STATE = {(r[0], r[1]) for r in RUNS}

# second-order branch (ACCUR = 1, DG(P1)): the reference's print_state_DGP1 writes the node-averaged
# reconstruction (FileIO.cpp:117-217), stored as n<k> (ny, nx, 4) per block
RUNS_DG = [
    ("Case_ffstep8", "dgp1_axi_cpg", 1, 1.4, 200),
    ("Case_rktnoz9", "dgp1_axi_tpg", 1, -1.4, 100),
]


def main():
    if not run_ref.available():
        raise SystemExit("oracle/_ref/MPI2DFV missing: run oracle/build_ref.sh first")
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for name, variant, iaxi, igam, steps in (RUNS if "--dg-only" not in sys.argv else []):
        if only and name not in only:
            continue
        case_dir = os.path.join(CASES, name)
        if name == "Case_spikefine24_bc5":
            import tempfile
            sys.path.insert(0, os.path.dirname(HERE))
            from conftest import overlay_case
            case_dir = overlay_case("Case_spikefine24", name, os.path.join(tempfile.mkdtemp(), "case"))
        case = Case.load(case_dir)
        shapes = [case.shape(b) for b in range(case.nblocks)]
        out = run_ref.run(case_dir, THERM, shapes, steps, iaxi=iaxi, igam=igam)
        assert not out["failed"], out["stdout_tail"]
        assert len(out["relres"]) == steps and len(out["dt"]) == steps
        data = dict(relres=out["relres"], dt=out["dt"], steps=np.int64(steps), iaxi=np.int64(iaxi), igam=np.float64(igam))
        for b, st in enumerate(out["states"]):
            data[f"sum{b}"] = st.reshape(-1, 4).sum(axis=0)
            if (name, variant) in FULL_STATE:
                data[f"b{b}"] = st
        path = os.path.join(HERE, f"golden_{name}_{variant}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {steps} steps, {os.path.getsize(path) / 1e6:.2f} MB, relres[-1]={out['relres'][-1]:.6e}")
    for name, variant, iaxi, igam, steps in (RUNS_DG if not only else []):
        case = Case.load(os.path.join(CASES, name))
        shapes = [case.shape(b) for b in range(case.nblocks)]
        out = run_ref.run(os.path.join(CASES, name), THERM, shapes, steps, iaxi=iaxi, igam=igam, accur=1)
        assert not out["failed"], out["stdout_tail"]
        assert len(out["relres"]) == steps and len(out["dt"]) == steps
        data = dict(relres=out["relres"], dt=out["dt"], steps=np.int64(steps), iaxi=np.int64(iaxi), igam=np.float64(igam))
        for b, nd in enumerate(out["nodal"]):
            data[f"n{b}"] = nd
        path = os.path.join(HERE, f"golden_{name}_{variant}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {steps} steps, {os.path.getsize(path) / 1e6:.2f} MB, relres[-1]={out['relres'][-1]:.6e}")


if __name__ == "__main__":
    main()
