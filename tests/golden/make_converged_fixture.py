"""Generates tests/golden/ref_committed_solution_ffstep8.npz from the reference's OWN committed outputs
(no reference code is run): FV/Outputs/solution.b1..8.tec = the converged state of the 8-block forward
step at the shipped macros after 497 227 iterations, and the tail of FV/Outputs/res.tec.
Run in the build container (needs /root/reference):  python tests/golden/make_converged_fixture.py
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from par2d_b200 import Case  # noqa: E402
from par2d_b200.host import read_solution  # noqa: E402

REF = os.environ.get("PAR2D_REFERENCE", "/root/reference")
case = Case.load(os.path.join(REPO, "tests", "golden", "cases", "Case_ffstep8"))
out = {}
for b in range(case.nblocks):
    nx, ny = case.shape(b)
    # rank b writes solution.b<b+1>.tec (FileIO.cpp:84)
    unk, xc, yc = read_solution(os.path.join(REF, "FV", "Outputs", f"solution.b{b + 1}.tec"), nx, ny)
    out[f"unk{b}"] = unk
    out[f"xc{b}"] = xc[0, :3].copy()      # three centroids per block pin the block <-> file association
    out[f"yc{b}"] = yc[0, :3].copy()
with open(os.path.join(REF, "FV", "Outputs", "res.tec")) as f:
    tail = f.readlines()[-3:]
out["res_tail_iter"] = np.array([int(line.split(",")[0]) for line in tail])
out["res_tail_relres"] = np.array([float(line.split(",")[1]) for line in tail])
path = os.path.join(REPO, "tests", "golden", "ref_committed_solution_ffstep8.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes;", "res.tec tail:", [line.strip() for line in tail])
