"""Diagnostic: N contexts on one GPU, optionally with the stop test across contexts; prints per-block sums and the
exchange snapshot of every context next to a single-context run."""
import ctypes as C, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from par2d_b200 import Case, Solver, _capi, default_config
from par2d_b200.local import LocalMultiSolver
from par2d_b200._capi import StepInfo
gpu = _capi.load_gpu()
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 0.9996
stop = int(sys.argv[4]) if len(sys.argv) > 4 else 1
name = sys.argv[5] if len(sys.argv) > 5 else "Case_ffstep8"
case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
cfg = default_config(gpu); cfg.device = 0; cfg.tol = tol
one = Solver(case, cfg, lib=gpu)
oi = one.step(nsteps)
print("single", [round(i.relres, 6) for i in oi])
print("single blk sums it0", one.block_history(nsteps)[0, :, 0].round(1).tolist())
m = LocalMultiSolver(case, cfg, nranks=nr, lib=gpu)
print("owners", m.owners)
for s in m.solvers:
    rc = gpu.step_async(s.handle, nsteps, stop); print("async", rc)
for r, s in enumerate(m.solvers):
    info = (StepInfo * nsteps)()
    rc = gpu.step_wait(s.handle, info)
    print("rank", r, "wait rc", rc, gpu.last_error(), [(i.iter, round(i.relres, 6), i.status) for i in info])
    fl = (C.c_uint64 * (16 * nr))(); res = (C.c_double * (8 * case.nblocks))(); cnt = (C.c_int64 * 6)()
    gpu.peer_snapshot(s.handle, fl, res, cnt)
    print("  flags", [list(fl[16 * q:16 * q + 2]) + [fl[16 * q + 15]] for q in range(nr)], "counters", list(cnt))
    n = max(sum(1 for i in info if i.iter >= 0), 1)
    print("  blk sums it0", dict(zip(s.owned, s.block_history(n)[0, :, 0].round(1).tolist())))
worst = max(np.abs(m.get_state(b) - one.get_state(b)).max() for b in range(case.nblocks))
print("max state diff vs single:", worst)
