"""Diagnostic: fast path stepped continuously (a) one step per call, (b) in one batch; reports the first bad step."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from par2d_b200 import Case, Solver, _capi, default_config
gpu = _capi.load_gpu()
_, cx, cy, bx, by = sys.argv[1].split(":")
steps = int(sys.argv[2])
case = Case.synthetic_ramp(int(cx), int(cy), int(bx), int(by))
def mk(mode):
    c = default_config(gpu); c.device = 0; c.mode = mode
    return Solver(case, c, lib=gpu)
a, s = mk(0), mk(1)
for it in range(steps):
    ia, is_ = a.step(1, check=False), s.step(1, check=False)
    fa = a.get_state(0); fs = s.get_state(0)
    nan = np.isnan(fa).any(axis=-1)
    d = np.abs(fa - fs)[..., 3].max()
    if ia[0].status or nan.any() or d > 1e-6:
        jj, ii = np.nonzero(nan | (np.abs(fa - fs)[..., 3] > 1e-6))
        print(f"one-step-per-call: first bad step {it}: status {ia[0].status} nan cells {nan.sum()} maxdT {d:.3e} first at j={jj[:5]} i={ii[:5]}")
        j, i = jj[0], ii[0]
        print("   fast", fa[j, i], "slow", fs[j, i])
        break
else:
    print(f"one-step-per-call: {steps} steps fine, max dT {d:.3e}")
for n in (2, 3, 4, 8, 16, 32, 64, 100):
    b = mk(0)
    info = b.step(n, check=False)
    st = [i.status for i in info]
    fb = b.get_state(0)
    print(f"batch of {n}: status {st[-1]} nan cells {np.isnan(fb).any(axis=-1).sum()} T201 cells {(fb[..., 3] == 201.0).sum()} relres {info[-1].relres:.6e}")
