"""Diagnostic: step the fast and the faithful device paths side by side and report the first step / cell
where they part by more than `thr` (relative to the variable scale).  Usage: diag_fast_vs_faithful.py case steps [thr]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from par2d_b200 import Case, Solver, _capi, default_config

name = sys.argv[1]
steps = int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-9
gpu = _capi.load_gpu()
if name.startswith("ramp:"):
    _, cx, cy, bx, by = name.split(":")
    case = Case.synthetic_ramp(int(cx), int(cy), int(bx), int(by))
else:
    case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
cfgs = []
for mode in (0, 1):
    c = default_config(gpu); c.device = 0; c.mode = mode
    cfgs.append(c)
fast, slow = Solver(case, cfgs[0], lib=gpu), Solver(case, cfgs[1], lib=gpu)
np.set_printoptions(precision=12, linewidth=200)
for it in range(steps):
    prev = [slow.get_state(b) for b in range(case.nblocks)]
    for b in range(case.nblocks):
        fast.set_state(b, prev[b])          # same input state for both: isolates one step
    rf = [fast.get_residual(b) for b in range(case.nblocks)]
    rs = [slow.get_residual(b) for b in range(case.nblocks)]
    fi, si = fast.step(1, check=False), slow.step(1, check=False)
    worst, where = 0.0, None
    for b in range(case.nblocks):
        f, s = fast.get_state(b), slow.get_state(b)
        vel = np.sqrt(s[..., 1] ** 2 + s[..., 2] ** 2).max()
        sc = np.array([np.abs(s[..., 0]).max(), vel, vel, np.abs(s[..., 3]).max()])
        e = np.abs(f - s) / sc
        if e.max() > worst:
            worst = e.max(); j, i, k = np.unravel_index(np.argmax(e), e.shape); where = (b, j, i, k)
    if worst > thr or fi[0].status or si[0].status:
        b, j, i, k = where
        print(f"step {it}: worst {worst:.3e} at block {b} j={j} i={i} var {k}; status fast {fi[0].status} slow {si[0].status}")
        print("  input state (3x3 neighbourhood rows j-1..j+1, cols i-1..i+1):")
        p = prev[b]
        for jj in range(max(j - 1, 0), min(j + 2, p.shape[0])):
            for ii in range(max(i - 1, 0), min(i + 2, p.shape[1])):
                print(f"    j={jj} i={ii} u={p[jj, ii]!r}")
        print(f"  residual fast {rf[b][j, i]!r}\n  residual slow {rs[b][j, i]!r}")
        print(f"  new fast {fast.get_state(b)[j, i]!r}\n  new slow {slow.get_state(b)[j, i]!r}")
        er = np.abs(rf[b] - rs[b]); jj, ii, kk = np.unravel_index(np.argmax(er), er.shape)
        print(f"  largest residual diff in that block at j={jj} i={ii} k={kk}: fast {rf[b][jj, ii]!r} slow {rs[b][jj, ii]!r}")
        break
else:
    print(f"{name}: {steps} single steps, fast == faithful within {thr:g} (last worst {worst:.3e})")
