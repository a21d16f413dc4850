"""Diagnostic: a few steps of a shipped case through the C ABI, printing the state error vs the oracle.
Usage: python tools/diag_step.py [case] [steps] [mode] [verbose]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from par2d_b200 import Case, Solver, _capi, default_config

name = sys.argv[1] if len(sys.argv) > 1 else "Case_ffstep8"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
verbose = len(sys.argv) > 4
gpu = _capi.load_gpu()
if name.startswith("ramp:"):
    _, cx, cy, bx, by = name.split(":")
    case = Case.synthetic_ramp(int(cx), int(cy), int(bx), int(by))
else:
    case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
cfg = default_config(gpu); cfg.device = 0; cfg.mode = mode
dev = Solver(case, cfg, lib=gpu)
orc = Solver(case, cfg, lib=_capi.load_oracle())
if verbose:
    for b in range(case.nblocks):
        rd, ro = dev.get_residual(b), orc.get_residual(b)
        e = np.abs(rd - ro)
        j, i, k = np.unravel_index(np.argmax(e), e.shape)
        print(f"  residual block {b} shape {ro.shape} max|d|={e.max():.3e} at j={j} i={i} k={k} dev={rd[j,i,k]:.6e} orc={ro[j,i,k]:.6e}")
info = dev.step(steps, check=False)
ref = orc.step(steps, check=False)
bad = [i.iter for i in info if i.status != 0]
if bad or len(info) != len(ref):
    print(f"  device status != 0 (status {info[-1].status}); history dev vs orc:")
    for a, b in list(zip(info, ref))[:8]:
        print(f"    it {a.iter} dev relres {a.relres:.9e} dt {a.dt:.6e} | orc relres {b.relres:.9e} dt {b.dt:.6e}")
worst = 0.0
for b in range(case.nblocks):
    s, r = dev.get_state(b), orc.get_state(b)
    for k in range(4):
        e = np.abs(s[..., k] - r[..., k])
        scale = np.sqrt(r[..., 1] ** 2 + r[..., 2] ** 2).max() if k in (1, 2) else np.abs(r[..., k]).max()
        w = e.max() / max(scale, 1e-300)
        worst = max(worst, w)
        if verbose and w > 1e-11:
            j, i = np.unravel_index(np.argmax(e), e.shape)
            print(f"  state block {b} var {k} rel={w:.3e} at j={j} i={i} dev={s[j,i,k]:.9e} orc={r[j,i,k]:.9e} nbad={(e > 1e-9 * np.abs(r[..., k]).max()).sum()}")
print(f"{name} steps={steps} mode={mode} tma=on "
      f"worst={worst:.3e} relres={info[-1].relres:.9e} ref={ref[-1].relres:.9e} dt={info[-1].dt:.6e}")
