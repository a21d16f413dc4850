"""Worker for the multi-rank tests (launched by torch.distributed.run).

  --backend gloo --impl oracle : host-side exchange/partition logic on CPU (world_size 2), the
                                 oracle standing in for the device library (test infrastructure)
  --backend nccl --impl gpu    : the real thing, one rank per GPU
Every rank advances its blocks; each rank then checks the blocks it owns against a
single-process run of the oracle on the same case: states and the residual history must be
identical (bitwise for the oracle backend, within tolerance for the CUDA path)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from par2d_b200 import Case, Solver, _capi, default_config  # noqa: E402
from par2d_b200.dist import DistSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--impl", default="oracle")
    ap.add_argument("--case", default="Case_spikefine24")
    ap.add_argument("--synthetic", default="")      # e.g. 128x128:4x4
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--transport", default="auto")
    ap.add_argument("--batches", type=int, default=1)   # the steps are taken in this many par2d_step batches
    ap.add_argument("--accur", type=int, default=0)     # 1: the second-order branch (lines carry two Gauss points per face)
    a = ap.parse_args()
    if a.backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group(a.backend)
    rank = dist.get_rank()
    lib = _capi.load_gpu() if a.impl == "gpu" else _capi.load_oracle()
    orc = _capi.load_oracle()

    def make_case():
        if a.synthetic:
            cells, blocks = a.synthetic.split(":")
            cx, cy = map(int, cells.split("x"))
            bx, by = map(int, blocks.split("x"))
            return Case.synthetic_ramp(cx, cy, bx, by)
        return Case.load(os.path.join(REPO, "tests", "golden", "cases", a.case))

    cfg = default_config(lib)
    cfg.accur = a.accur
    if a.impl == "gpu":
        cfg.device = torch.cuda.current_device()
    ds = DistSolver(make_case(), cfg, lib=lib, transport=a.transport)
    if a.transport in ("peer", "external"):
        assert ds.transport == a.transport, ds.transport
    per = a.steps // a.batches
    parts = [ds.step(per) for _ in range(a.batches)]
    dt = np.concatenate([p[0] for p in parts])
    relres = np.concatenate([p[1] for p in parts])

    ocfg = default_config(orc)
    ocfg.accur = a.accur
    single = Solver(make_case(), ocfg, lib=orc)
    info = single.step(a.steps)
    ref_rel = np.array([i.relres for i in info])
    ref_dt = np.array([i.dt for i in info])
    worst = 0.0
    for b in ds.owned:
        got, ref = ds.get_state(b), single.get_state(b)
        if a.tol == 0.0:
            assert np.array_equal(got, ref), f"rank {rank}: block {b} differs"
        vel = np.sqrt(ref[..., 1] ** 2 + ref[..., 2] ** 2).max()
        for k, scale in enumerate((np.abs(ref[..., 0]).max(), vel, vel, np.abs(ref[..., 3]).max())):
            worst = max(worst, np.abs(got[..., k] - ref[..., k]).max() / max(scale, 1e-300))
    if a.tol == 0.0:
        assert np.array_equal(relres, ref_rel) and np.array_equal(dt, ref_dt), (relres, ref_rel)
    else:
        assert worst <= a.tol, worst
        assert np.abs(relres / ref_rel - 1).max() <= 1e-9 and np.abs(dt / ref_dt - 1).max() <= 1e-12
    if a.impl == "gpu":
        # the same case advanced by ONE process on this rank's GPU: sharding over GPUs must not change a bit
        one = Solver(make_case(), cfg, lib=lib)
        oinfo = one.step(a.steps)
        for b in ds.owned:
            assert np.array_equal(ds.get_state(b), one.get_state(b)), f"rank {rank}: block {b} differs from the single-GPU run"
            if a.accur == 1:
                for x, y in zip(ds.solver.get_slopes(b), one.get_slopes(b)):
                    assert np.array_equal(x, y), f"rank {rank}: slopes of block {b} differ from the single-GPU run"
            if ds.transport == "peer":   # the arenas hold the final state's halos (trailing put of par2d_step)
                if a.accur == 1:
                    for x, y in zip(ds.solver.get_residual_dg(b), one.get_residual_dg(b)):
                        assert np.array_equal(x, y), f"rank {rank}: right-hand sides of block {b} differ"
                else:
                    assert np.array_equal(ds.solver.get_residual(b), one.get_residual(b)), f"rank {rank}: residual of block {b} differs"
        assert np.array_equal(dt, np.array([i.dt for i in oinfo]))
        assert np.array_equal(relres, np.array([i.relres for i in oinfo])), (relres, [i.relres for i in oinfo])
        one.close()
    owned = len(ds.owned)
    dist.barrier()
    print(f"RANK{rank} OK owned={owned} worst={worst:.3e} transport={ds.transport}", flush=True)
    ds.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
