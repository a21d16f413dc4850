"""Throughput of the shipped reference cases (BASELINE.json configs[0..3]) through the C ABI on one GPU, next to the
unmodified reference binary (one shim MPI rank per block) on the box's host cores.  These meshes hold 43-55 k cells:
a step is launch-latency bound, not HBM bound - reported for completeness, the roofline claim is bench.py's.
usage: python tools/bench_cases.py [--steps N] [--out file.json]"""
import argparse, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
import torch  # noqa: F401  (device bring-up)
from par2d_b200 import Case, Solver, _capi, default_config

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20000)
ap.add_argument("--batch", type=int, default=2000)
ap.add_argument("--ref-iters", type=int, default=1500)
ap.add_argument("--out", default=None)
ap.add_argument("--no-ref", action="store_true", help="skip the reference CPU runs")
args = ap.parse_args()
gpu = _capi.load_gpu()
import run_ref
therm = os.path.join(REPO, "tests", "golden", "cases", "therm_air.dat")
rows = []
for name, igam in (("Case_ffstep1", 1.4), ("Case_ffstep8", 1.4), ("Case_rktnoz9", 1.4), ("Case_rktnoz9", -1.4), ("Case_spikefine24", 1.4)):
    path = os.path.join(REPO, "tests", "golden", "cases", name)
    case = Case.load(path)
    cfg = default_config(gpu); cfg.device = 0; cfg.igam = igam
    s = Solver(case, cfg, lib=gpu)
    s.step(200)
    t0 = time.perf_counter(); done = 0
    while done < args.steps:
        info = s.step(args.batch); done += len(info)
    dt = time.perf_counter() - t0
    row = {"case": name, "gas": "CPG" if igam > 0 else "TPG (NASA-7)", "blocks": case.nblocks, "cells": case.ncells, "steps": done,
           "gpu_us_per_step": 1e6 * dt / done, "gpu_cell_updates_per_s": case.ncells * done / dt, "relres_last": info[-1].relres}
    if run_ref.available() and igam > 0 and not args.no_ref:
        shapes = [case.shape(b) for b in range(case.nblocks)]
        tw = run_ref.run(path, therm, shapes, 100)["wall"]
        tf = run_ref.run(path, therm, shapes, 100 + args.ref_iters)["wall"]
        row.update({"ref_cpu_cell_updates_per_s": case.ncells * args.ref_iters / max(tf - tw, 1e-9), "ref_ranks": case.nblocks,
                    "host_cores": os.cpu_count(), "ref_note": "unmodified reference binary, g++ -O2, one shim rank per block, t(100+N)-t(100)"})
        row["speedup"] = row["gpu_cell_updates_per_s"] / row["ref_cpu_cell_updates_per_s"]
    rows.append(row)
    print(json.dumps(row), flush=True)
    s.close()
if args.out:
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)
