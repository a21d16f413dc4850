"""Throughput of the second-order branch (cfg.accur = 1, DG(P1)) on the benchmark mesh (8 blocks of 2048^2 cells,
one GPU).  Algorithmic bytes per cell-update: mean 32 + slopes 64 read, mean 32 + slopes 64 written, nodes 16
= 208 B.  usage: python tools/bench_dg.py [--block 2048] [--steps 20] [--out file.json]"""
import argparse, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from par2d_b200 import Case, Solver, _capi, default_config

ap = argparse.ArgumentParser()
ap.add_argument("--block", type=int, default=2048)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--out", default=None)
args = ap.parse_args()
gpu = _capi.load_gpu()
case = Case.synthetic_ramp(4 * args.block, 2 * args.block, 4, 2)
cfg = default_config(gpu); cfg.device = 0; cfg.accur = 1; cfg.mode = args.mode
s = Solver(case, cfg, lib=gpu)
s.step(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
info = s.step(args.steps)
dt = time.perf_counter() - t0
peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0
row = {"workload": f"synthetic ramp, 8 blocks of {args.block}^2 cells, accur=1 (DG(P1)), mode={args.mode}", "cells": case.ncells, "steps": len(info),
       "ms_per_step": 1e3 * dt / len(info), "cell_updates_per_s": case.ncells * len(info) / dt,
       "algorithmic_bytes_per_cell": 208, "hbm_gbs_algorithmic": 208 * case.ncells * len(info) / dt / 1e9, "hbm_peak_gbs": peak,
       "frac_of_hbm_roofline": 208 * case.ncells * len(info) / dt / 1e9 / peak, "relres_last": info[-1].relres, "status": info[-1].status}
print(json.dumps(row))
if args.out:
    json.dump(row, open(args.out, "w"), indent=1)
