#!/usr/bin/env python
"""bench.py -- cell-updates/s of the explicit FV residual+update path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation

Workload (config.workload): the synthetic 10-degree ramp mesh of BASELINE.json configs[4], weak
scaling = 8 blocks of 2048x2048 cells per GPU, so that N = 8 is the full 16k x 16k / 64-block mesh
and N = 1 is its per-GPU share (4x2 blocks, 33.5 M cells, 1.07 GB of state: larger than the 126 MB
L2, so no flush between steps is needed).  A "step" is one explicit iteration over all blocks:
ghost fill + dt, residual assembly (LDFSS fluxes), norms, per-cell solve and update.

One JSON line is printed by rank 0 (see README/DESIGN for the keys).  value = device-timed whole-job
throughput with state resident in HBM; e2e = the same job through the public C ABI starting from
pinned HOST buffers (state upload, K steps with the per-step dt/relres log read back, state
download), all inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

BLOCK = 2048                      # cells per block side (BASELINE.json configs[4]: 64 blocks of 2048^2)
LAYOUT = {1: (4, 2), 2: (4, 4), 4: (8, 4), 8: (8, 8)}   # blocks in x, y for N GPUs (8 per GPU)
ALGO_BYTES_PER_CELL = 80          # state read 32 + state write 32 + node coordinates 16
LAUNCHES_PER_STEP = 2             # ghost_kernel (carries the previous step's residual bookkeeping), march_kernel (+1 pack_kernel for N > 1); +1 finalize_kernel per batch


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def count(self):
        return len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference / CPU-baseline leg (the one place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------

def _pow2_floor(n):
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def cpu_reference_run(iters: int, warm: int, max_ranks: int = 32):
    """Times the reference's CPU implementation of the path on the host cores on a bounded sample
    of the benchmark workload: the same ramp generator scaled down to 128x128-cell blocks, one block
    per reference MPI rank (the reference is one rank per block, FV/main.cpp:125-128), as many ranks
    as the box has cores (power of two, <= 32).  With oracle/_ref present the UNMODIFIED reference
    binary runs (kind "reference"); its step loop is isolated by differencing a `warm`-iteration and a
    `warm+iters`-iteration run (mesh read, metrics and solution writes are identical in both).
    Otherwise the oracle port (kind "port") is stepped in-process with OpenMP over blocks."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import run_ref
    from par2d_b200 import Case, Solver, _capi, default_config
    ncores = os.cpu_count() or 1
    ranks = min(max_ranks, _pow2_floor(ncores))
    nbx = 1
    while nbx * nbx * 2 <= ranks:
        nbx *= 2
    nby = ranks // nbx
    side = 128
    case = Case.synthetic_ramp(side * nbx, side * nby, nbx, nby)
    cells = case.ncells
    sample = (f"synthetic ramp {side * nbx}x{side * nby} cells in {nbx}x{nby} blocks of {side}x{side} (benchmark generator, "
              f"blocks scaled 16x down per side), {iters} iterations after {warm} warm-up")
    if run_ref.available():
        tmp = tempfile.mkdtemp(prefix="par2d_bench_case_")
        case.write(tmp)
        shapes = [case.shape(b) for b in range(case.nblocks)]
        therm = os.path.join(REPO, "tests", "golden", "cases", "therm_air.dat")
        t_warm = run_ref.run(tmp, therm, shapes, warm)["wall"]
        t_full = run_ref.run(tmp, therm, shapes, warm + iters)["wall"]
        dt = max(t_full - t_warm, 1e-9)
        kind, cores = "reference", min(ranks, ncores)
        sample += f"; unmodified reference binary, SHIM_NP={ranks} (g++ -O2), step loop = t({warm + iters}) - t({warm})"
    else:
        os.environ.setdefault("OMP_NUM_THREADS", str(ranks))
        orc = _capi.load_oracle()
        s = Solver(case, default_config(orc), lib=orc)
        s.step(warm)
        t0 = time.perf_counter()
        s.step(iters)
        dt = time.perf_counter() - t0
        kind, cores = "port", min(ranks, ncores)
        sample += f"; oracle port (gcc -O2, OpenMP over blocks, {ranks} threads)"
    return {"value": cells * iters / dt, "unit": "cell-updates/s", "cores": cores, "kind": kind, "sample": sample,
            "ms_per_iteration": 1e3 * dt / iters, "sample_cells": cells, "host_cores": ncores}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inner = 25   # reference iterations per bench "step" so that K steps last seconds, not milliseconds
    res = cpu_reference_run(iters=args.steps * inner, warm=args.warmup * inner)
    nbx, nby = LAYOUT[args.gpus]
    line = {
        "impl": "reference", "metric": "cell-updates/sec (fp64 residual+update)", "value": res["value"],
        "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_iteration"] * inner, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic 10deg ramp, {nbx * nby} blocks of {BLOCK}x{BLOCK} cells (8 per GPU)",
                   "reference_step": f"{inner} reference iterations on the bounded sample", "l2": "n/a (CPU)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

# ------------------------------------------------------------------------------------------------
# parity gate: runs before anything is timed, on the same ranks / transport as the timed job
# ------------------------------------------------------------------------------------------------

GATE_STATE_TOL = 1e-11     # max|d var| / max|var| per variable after N steps (BASELINE.json north_star)
GATE_HIST_TOL = 2e-6       # the reference prints relres with %le (7 digits) and dt with %7.4e (5 digits)


def parity_gate(gpu, world, rank, local, transport):
    """Advances two shipped cases on the N ranks of this job and compares the result with fixtures generated by
    the UNMODIFIED reference binary (tests/golden/golden_*.npz, written by tests/golden/make_golden.py; only the
    files are read - nothing under oracle/ is imported).  Replaces the reference's own check, which is the eye
    (SURVEY 4).  The blocks are sharded exactly like the timed job's (Case.partition, halo exchange
    SpatialDiscretization.cpp:217-471 over the peer-mapped arenas, residual gather main.cpp:379-392 in global
    block order).
      * Case_spikefine24 (24 blocks, irregular neighbours, subsonic far field), PAR2D_MODE_FAST, 200 steps:
        every cell of every block within GATE_STATE_TOL, relres/dt history within the printed precision;
      * Case_ffstep8 (8 blocks), PAR2D_MODE_FAITHFUL, 200 steps: every cell BITWISE equal to the reference.
    Returns the list of gate records; raises SystemExit(3) on failure (after all ranks agreed on it)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from par2d_b200 import Case, Solver, default_config
    multi = world > 1
    records, failed = [], False
    for name, mode, bitwise in (("Case_spikefine24", 0, False), ("Case_ffstep8", 1, True)):
        gold = np.load(os.path.join(REPO, "tests", "golden", f"golden_{name}_axi_cpg.npz"))
        steps = int(gold["steps"])
        case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
        cfg = default_config(gpu)
        cfg.device, cfg.mode = local, mode
        if multi:
            from par2d_b200.dist import DistSolver
            s = DistSolver(case, cfg, lib=gpu, transport=transport)
            dt, relres = s.step(steps)                         # relres: all blocks, global block order
        else:
            s = Solver(case, cfg, lib=gpu)
            info = s.step(steps)
            dt, relres = np.array([i.dt for i in info]), np.array([i.relres for i in info])
        err, nbad = 0.0, 0
        for b in s.owned:
            got, ref = s.get_state(b), gold[f"b{b}"]
            nbad += int(np.count_nonzero(got != ref))
            vel = np.sqrt(ref[..., 1] ** 2 + ref[..., 2] ** 2).max()
            for k, scale in enumerate((np.abs(ref[..., 0]).max(), vel, vel, np.abs(ref[..., 3]).max())):
                err = max(err, float(np.abs(got[..., k] - ref[..., k]).max() / scale))
        used = s.transport if multi else None
        s.close()
        agg = torch.tensor([err, float(nbad)], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(agg[0:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(agg[1:2], op=dist.ReduceOp.SUM)
        err, nbad = float(agg[0].item()), int(agg[1].item())
        hist_err = float(max(np.abs(relres / gold["relres"] - 1.0).max(), np.abs(dt / gold["dt"] - 1.0).max() / 50.0))
        ok = (nbad == 0 if bitwise else err <= GATE_STATE_TOL) and hist_err <= GATE_HIST_TOL and len(relres) == steps
        failed |= not ok
        records.append({"case": name, "ranks": world, "steps": steps, "mode": "FAITHFUL" if mode else "FAST",
                        "state_err": err, "hist_err": hist_err, "bitwise": nbad == 0, "cells_differing": nbad,
                        "transport": used, "relres_last": float(relres[-1]), "ok": bool(ok),
                        "against": f"tests/golden/golden_{name}_axi_cpg.npz (unmodified reference binary, {steps} steps)"})
    if failed:
        if rank == 0:
            print(json.dumps({"parity_gate": records, "error": "parity gate failed; nothing was timed"}), flush=True)
        raise SystemExit(3)
    return records


def shipped_cases(gpu, world, rank, local, transport, steps=3000):
    """BASELINE.json configs[0..3] on the same N ranks: us per step and cell-updates/s of the four shipped case
    directories (43-55 k cells: a step is bound by latency, not by HBM - the small-mesh path: one persistent launch per
    batch on one GPU, one fused launch per step across GPUs).  A case with fewer blocks than ranks is skipped."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from par2d_b200 import Case, Solver, default_config
    multi = world > 1
    rows = []
    for name, igam in (("Case_ffstep1", 1.4), ("Case_ffstep8", 1.4), ("Case_rktnoz9", -1.4), ("Case_spikefine24", 1.4)):
        case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
        if case.nblocks < world:
            rows.append({"case": name, "blocks": case.nblocks, "skipped": f"{case.nblocks} block(s) cannot be spread over {world} ranks"})
            continue
        cfg = default_config(gpu)
        cfg.device, cfg.igam = local, igam
        if multi:
            from par2d_b200.dist import DistSolver
            s = DistSolver(case, cfg, lib=gpu, transport=transport)
            adv = s.advance
        else:
            s = Solver(case, cfg, lib=gpu)
            adv = s.step
        adv(200)
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        t0 = time.perf_counter()
        left = steps
        while left > 0:
            info = adv(min(left, 1000))
            left -= min(left, 1000)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
        ok = all(i.status == 0 for i in info)
        rows.append({"case": name, "gas": "CPG" if igam > 0 else "TPG (NASA-7)", "blocks": case.nblocks, "cells": case.ncells, "ranks": world,
                     "steps": steps, "us_per_step": 1e6 * sec / steps, "cell_updates_per_s": case.ncells * steps / sec, "status_ok": ok,
                     "timing": "host wall clock around batches of 1000 steps (each batch ends with a synchronise + history read-back), max over ranks"})
        s.close()
    return rows


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from par2d_b200 import Case, Solver, _capi, default_config
    from par2d_b200._capi import StepInfo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gpu = _capi.load_gpu()      # raises when the CUDA library is missing: there is no fallback path
    gate = None if args.no_gate else parity_gate(gpu, world, rank, local, args.transport)
    shipped = None if args.no_gate else shipped_cases(gpu, world, rank, local, args.transport)
    # weak scaling (default, the driver's contract): 8 blocks per GPU; --strong: the full 8 x 8 block mesh on any N
    nbx, nby = LAYOUT[8] if args.strong else LAYOUT[args.gpus]
    blk = args.block
    case = Case.synthetic_ramp(blk * nbx, blk * nby, nbx, nby, nranks=world, rank=rank)
    total_cells = case.ncells
    cfg = default_config(gpu)
    cfg.device = local
    if args.faithful:
        cfg.mode = 1     # PAR2D_MODE_FAITHFUL: the reference's operation order, bit-identical results (not the headline)

    def sync_all():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    if multi:
        from par2d_b200.dist import DistSolver
        solver = DistSolver(case, cfg, lib=gpu, transport=args.transport)
        ctx = solver.solver
    else:
        solver = ctx = Solver(case, cfg, lib=gpu)
    # the library's work and the timing events share one real torch stream
    stream = solver.stream if multi else torch.cuda.Stream()
    if not multi:
        gpu.check(gpu.set_stream(ctx.handle, stream.cuda_stream), "set_stream")
    torch.cuda.set_stream(stream)
    h = ctx.handle
    owned_cells = ctx.owned_cells

    def advance(n):
        if multi:
            return solver.advance(n)
        return ctx.step(n)

    # ---- device-resident throughput -------------------------------------------------------
    # order: warm-up -> calibration -> ~1.5 s pre-heat of the same step loop (the board settles at its power
    # cap) -> the timed K steps -> >= 1 s continuation (value_sustained).  The clock sampler runs from the
    # start of the pre-heat to the end of the continuation, so its window CONTAINS the timed region.
    def timed(n, chunk=500):
        """n steps in batches of <= chunk; returns (ms, max over ranks; records of the last batch)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        left, last = n, None
        while left > 0:
            m = min(left, chunk)
            last = advance(m)
            left -= m
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last

    advance(args.warmup)
    cal_ms, _ = timed(10)
    per_step = max(cal_ms / 10, 1e-3)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    timed(int(args.preheat_ms / per_step) + 1)
    gpu.check(gpu.kernel_timing(h, 1), "kernel_timing")
    elapsed_ms, info = timed(args.steps, chunk=args.steps)          # THE timed region: exactly K steps, one batch
    kms, kn, nlaunch = C.c_double(), C.c_int64(), C.c_int64()
    gpu.check(gpu.kernel_time(h, C.byref(kms), C.byref(kn)), "kernel_time")
    gpu.check(gpu.launch_count(h, C.byref(nlaunch)), "launch_count")
    relres_last = (solver.global_relres(args.steps)[-1] if multi else info[-1].relres)
    n_sus = int(args.sustain_ms / per_step) + 1
    sus_ms, sus_info = timed(n_sus)
    kms_s, kn_s = C.c_double(), C.c_int64()
    gpu.check(gpu.kernel_time(h, C.byref(kms_s), C.byref(kn_s)), "kernel_time")
    gpu.check(gpu.kernel_timing(h, 0), "kernel_timing")
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = (f"~{args.preheat_ms / 1e3:.1f} s pre-heat + the timed region + ~{args.sustain_ms / 1e3:.1f} s continuation "
                            "of the same step loop (one uninterrupted sampling interval that contains the timed region)")
    assert all(i.status == 0 for i in info) and all(i.status == 0 for i in sus_info), "device raised an error flag during the timed region"
    value = total_cells * args.steps / (elapsed_ms * 1e-3)
    value_sustained = total_cells * n_sus / (sus_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers ------------------------------
    e2e = None
    if 32 * owned_cells > 6e9:
        e2e = {"value": None, "unit": "cell-updates/s", "note": "skipped: more than 6 GB of state per GPU would have to be pinned on the host (strong-scaling run)"}
    else:
        host_bufs = {}
        for b in ctx.owned:
            nx, ny = ctx.shapes[b]
            buf = torch.empty((ny - 1, nx - 1, 4), dtype=torch.float64, pin_memory=True)
            buf[..., 0], buf[..., 1], buf[..., 2], buf[..., 3] = cfg.uFS[0], cfg.uFS[1], cfg.uFS[2], cfg.uFS[3]
            host_bufs[b] = buf
        dp = C.POINTER(C.c_double)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids = (C.c_int32 * len(host_bufs))(*host_bufs.keys())
        ptrs = (dp * len(host_bufs))(*[C.cast(buf.data_ptr(), dp) for buf in host_bufs.values()])
        gpu.check(gpu.set_states(h, len(host_bufs), ids, ptrs), "set_states")     # copy of block i+1 overlaps the transpose of block i
        einfo = advance(args.steps)                    # per-step dt/relres records are copied back
        gpu.check(gpu.get_states(h, len(host_bufs), ids, ptrs), "get_states")
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        state_bytes = 32 * total_cells
        log_bytes = (C.sizeof(StepInfo) + 16 * case.nblocks) * args.steps
        e2e = {"value": total_cells * args.steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": state_bytes / args.steps, "d2h_bytes_per_step": (state_bytes + log_bytes) / args.steps,
               "note": "one job = state upload from pinned host memory + K steps with per-step log read-back + state download; "
                       "bytes are job totals divided by K"}
        assert all(i.status == 0 for i in einfo)
    transport = solver.transport if multi else None
    solver.close()        # collective for the peer-mapped transport: arenas are unmapped before anybody frees

    if rank != 0:
        if multi:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------
    peaks, which = measured_peaks()
    launches = max(int(kn.value), 1)
    k_ms = kms.value / launches
    achieved = ALGO_BYTES_PER_CELL * owned_cells / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("step_kernel_dram_bytes_per_launch")
    fp64 = C.c_double()
    gpu.check(gpu.probe_fp64(local, C.byref(fp64)), "probe_fp64")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "kernel": "march_kernel (fused residual assembly + norms + per-cell solve + update)", "kernel_ms": k_ms, "kernel_share_of_step": kms.value / elapsed_ms,
                "frac_sustained": (ALGO_BYTES_PER_CELL * owned_cells / (kms_s.value / max(int(kn_s.value), 1) * 1e-3) / 1e9) / peaks["hbm_gbs"],
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which}, burst copy)",
                "algorithmic_bytes_per_cell": ALGO_BYTES_PER_CELL, "cells_per_launch": owned_cells,
                "fp64_fma_tflops_measured": fp64.value,
                "cell_updates_per_s_per_gpu": owned_cells / (k_ms * 1e-3)}

    cpu = None
    if not args.no_cpu_baseline and args.gpus == 1:      # the CPU baseline is reported at N = 1 only
        res = cpu_reference_run(iters=500, warm=25)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": "cell-updates/sec (fp64 residual+update)", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic 10deg ramp, {nbx * nby} blocks of {blk}x{blk} cells ({nbx * nby // args.gpus} per GPU), "
                               f"{total_cells} cells, IAXI=1, CPG, CFL=0.01 (shipped macros)",
                   "global_cells": total_cells, "parallelism": f"blocks sharded over {args.gpus} GPU(s), 1 process per GPU",
                   "halo_exchange": (f"{transport} (" + ("kernel stores into NVLink peer-mapped arenas + sequence flags, no host transport; edge strips + put on a side stream beside the interior strips"
                                     if transport == "peer" else "NCCL send/recv + all-reduce(MAX) per step") + ")") if multi else "none (single GPU)",
                   "l2": "state per GPU 1.07 GB read + 1.07 GB written per step >> 126 MB L2; no flush needed",
                   "arith_mode": "PAR2D_MODE_FAITHFUL" if args.faithful else "PAR2D_MODE_FAST"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        # counted by the library: every kernel rank 0 launched inside the timed region (large mesh: ghost + march per step,
        # across ranks ghost + edge strips + interior strips + put; + the put and the bookkeeping launch that open / close the batch)
        "gpu_launches": int(nlaunch.value),
        "relres_last": relres_last, "dt": info[-1].dt,
        "value_sustained": value_sustained, "sustained_steps": n_sus,
        "parity_gate": gate,
        "shipped_cases": shipped,
    }
    print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1, choices=sorted(LAYOUT))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, default=BLOCK, help="cells per block side (default: the benchmark's 2048)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gate", action="store_true", help="skip the parity gate (development runs only)")
    ap.add_argument("--preheat-ms", type=float, default=1500.0, help="untimed run of the step loop before the timed region")
    ap.add_argument("--sustain-ms", type=float, default=1000.0, help="continuation after the timed region (value_sustained)")
    ap.add_argument("--faithful", action="store_true", help="time the bit-faithful arithmetic mode instead of the default fast one")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the full 16k x 16k / 64-block mesh on N GPUs (default: weak, 8 blocks per GPU)")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "external"],
                    help="multi-GPU halo exchange: NVLink peer-mapped arenas (default) or NCCL send/recv per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
