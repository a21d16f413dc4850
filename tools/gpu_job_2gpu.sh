#!/usr/bin/env bash
# usage (2-GPU box): bash tools/gpu_job_2gpu.sh TAG   - the whole GPU suite (two-GPU tests included), then the driver's 2-GPU bench line
tag="${1:-x}"
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q -rs > gpurun_out/${tag}_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_2gpu.log
tail -n 8 gpurun_out/${tag}_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_2gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.2f G  step %.4f ms  frac %.3f  e2e %.2f G" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e9))
        for g in d["parity_gate"]: print("  gate", g["case"], g["mode"], g["ok"], g["state_err"], g["bitwise"])
        for r in d["shipped_cases"]: print("  ", r.get("case"), r.get("us_per_step"), r.get("skipped"))
PY
