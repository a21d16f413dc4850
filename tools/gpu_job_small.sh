#!/usr/bin/env bash
# GPU job: full GPU test suite, then the per-step profile and the timing of the shipped cases (persistent launch).
# usage (GPU box): bash tools/gpu_job_small.sh TAG [notest]
tag="${1:-x}"
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -n 3 gpurun_out/${tag}_pytest.log
fi
timeout 300 python tools/diag_persist.py Case_ffstep8 Case_ffstep1 Case_spikefine24 Case_rktnoz9 > gpurun_out/${tag}_diag.log 2>&1
cat gpurun_out/${tag}_diag.log
timeout 300 python tools/bench_cases.py --no-ref --steps 6000 > gpurun_out/${tag}_cases.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/${tag}_cases.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["case"], d["gas"], "%.2f us/step" % d["gpu_us_per_step"], "relres", d["relres_last"])
PY
