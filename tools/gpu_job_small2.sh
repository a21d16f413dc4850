#!/usr/bin/env bash
# usage (GPU box): bash tools/gpu_job_small2.sh TAG   - fused/persistent tests, the per-step profile, the shipped cases
tag="${1:-x}"
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_driver.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -n 6 gpurun_out/${tag}_pytest.log
timeout 300 python tools/diag_persist.py Case_ffstep8 Case_ffstep1 Case_spikefine24 Case_rktnoz9 > gpurun_out/${tag}_diag.log 2>&1
grep -v "longest\|bookkeeping CTA" gpurun_out/${tag}_diag.log
timeout 300 python tools/bench_cases.py --no-ref --steps 6000 > gpurun_out/${tag}_cases.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/${tag}_cases.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["case"], d["gas"], "%.2f us/step" % d["gpu_us_per_step"], "relres", d["relres_last"])
PY
