#!/usr/bin/env bash
# compute-sanitizer over small runs of every kernel family (GPU box): memcheck + racecheck + synccheck.
# usage: tools/sanitize.sh out.log
out="${1:-gpurun_out/sanitizer.log}"
: > "$out"
for tool in memcheck racecheck synccheck; do
    for args in "Case_spikefine24 3 0" "Case_rktnoz9 3 1" "ramp:256:192:2:2 2 0"; do
        echo "=== $tool: diag_step $args" >> "$out"
        timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/diag_step.py $args 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|worst=|Invalid|hazard|error" | head -8 >> "$out"
    done
    echo "=== $tool: second-order branch" >> "$out"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/bench_dg.py --block 96 --steps 2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|status" | cut -c1-300 | head -8 >> "$out"
done
cat "$out"
