#!/usr/bin/env bash
# usage (on the GPU box): tools/bench_variants.sh out.log name[:ENV=val,...] ...   "default" = the in-tree library
out="$1"; shift
: > "$out"
for spec in "$@"; do
    name="${spec%%:*}"; envs=""
    if [[ "$spec" == *:* ]]; then envs="$(echo "${spec#*:}" | tr ',' ' ')"; fi
    lib="par2d_b200/csrc/build/variants/libpar2d_gpu_$name.so"
    if [ "$name" = default ]; then lib="par2d_b200/libpar2d_gpu.so"; fi
    line=$(env $envs PAR2D_GPU_LIB="$PWD/$lib" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gate --preheat-ms 400 --sustain-ms 400 2>&1 | grep '^{' | tail -1)
    python - "$spec" "$line" >> "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    r = d["roofline"]
    print(f"{sys.argv[1]:40s} value {d['value']/1e9:7.2f} G/s  step {d['ms_per_step']:.4f} ms  kernel {r['kernel_ms']:.4f} ms  frac {r['frac']:.4f}  clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"{sys.argv[1]:40s} FAILED {e} {sys.argv[2][:200]}")
PY
done
cat "$out"
