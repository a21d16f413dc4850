#!/usr/bin/env bash
# Same-box A/B of the large-mesh step (200 steps each): the round-1 tree (if built under build/r1tree) next to variants
# of the current library.  usage (GPU box): tools/ab_large.sh out.log name[:ENV=val,...] ...   ("default" = in-tree lib)
out="$1"; shift
: > "$out"
fmt='import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("%-44s value %6.2f G  step %.4f  kernel %.4f  frac %.4f  clk %s" % (sys.argv[1], d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["frac"], d["clocks"]["sm_mhz"]))'
r1() { [ -d par2d_b200/csrc/build/r1tree ] && (cd par2d_b200/csrc/build/r1tree && timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | grep '^{' | python -c "$fmt" "r1-tree" >> "$OLDPWD/$out"); }
:
for spec in "$@"; do
    name="${spec%%:*}"; envs=""
    if [[ "$spec" == *:* ]]; then envs="$(echo "${spec#*:}" | tr ',' ' ')"; fi
    lib="par2d_b200/csrc/build/variants/libpar2d_gpu_$name.so"
    if [ "$name" = default ]; then lib="par2d_b200/libpar2d_gpu.so"; fi
    env $envs PAR2D_GPU_LIB="$PWD/$lib" timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gate --preheat-ms 1200 --sustain-ms 100 2>&1 | grep '^{' | python -c "$fmt" "$spec" >> "$out"
done
:
cat "$out"
