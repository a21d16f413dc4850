#!/usr/bin/env bash
# usage (GPU box): tools/bench_dg_variants.sh out.log variant...   ("default" = the in-tree library)
out="$1"; shift
: > "$out"
for v in "$@"; do
    lib="par2d_b200/csrc/build/variants/libpar2d_gpu_$v.so"
    [ "$v" = default ] && lib="par2d_b200/libpar2d_gpu.so"
    line=$(PAR2D_GPU_LIB="$PWD/$lib" timeout 300 python tools/bench_dg.py --steps 10 2>&1 | grep '^{' | tail -1)
    echo "$v $(echo "$line" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms/step %.3f  G/s %.3f  frac %.4f status %d" % (d["ms_per_step"], d["cell_updates_per_s"]/1e9, d["frac_of_hbm_roofline"], d["status"]))' 2>&1)" >> "$out"
done
cat "$out"
