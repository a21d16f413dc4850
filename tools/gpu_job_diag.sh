#!/usr/bin/env bash
# usage (GPU box): bash tools/gpu_job_diag.sh TAG "case case ..." name[:ENV=val,...] ...   ("default" = in-tree library)
tag="$1"; cases="$2"; shift 2
out="gpurun_out/${tag}_diag.log"; : > "$out"
for spec in "$@"; do
    name="${spec%%:*}"; envs=""
    if [[ "$spec" == *:* ]]; then envs="$(echo "${spec#*:}" | tr ',' ' ')"; fi
    lib="par2d_b200/csrc/build/variants/libpar2d_gpu_$name.so"
    if [ "$name" = default ]; then lib="par2d_b200/libpar2d_gpu.so"; fi
    echo "== $spec" >> "$out"
    env $envs PAR2D_GPU_LIB="$PWD/$lib" timeout 300 python tools/diag_persist.py $cases >> "$out" 2>&1
done
cat "$out"
