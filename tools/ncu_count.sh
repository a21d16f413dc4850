#!/usr/bin/env bash
# usage (GPU box): tools/ncu_count.sh out.txt variant...   executed-instruction counts of one march_kernel launch per library variant
out="$1"; shift
: > "$out"
for v in "$@"; do
    lib="par2d_b200/csrc/build/variants/libpar2d_gpu_$v.so"
    [ "$v" = default ] && lib="par2d_b200/libpar2d_gpu.so"
    PAR2D_GPU_LIB="$PWD/$lib" ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
        --clock-control none -k regex:march_kernel -s 4 -c 1 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null \
        | grep march_kernel | python -c "
import csv, sys
for r in csv.reader(sys.stdin):
    print('$v', r[-3], r[-1])
" >> "$out"
done
cat "$out"
