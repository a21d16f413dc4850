"""Summarise an .ncu-rep (read here on the CPU box): key roofline metrics + stall reasons + per-pipe instruction mix."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_', 'launch__occupancy_limit', 'smsp__issue_active.avg.pct', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled', 'lts__t_sector_hit_rate', 'l1tex__t_sector_hit_rate', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__inst_executed_op_shfl', 'smsp__thread_inst_executed_per_inst_executed', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__sass_inst_executed_op_', 'sm__sass_inst_executed_op_', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
for i, h in enumerate(hdr):
    if any(w in h for w in want) and not any(x in h for x in ('.max.', '.min.', '.sum.pct', '.sum.per_second')) :
        vals = " | ".join(r[i] for r in data)
        if 'pipe_' in h and 'pct' not in h: continue
        print(f"{h:100s} {units[i]:16s} {vals}")
