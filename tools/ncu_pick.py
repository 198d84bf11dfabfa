#!/usr/bin/env python
"""Pick the judged columns out of `ncu -i X.ncu-rep --page raw --csv` (developer tool): python tools/ncu_pick.py raw.csv > summary.csv"""
import csv
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_src_fp64.sum.per_second",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "smsp__cycles_active.avg",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([w for w, _ in idx])
out.writerow([units[i] for _, i in idx])
for r in rows[2:]:
    out.writerow([r[i] for _, i in idx])
