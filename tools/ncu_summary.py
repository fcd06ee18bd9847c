#!/usr/bin/env python
"""Print the handful of ncu metrics the roofline discussion uses from a .ncu-rep (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/permute_w4.ncu-rep > profiles/r01_permute_w4.txt"""
import csv
import re
import subprocess
import sys

PAT = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|"
                 r"launch__grid_size|launch__block_size|launch__cluster|launch__occupancy_limit_\w+|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
                 r"smsp__inst_executed.sum|sm__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"lts__throughput.avg.pct_of_peak_sustained_elapsed|smsp__issue_active.avg.pct_of_peak_sustained_active|"
                 r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|"
                 r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_\w+_per_issue_active.ratio|"
                 r"smsp__cycles_active.avg|sm__cycles_elapsed.max|lts__t_sectors_srcunit_tex_op_read.sum|lts__t_sector_hit_rate.pct|"
                 r"launch__shared_mem_per_block_\w+|smsp__inst_executed_op_\w+\.sum)$")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for n, r in enumerate(rows[2:]):
    print(f"--- launch {n}: {r[hdr.index('Kernel Name')][:110]}")
    for k, u, v in zip(hdr, units, r):
        if PAT.match(k):
            print(f"{k:90s} {v:>16s} {u}")
