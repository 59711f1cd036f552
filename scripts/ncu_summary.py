#!/usr/bin/env python
"""Print the handful of ncu metrics that matter for this project from a .ncu-rep (reads via `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
    "sm__pipe_tensor_subpipe", "tensor",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_read.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct",
    "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second",
    "lts__cycles_elapsed.avg.per_second", "smem", "shared",
]

rep = sys.argv[1]
flt = sys.argv[2:] if len(sys.argv) > 2 else KEYS
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
for r in data:
    print("==", r[name_col][:80], "id", r[0])
for i, h in enumerate(hdr):
    if any(k in h for k in flt):
        print(f"{h:90s} {units[i]:>12s}  " + "  ".join(r[i] for r in data))
