#!/usr/bin/env python3
"""Key metrics per launch of an `ncu --set full` report (read here, on the build container: `ncu -i rep --page raw --csv`).
usage: python tools/ncu_summary.py gpurun_out/r2_prof_frontend.ncu-rep > profiles/r2_ncu_full_frontend.csv"""
import csv
import io
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        r = list(r)
        r[ki] = r[ki].split("(")[0].replace("b200::", "").replace("void ", "")[:60]
        w.writerow([r[i] for i in idx])
