#!/usr/bin/env python
"""python scripts/ncu_pick.py rep.ncu-rep [regex] -- print the metrics matching regex (default: a useful set)"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|smsp__inst_executed.sum$|issue_active.avg.pct|warps_active.avg.pct|"
                 r"l1tex__data_pipe_(lsu|tex)_wavefronts(_mem_shared)?.sum$|l1tex__data_pipe.*pct_of_peak_sustained_elapsed|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__inst_executed_pipe_(lsu|fma|alu|tex|fmaheavy|fmalite|uniform).sum$|"
                 r"stalled_.*per_issue_active|lts__t_bytes.sum$|lts__t_sector_hit_rate|sm__cycles_active.avg$|sm__cycles_elapsed.max|registers_per_thread|"
                 r"smsp__inst_executed_op_shared_ld.sum$|l1tex__throughput.avg.pct|lts__throughput.avg.pct|gpu__dram_throughput.avg.pct|sm__throughput.avg.pct|"
                 r"launch__occupancy_limit|achieved_occupancy|sm__inst_executed_pipe_.*pct")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print("  %-90s %s %s" % (h, r[i], units[i]))
