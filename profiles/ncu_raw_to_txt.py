#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the `metric = value unit` summary kept under profiles/ (a selection of
the sections the roofline discussion uses: dram, launch, occupancy, issue, pipes, stalls, shared-memory conflicts).
usage: ncu_raw_to_txt.py <raw.csv> "<header line>" > profiles/<label>_ncu_full.txt"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
names, units, vals = rows[0], rows[1], rows[2]
keep = re.compile(r"^(dram__|gpu__time|gpu__dram|gpc__cycles_elapsed.max|launch__|sm__cycles_active.avg|sm__inst_executed|sm__throughput|"
                  r"sm__warps_active|smsp__average_warps_issue_stalled|smsp__inst_executed|smsp__issue_active|smsp__pcsamp|smsp__thread_inst|"
                  r"smsp__warps|l1tex__data_bank_conflicts|l1tex__data_pipe_lsu_wavefronts_mem_shared|lts__t_sector_hit|lts__t_bytes|"
                  r"sm__sass_thread_inst|smsp__sass_average|sm__maximum_warps|sm__ctas|.*TriageCompute.*(dram__throughput|alu|xu|uniform|cycles_active))")
print(sys.argv[2] if len(sys.argv) > 2 else "")
for n, u, v in sorted(zip(names, units, vals)):
    if keep.search(n):
        print("%s = %s %s" % (n, v, u))
