"""Development tool: print the interesting metrics of an `ncu --page raw --csv` export."""
import csv
import re
import sys

PAT = (r"gpu__time_duration.sum|smsp__average_warps_issue_stalled.*per_issue_active|smsp__warps_active.avg.per_cycle_active|"
       r"smsp__inst_executed.sum$|sm__cycles_elapsed.avg$|smsp__issue_active.avg.pct|launch__registers_per_thread$|"
       r"launch__occupancy_limit|dram__bytes_(read|write).sum$|sm__inst_executed_pipe_[a-z]+.avg.pct_of_peak_sustained_active|"
       r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
       r"launch__waves_per_multiprocessor|sm__warps_active.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
       r"smsp__thread_inst_executed_per_inst_executed.ratio")
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d["Kernel Name"][:70])
    for k in hdr:
        if re.search(PAT, k) and d[k] not in ("", "0"):
            print(f"  {k:90s} {d[k]}")
