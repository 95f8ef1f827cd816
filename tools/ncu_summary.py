"""Summarise an ncu --page raw --csv dump: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [regex]`."""
import csv, re, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__throughput.avg.pct|"
                 r"pipe_(fp64|tensor|fma|alu|xu|lsu|fmaheavy).*pct_of_peak_sustained_active|sm__warps_active.avg.pct|"
                 r"registers_per_thread|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|issue_stalled.*_pct|"
                 r"smsp__average_warps_issue_stalled.*per_issue_active|shared_mem_per_block|occupancy_limit|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
                 r"lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_.*sum$|sm__cycles_elapsed.max|smsp__warps_eligible.avg.per_cycle_active")
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
for r in rows[2 if units else 1:]:
    print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print("  %-100s %s %s" % (h, r[i], units[i] if units else ""))
