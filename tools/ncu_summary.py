"""Summarise an ncu --page raw --csv dump:
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [--traffic-json out.json workload kernel_regex]
With --traffic-json it also writes {"workload", "kernel", "dram_bytes_per_launch"} for bench.py's roofline.traffic."""
import csv, json, re, sys
args = sys.argv[1:]
tj = None
if args and args[0] == "--traffic-json":
    tj = (args[1], args[2], re.compile(args[3])); args = args[4:]
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
pat = re.compile(args[0] if args else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__throughput.avg.pct|"
                 r"pipe_(fp64|tensor|fma|alu|xu|lsu|fmaheavy).*pct_of_peak_sustained_active|sm__warps_active.avg.pct|"
                 r"registers_per_thread|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|issue_stalled.*_pct|"
                 r"smsp__average_warps_issue_stalled.*per_issue_active|shared_mem_per_block|occupancy_limit|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
                 r"lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_.*sum$|sm__cycles_elapsed.max|smsp__warps_eligible.avg.per_cycle_active|"
                 r"lts__t_bytes.sum$|l1tex__data_pipe_lsu_wavefronts.sum$|smem|tmem")
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
for r in rows[2 if units else 1:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    vals = {}
    for i, h in enumerate(hdr):
        if pat.search(h):
            print("  %-100s %s %s" % (h, r[i], units[i] if units else ""))
        vals[h] = (r[i], units[i] if units else "")
    if tj and tj[2].search(name):
        def to_bytes(key):
            v, u = vals[key]
            f = float(v.replace(",", ""))
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tot = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        json.dump({"workload": tj[1], "kernel": name[:60], "dram_bytes_per_launch": tot,
                   "dram_read": to_bytes("dram__bytes_read.sum"), "dram_write": to_bytes("dram__bytes_write.sum")}, open(tj[0], "w"))
        tj = None
