"""Developer tool: print selected metrics from an `ncu --page raw --csv` dump.  usage: ncu_pick.py raw.csv [regex]"""
import csv
import re
import sys

DEFAULT = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__throughput.avg.pct|smsp__issue_active.avg.pct|"
           r"sm__warps_active.avg.pct|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|launch__registers_per_thread|"
           r"launch__grid_size|launch__occupancy_limit|smsp__inst_executed.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|"
           r"lts__t_bytes.sum$|l1tex__t_bytes.sum$|warp_issue_stalled.*per_warp_active.pct|smsp__warps_eligible.avg.per_cycle_active|"
           r"sm__inst_executed_pipe_(xu|fma|alu|lsu|fmaheavy|fp32).*sum$|gpu__dram_throughput.avg.pct|sm__pipe.*pct_of_peak_sustained_active")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else DEFAULT)
for i, h in enumerate(hdr):
    if pat.search(h):
        print(f"{h} [{units[i]}]: " + "  ".join(r[i] for r in rows[2:]))
