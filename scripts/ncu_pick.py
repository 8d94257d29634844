"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
pats = sys.argv[1:] or ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "sm__warps_active.avg.pct_of_peak_sustained_active", "bank_conflicts", "issue_stalled",
                        "sm__throughput.avg.pct", "lts__t_sectors_op", "registers_per_thread", "l1tex__throughput.avg.pct",
                        "lts__throughput.avg.pct", "dram__throughput.avg.pct"]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if any(p in h for p in pats) and "_not_issued" not in h:
            print(f"{h:90s} {r[i]:>16s} {rows[1][i]}")
    print()
