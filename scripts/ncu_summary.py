"""ncu report -> the metrics quoted in profiles/README.md as a small CSV:  python scripts/ncu_summary.py in.ncu-rep out.csv"""
import csv, subprocess, sys
KEEP = ("gpu__time_duration.sum", "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg",
        "sm__warps_active.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed", "launch__", "smsp__average_warps_issue_stalled",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__sass_thread_inst_executed_op_d", "sm__throughput.avg", "gpu__dram_throughput")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["launch", "kernel", "metric", "unit", "value"])
    ki = hdr.index("Kernel Name")
    for li, r in enumerate(rows[2:]):
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) for k in KEEP) and not h.endswith(("min", "max")) and ".min." not in h and ".max." not in h and ".sum.p" not in h:
                w.writerow([li, r[ki][:60], h, u, v])
