"""Compact per-launch table from an .ncu-rep: python tools/ncu_raw_summary.py <rep> [out.csv]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
idx = [h.index(w) for w in want if w in h]
out = [[h[i] for i in idx], [u[i] for i in idx]] + [[r[i][:90] for i in idx] for r in rows[2:]]
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w", newline="")).writerows(out)
for r in out[2:]:
    print(" | ".join(f"{a.split('__')[-1][:28]}={b}" for a, b in zip(out[0], r)))
