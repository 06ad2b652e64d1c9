"""Selected metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`) -> one CSV row per kernel.
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep "label" >> profiles/ncu_x.csv"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
if len(sys.argv) < 4:
    w.writerow(["what", "Kernel Name", "Grid Size", "Block Size"] + METRICS)
    w.writerow(["(units)", "", "", ""] + [units[hdr.index(m)] for m in METRICS])
for r in rows[2:]:
    w.writerow([sys.argv[2], r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]] +
               [r[hdr.index(m)] for m in METRICS])
