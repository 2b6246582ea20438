"""Selected per-kernel metrics of an .ncu-rep as CSV (the file committed under profiles/).
Usage: python tools/ncu_summary.py report.ncu-rep > summary.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "derived__lts__lts2xbar_bytes.sum.per_second",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [hdr.index("Kernel Name")] + [hdr.index(w) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] for i in cols])
out.writerow([units[i] for i in cols])
for r in rows[2:]:
    out.writerow([r[i] for i in cols])
