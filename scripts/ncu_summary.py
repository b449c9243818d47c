"""Extract the judged metrics from .ncu-rep files (run where ncu is installed, no GPU needed) into one CSV.
usage: python scripts/ncu_summary.py out.csv a.ncu-rep [b.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["report", "kernel", "launch"] + WANT)
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for k, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.writerow([rep.split("/")[-1], d.get("Kernel Name", "")[:60], k] + [f"{d.get(m, '')} {u.get(m, '')}".strip() for m in WANT])
print("wrote", sys.argv[1])
