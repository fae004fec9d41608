"""Summaries of the ncu captures for profiles/ (run here, no GPU needed):
    python scratch/summarize_ncu.py list gpurun_out/launches.csv  > profiles/..._launch_shares.txt
    python scratch/summarize_ncu.py full gpurun_out/prof.ncu-rep  > profiles/..._ncu_full_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpc__cycles_elapsed.avg.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__inst_executed_pipe_xu.avg.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio")


def rows_of(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


mode, path = sys.argv[1], sys.argv[2]
if mode == "list":
    rows = rows_of(open(path).read())
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) > mv and r[hdr.index("Metric Name")] == "gpu__time_duration.sum":
            unit = r[hdr.index("Metric Unit")]
            v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
            tot[r[kn]] += v
            cnt[r[kn]] += 1
    s = sum(tot.values())
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k[:70]:70s} n={cnt[k]:3d} total={tot[k]:9.3f} ms share={100 * tot[k] / s:5.1f}%")
else:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = rows_of(txt)
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("-----")
        print(f"{'Kernel Name':95s} {r[kn]}")
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEEP):
                print(f"{h:95s} {r[i]} {units[i]}")
