"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + hottest SASS lines.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top N] > profiles/xxx.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for kr in rows[2:]:
    name = kr[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name)
    for h, u, v in zip(hdr, units, kr):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address") + 1
body = [r for r in rows[start:] if len(r) > 6 and r[0].startswith("0x")]
tot_i = sum(int(r[5]) for r in body)
tot_s = sum(int(r[2]) for r in body)
print(f"\n== SASS: {len(body)} instructions, {tot_i} warp-instructions executed, {tot_s} stall samples")
fp64 = sum(int(r[5]) for r in body if r[1].split()[0].lstrip("@!P0123456789 ").startswith(("DFMA", "DMUL", "DADD", "DSETP"))
           or any(t in r[1] for t in (" DFMA ", " DMUL ", " DADD ", " DSETP ")))
print(f"   FP64-pipe warp-instructions: {fp64} ({100.0 * fp64 / tot_i:.1f} %)")
print(f"\n== top {top} instructions by stall samples (idx, executed, samples, sass)")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][2]))[:top]:
    print(f"{i:5d} {int(r[5]):12d} {int(r[2]):8d}  {r[1].strip()[:90]}")
