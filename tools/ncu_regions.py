"""Per-region stall breakdown from an .ncu-rep source page: instructions grouped by execution count
(= code region), with the stall-reason mix of each region.
usage: python tools/ncu_regions.py rep.ncu-rep [--top N]"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 8
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = [r for r in rows[start + 1:] if r and r[0].startswith("0x")]
reasons = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
n_instr, samples, fp64 = Counter(), Counter(), Counter()
mix = defaultdict(Counter)
for r in body:
    ex = int(r[5])
    n_instr[ex] += 1
    samples[ex] += int(r[2])
    op = r[1].split()[1] if r[1].startswith("@") else r[1].split()[0]
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64")):
        fp64[ex] += 1
    for i, name in reasons:
        mix[ex][name] += int(r[i] or 0)
tot_i = sum(k * v for k, v in n_instr.items())
tot_s = sum(samples.values())
print(f"{tot_i} warp-instructions, {tot_s} stall samples")
for ex, _ in sorted(n_instr.items(), key=lambda t: -samples[t[0]])[:top]:
    m = mix[ex]
    ms = sum(m.values()) or 1
    tops = ", ".join(f"{k} {100 * v / ms:.0f}%" for k, v in m.most_common(6))
    print(f"executed {ex:>10d} x {n_instr[ex]:4d} instrs ({fp64[ex]:3d} fp64): {100 * ex * n_instr[ex] / tot_i:5.1f}% of instructions, "
          f"{100 * samples[ex] / tot_s:5.1f}% of samples | {tops}")
