"""Static evidence from the shipped library (no GPU needed): the cubins it holds, the instruction mix of the scoring
kernel and of its hot loop, and the mnemonics that show the 1-D bulk TMA copies + mbarriers.
usage: python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "structure_from_motion_b200", "libsfm_b200.so")
KERNEL = "_ZN3sfm7k_scoreILi2ELi16ELi1EEEvNS_9ScoreArgsE"

print("$ cuobjdump -lelf", os.path.relpath(SO, ROOT))
print(subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout.strip())
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(sass) if "Function :" in l] + [len(sass)]
funcs = {sass[a].split(":")[1].strip(): sass[a:b] for a, b in zip(starts, starts[1:])}
print("\n%d kernels in the cubin; mnemonic totals over all of them:" % len(funcs))
tot = collections.Counter()
for body in funcs.values():
    for l in body:
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            tot[m.group(1).split(".")[0]] += 1
for k in ("DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDS", "LDG", "ATOMS", "ATOMG", "REDUX", "MATCH", "SHF", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA", "DMMA"):
    print("  %-8s %6d" % (k, tot.get(k, 0)))
body = funcs[KERNEL]
print("\n== %s (k_score<HPT 2, G 16, SCREEN>) ==" % KERNEL)
ins = []
for l in body:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?) ;", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
# hot loop = the backward branch with the largest body that contains only DFMA/LDS/SHF-class work
best = None
for pc, text in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", text)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < pc:
            span = [t for p, t in ins if tgt <= p <= pc]
            nd = sum(t.split()[0].startswith("DFMA") or " DFMA" in t[:12] for t in span)
            if nd >= 300 and (best is None or len(span) < len(best[3])):  # the innermost loop that holds the screen
                best = (nd, tgt, pc, span)
nd, lo, hi, span = best
mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in span)
print("hot loop 0x%04x..0x%04x: %d instructions per 16 correspondences x 64 hypotheses" % (lo, hi, len(span)))
print("  " + ", ".join("%s %d" % kv for kv in mix.most_common()))
print("  LDS.128: %d   UBLKCP in kernel: %d   SYNCS in kernel: %d" % (
    sum("LDS.128" in t for t in span), sum("UBLKCP" in t for _, t in ins), sum("SYNCS" in t for _, t in ins)))
print("\nfirst 40 instructions of the hot loop:")
for t in span[:40]:
    print("   ", t)
print("\nbulk-copy / mbarrier instructions of the kernel:")
for pc, t in ins:
    if "UBLKCP" in t or "SYNCS" in t:
        print("    /*%04x*/ %s" % (pc, t))
