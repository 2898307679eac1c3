"""profiles/r2_sass_summary.txt: per-kernel counts of the SASS instructions that show which hardware paths libqpb200.so
uses (run after a build; needs cuobjdump and c++filt, no GPU):   python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quadraticprogramsolver_b200", "libqpb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
COLS = ["UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "SHFL", "ATOMG", "RED", "CCTL", "BAR"]
FIRST = ["UBLKCP", "DMMA", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK"]
funcs, first, cur = [], {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = {"name": m.group(1), "n": dict.fromkeys(COLS, 0)}
        funcs.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not (m and cur):
        continue
    op = m.group(1)
    for c in COLS:
        if op == c or op.startswith(c + "."):
            cur["n"][c] += 1
    for f in FIRST:
        if op.startswith(f) and f not in first:
            first[f] = line.rstrip()
names = subprocess.run(["c++filt"], input="\n".join(f["name"] for f in funcs), capture_output=True, text=True).stdout.splitlines()
print("cuobjdump -sass quadraticprogramsolver_b200/libqpb200.so (sm_100a) -- per kernel: counts of the instructions that show\n"
      "which hardware paths the code uses.  UBLKCP = 1-D TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, DMMA = FP64\n"
      "tensor pipe (mma.sync.m8n8k4.f64), DFMA/DMUL/DADD = FP64 pipe, LDS/STS = shared memory, SHFL = warp shuffles,\n"
      "ATOMG/RED = global atomics, CCTL = cache control (L1 invalidate in the grid barrier).  Written by scripts/sass_summary.py.\n")
print(f"{'kernel':<84}" + "".join(f"{c:>7}" for c in COLS))
for f, nm in zip(funcs, names):
    nm = re.sub(r"\(.*$", "", nm)
    print(f"{nm:<84}" + "".join(f"{f['n'][c]:>7}" for c in COLS))
for f in FIRST:
    if f in first:
        print(f"\nfirst occurrence of {f}:\n{first[f]}")
