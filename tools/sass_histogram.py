#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): the evidence for which Blackwell instructions
the hand-written kernels use (FFMA2 = packed fp32 FMA, DMMA.8x8x4 = FP64 tensor cores, UBLKCP = TMA bulk copy, SYNCS =
mbarrier, LDGSTS = cp.async) and which they do not (UTMALDG / UTC*MMA / LDTM: no tcgen05 kind exists for FP64 or exact FP32).

    python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "go-audio-resampler_b200" / "_build" / "libgar_b200.so")
KEY = ["FFMA2", "FFMA", "DFMA", "DMMA", "HMMA", "UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "SYNCS", "LDGSTS",
       "LDS", "STS", "LDG", "STG", "BAR", "ATOMS", "MATCH", "SHFL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"gar::\(anonymous namespace\)::|\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(gar::\w+(, .*)?\)$|\(.*\)$", "", cur)
        per[cur] = Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_x]+)*)", line)
    if m and cur:
        op = m.group(1)
        per[cur][op.split(".")[0]] += 1
        if op.startswith("DMMA") or op.startswith("FFMA2"):
            per[cur][op.split(" ")[0]] += 0
total = Counter()
for c in per.values():
    total.update(c)
print(f"# {Path(lib).name}: {len(per)} kernels, {sum(total.values())} SASS instructions (sm_100a)")
print("# totals of the instructions of interest: " + ", ".join(f"{k}={total.get(k, 0)}" for k in KEY))
print("# absent (by design, see DESIGN.md): " + ", ".join(k for k in ("UTMALDG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "HMMA") if not total.get(k)))
print()
print(f"{'kernel':100s} {'instr':>7s}  " + " ".join(f"{k:>6s}" for k in KEY[:5] + KEY[5:6] + KEY[11:13]))
for name in sorted(per, key=lambda n: -sum(per[n].values())):
    c = per[name]
    cols = KEY[:5] + KEY[5:6] + KEY[11:13]
    print(f"{name[:100]:100s} {sum(c.values()):7d}  " + " ".join(f"{c.get(k, 0):6d}" for k in cols))
