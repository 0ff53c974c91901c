#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` for one kernel: stall-sample totals by
reason, instruction mix, and the hottest SASS regions (contiguous address ranges).  Usage:
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_source_summary.py src.csv [kernel-index]
"""
import csv
import sys
from collections import Counter, defaultdict


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(open(path)))
    # split into kernels
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    starts.append(len(rows))
    a, b = starts[which], starts[which + 1]
    print(rows[a][1][:120])
    hdr = rows[a + 1]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[a + 2:b] if len(r) >= len(hdr) - 2]
    samp = col["# Samples"]
    ex = col["Instructions Executed"]
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = Counter()
    mix = Counter()
    mix_ex = Counter()
    total_samples = 0
    for r in body:
        s = int(r[samp] or 0)
        total_samples += s
        op = r[col["Source"]].split()
        op = [o for o in op if not o.startswith("@")]
        m = op[0].split(".")[0] if op else "?"
        mix[m] += s
        mix_ex[m] += int(r[ex] or 0)
        for n in stall_cols:
            tot[n] += int(r[col[n]] or 0)
    print("total samples", total_samples)
    print("stalls:", ", ".join(f"{k[6:]}={v} ({100*v/max(1,total_samples):.1f}%)" for k, v in tot.most_common(10)))
    print("samples by opcode:", ", ".join(f"{k}={v}" for k, v in mix.most_common(14)))
    print("executed by opcode:", ", ".join(f"{k}={v}" for k, v in mix_ex.most_common(14)))
    # hot regions: windows of 64 instructions
    W = 64
    reg = []
    for i in range(0, len(body), W):
        chunk = body[i:i + W]
        s = sum(int(r[samp] or 0) for r in chunk)
        e = sum(int(r[ex] or 0) for r in chunk)
        ops = Counter((r[col["Source"]].split() or ["?"])[0].split(".")[0] for r in chunk)
        st = Counter()
        for r in chunk:
            for n in stall_cols:
                st[n[6:]] += int(r[col[n]] or 0)
        reg.append((s, i, e, ops, st))
    print("regions (64 instr): idx samples% executed  top-ops")
    for s, i, e, ops, st in reg:
        if s * 100 / max(1, total_samples) >= 1.0:
            print(f"  {i:6d} {100*s/total_samples:5.1f}% {e:10d}  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(4)) +
                  "  | " + " ".join(f"{k}={100*v/max(1,s):.0f}%" for k, v in st.most_common(4)))


if __name__ == "__main__":
    main()
