#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on (needs -lineinfo).

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
SORT = 0 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out = {}
h = None
for r in rows:
    if r and r[0] == "Line No":
        h = r
        continue
    if h is None or len(r) < 8 or not r[0].isdigit():
        continue                     # SASS rows carry an empty "Line No"; CUDA rows aggregate their instructions
    ie = h.index("Instructions Executed"); ws = h.index("Warp Stall Sampling (All Samples)")
    try:
        n = int(r[ie]); w = int(r[ws])
    except ValueError:
        continue
    key = (r[0], r[1])
    a, b = out.get(key, (0, 0))
    out[key] = (a + n, b + w)
ti = sum(v[0] for v in out.values()) or 1
tw = sum(v[1] for v in out.values()) or 1
print("instructions %d, stall samples %d" % (ti, tw))
for (ln, src), (n, w) in sorted(out.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print("%5.1f%% inst %5.1f%% stall  L%-5s %s" % (100.0 * n / ti, 100.0 * w / tw, ln, src.strip()[:150]))
