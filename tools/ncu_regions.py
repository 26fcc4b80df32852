#!/usr/bin/env python
"""SASS-level execution profile of one kernel from an ncu report: consecutive instructions with the same execution
count are merged into regions (count per warp, share of all executed instructions, active threads per instruction).

  python tools/ncu_regions.py <report.ncu-rep> <kernel regex> [min share %]
"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ie, te, ws = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[ie].isdigit():
        data.append((r[1].strip(), int(r[ie]), int(r[te]), int(r[ws])))
    elif r and r[0] == "Address":
        break            # next launch of the same kernel
tot = sum(d[1] for d in data); tw = sum(d[3] for d in data) or 1
nwarps = data[0][1]
print("instructions %d  warps %d  per warp %.0f  stall samples %d" % (tot, nwarps, tot / nwarps, tw))
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[start][1]) > 0.02 * max(data[start][1], 1):
        seg = data[start:i]
        a = sum(d[1] for d in seg)
        if 100.0 * a / tot >= minshare:
            print("%4d-%4d n=%3d  x%7.2f/warp  %5.1f%% inst %5.1f%% stall  thr %4.1f   %s" % (
                start, i - 1, len(seg), seg[0][1] / nwarps, 100.0 * a / tot, 100.0 * sum(d[3] for d in seg) / tw,
                sum(d[2] for d in seg) / max(1, a), seg[0][0][:60]))
        start = i
