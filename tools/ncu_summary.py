#!/usr/bin/env python
"""Summarise an ncu capture (gpurun_out/*.ncu-rep + launch-list csv) into profiles/<name>.md (committed evidence).

  python tools/ncu_summary.py <name> <report.ncu-rep> [launches.csv]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def main():
    name, rep = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# ncu summary `%s`" % name, "", "source: `%s` (`ncu --set full --clock-control none --import-source on`), per launch" % os.path.basename(rep), "",
           "| kernel | " + " | ".join(m[1] for m in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        cells = []
        for m, _ in METRICS:
            if m in hdr:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, units[i]))
            else:
                cells.append("-")
        out.append("| `%s` | " % r[ki].split("(")[0] + " | ".join(cells) + " |")
    if len(sys.argv) > 3:
        rows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 10]
        h = rows[0]
        k, v, g = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            agg.setdefault(r[k].split("(")[0], []).append(float(r[v].replace(",", "")))
        tot = sum(sum(x) for x in agg.values())
        out += ["", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: shares, not absolutes)", "",
                "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for kk, x in agg.items():
            out.append("| `%s` | %d | %.3f | %.1f%% |" % (kk, len(x), sum(x) / 1e6, 100 * sum(x) / tot))
    path = os.path.join(ROOT, "profiles", name + ".md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
