#!/usr/bin/env python
"""Extract per-launch DRAM traffic of the kernels of an `ncu --set full` capture into profiles/ncu_traffic.json (read by bench.py
for roofline.traffic).

  python tools/ncu_traffic.py <report.ncu-rep> <workload, e.g. C3> <batch>
"""
import csv, json, os, subprocess, sys
rep, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi, ti = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("b200::", "").replace("void ", "").split("<")[0]
    rd = float(r[ri].replace(",", "")) * scale[units[ri]]
    wr = float(r[wi].replace(",", "")) * scale[units[wi]]
    e = out.setdefault(name, {"workload": workload, "batch": batch, "launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0,
                              "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"})
    e["launches"] += 1; e["dram_bytes_read"] += rd; e["dram_bytes_write"] += wr
for e in out.values():          # per launch
    e["dram_bytes_read"] = round(e["dram_bytes_read"] / e["launches"]); e["dram_bytes_write"] = round(e["dram_bytes_write"] / e["launches"])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
