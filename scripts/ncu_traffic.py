#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of one sub-step: DRAM bytes (read + write) per launch,
summed per bench.py phase.  usage: ncu_traffic.py report.ncu-rep particles out.json"""
import csv, io, json, subprocess, sys
PHASE_OF = {"k_hash_count": "grid", "k_scan_cells": "grid", "k_fill_incremental": "grid", "k_cell_lists_density": "density", "k_lists_density_tp": "density", "k_collide_predict": "force_np+predict", "k_collide_integrate": "pressure_force+integrate",
            "k_force_np_predict": "force_np+predict", "k_pressure_force": "pressure_force+integrate", "k_pressure": "pressure"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
path, particles, outp = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot, per_kernel, seen = {}, {}, {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    b = sum(float(r[ix[k]].replace(",", "")) * UNIT[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per_kernel[name] = per_kernel.get(name, 0) + b
    seen[name] = seen.get(name, 0) + 1
for name in per_kernel:  # a capture may hold more than one launch of a kernel: average per launch
    per_kernel[name] /= seen[name]
    ph = PHASE_OF.get(name)
    if ph:
        tot[ph] = tot.get(ph, 0) + per_kernel[name]
json.dump({"source": path.split("/")[-1], "particles": particles, "dram_bytes_per_launch": tot, "dram_bytes_per_kernel": per_kernel,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of one sub-step, ncu --set full --clock-control none"}, open(outp, "w"), indent=1)
print(json.dumps(tot))
