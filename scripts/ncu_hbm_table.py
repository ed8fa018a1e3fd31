#!/usr/bin/env python
"""Per-kernel achieved DRAM bandwidth from `ncu --set full` reports: (dram__bytes_read.sum + dram__bytes_write.sum) /
gpu__time_duration.sum against MEASURED_PEAKS.json's hbm_gbs.  usage: ncu_hbm_table.py TAG=report.ncu-rep [TAG=report ...]
(prints markdown rows; profiles/r02_v12_hbm_per_kernel.md was made from the r02x captures with PCISPH= and SPH=)."""
import csv, io, json, os, subprocess, sys
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TU = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
print("| run | kernel | us | DRAM read MB | DRAM write MB | achieved GB/s | % of HBM peak | ncu gpu__dram_throughput % |\n|---|---|---|---|---|---|---|---|")
for arg in sys.argv[1:]:
    tag, path = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt))); h, u = rows[0], rows[1]; ix = {k: i for i, k in enumerate(h)}
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        val = lambda k: float(r[ix[k]].replace(",", ""))
        dur = val("gpu__time_duration.sum") * TU[u[ix["gpu__time_duration.sum"]]]
        rd = val("dram__bytes_read.sum") * UNIT[u[ix["dram__bytes_read.sum"]]]
        wr = val("dram__bytes_write.sum") * UNIT[u[ix["dram__bytes_write.sum"]]]
        gbs = (rd + wr) / dur / 1e9
        print(f"| {tag} | `{name}` | {dur * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} | {val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
