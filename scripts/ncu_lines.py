#!/usr/bin/env python
"""Per-CUDA-source-line profile of one kernel in an .ncu-rep (needs -lineinfo and --import-source on):
warp instructions executed and stall samples per line, top N by instructions.
usage: ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys
path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
fname, rows, hdr = None, [], None
seen_kernel = 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr and r[0].isdigit():
        try:
            rows.append((fname, int(r[0]), r[1].strip(), int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]])))
        except (ValueError, IndexError):
            pass
# several launches of the kernel repeat the listing: keep the first occurrence of each (file, line)
agg = {}
for f, ln, src, s, i in rows:
    agg.setdefault((f, ln), (src, s, i))
tot_i = sum(v[2] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot_i}  samples {tot_s}")
for (f, ln), (src, s, i) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f:18s} {ln:5d} inst={i:>10d} ({100.0 * i / max(tot_i, 1):5.1f}%) samp={s:>6d} ({100.0 * s / max(tot_s, 1):5.1f}%)  {src[:110]}")
