#!/usr/bin/env python
"""Hot spots of one kernel in an .ncu-rep: per-SASS-instruction executed counts and stall samples,
grouped in address order; prints the top instructions and cumulative per region."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# may contain several kernels: split on "Kernel Name" rows
blocks, cur = [], None
for l in lines:
    if l.startswith('"Kernel Name"'):
        cur = [l]; blocks.append(cur)
    elif cur is not None:
        cur.append(l)
for b in blocks:
    name = next(csv.reader([b[0]]))[1][:80]
    rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]
    tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"== {name}: {len(data)} SASS instr, {tot_inst} warp-instr executed, {tot_samp} samples")
    ranked = sorted(range(len(data)), key=lambda k: -int(data[k][ix["# Samples"]]))[:top]
    for k in sorted(ranked):
        r = data[k]
        print(f"{k:6d} {r[ix['Source']].strip()[:70]:70s} samp={r[ix['# Samples']]:>6s} exec={r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>5s}")
    # region histogram: 20 buckets
    nb = 25; sz = (len(data) + nb - 1) // nb
    for b0 in range(0, len(data), sz):
        seg = data[b0:b0 + sz]
        print(f"  [{b0:5d}-{b0 + len(seg):5d}) inst={sum(int(r[ix['Instructions Executed']]) for r in seg):>11d} samp={sum(int(r[ix['# Samples']]) for r in seg):>7d}")
