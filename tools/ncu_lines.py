#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture, read here without a GPU:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_cs.csv ; tools/ncu_lines.py X_cs.csv [N]
Prints the N source lines with the most executed warp instructions, their share, and stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "Instructions Executed" in r)
ci = {n: i for i, n in enumerate(hdr)}
iline, isrc, iinst, isamp = 0, 1, hdr.index("Instructions Executed"), hdr.index("# Samples")
lines = []
for r in rows:
    if r[iline].strip().isdigit():
        try:
            lines.append((int(r[iline]), r[isrc].strip(), int(r[iinst] or 0), int(r[isamp] or 0)))
        except ValueError:
            pass
tot_i = sum(l[2] for l in lines) or 1
tot_s = sum(l[3] for l in lines) or 1
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for ln, src, ins, smp in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{ln:5d} {100 * ins / tot_i:5.1f}% inst {100 * smp / tot_s:5.1f}% samp  {src[:110]}")
