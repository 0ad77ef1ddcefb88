"""Instruction counts of a kernel split at barriers, from an ncu report's source page.
Usage: python tools/sass_regions.py report.ncu-rep [lo hi]  (lo/hi: print SASS rows in that range)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > ie]
if len(sys.argv) > 3:
    for n in range(int(sys.argv[2]), int(sys.argv[3])):
        print(n, body[n][ie], body[n][isamp], body[n][ia][:100])
    sys.exit(0)
total = sum(int(r[ie] or 0) for r in body)
print("total instructions", total, "rows", len(body))
cur, tot, samp, start = collections.Counter(), 0, 0, 0
for n, r in enumerate(body):
    s = r[ia].strip()
    e, sm = int(r[ie] or 0), int(r[isamp] or 0)
    parts = s.split()
    op = (parts[1] if s.startswith("@") else parts[0]).split(".")[0]
    cur[op] += e
    tot += e
    samp += sm
    if op == "BAR" or ("SYNCS" in s and "TRYWAIT" in s) or n == len(body) - 1:
        print(f"{start:5d}-{n:5d} instr {tot:9d} {100 * tot / total:5.1f}% samples {samp:5d}", cur.most_common(8))
        cur, tot, samp, start = collections.Counter(), 0, 0, n + 1
