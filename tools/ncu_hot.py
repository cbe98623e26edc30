#!/usr/bin/env python
"""Per-instruction stall samples of one kernel from an ncu report: prints the hot loop (instructions executed at
least --frac of the most-executed one) with cycles-per-iteration estimates.  usage: ncu_hot.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        tables.append(cur)
    elif r and r[0] != "Address" and cur is not None and len(r) == len(hdr):
        cur.append(r)
data = max(tables, key=lambda t: sum(int(r[isamp]) for r in t))
tot = sum(int(r[isamp]) for r in data)
mx = max(int(r[iex]) for r in data)
agg = collections.Counter()
for r in data:
    for c in stall:
        agg[hdr[c][6:]] += int(r[c])
print("samples", tot, "instructions", len(data), "max executed", mx)
print("stall totals:", ", ".join(f"{k}={v} ({100*v/tot:.0f}%)" for k, v in agg.most_common(9)))
hot = [r for r in data if int(r[iex]) >= frac * mx]
hs = sum(int(r[isamp]) for r in hot)
print(f"hot region: {len(hot)} instructions, {hs} samples ({100*hs/tot:.0f}%)")
for r in hot:
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall), reverse=True)[:2]
    print(f"{r[isrc].strip()[:64]:64s} ex={int(r[iex])//1000:>7}k smp={int(r[isamp]):>6} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
