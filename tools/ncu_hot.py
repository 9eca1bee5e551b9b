#!/usr/bin/env python3
"""Top stall lines of one kernel from an ncu source page CSV (ncu -i rep --page source --csv -k regex:NAME).
usage: tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"] and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top}")
