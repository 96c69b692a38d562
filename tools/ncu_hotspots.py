#!/usr/bin/env python
"""Top stall locations of an `ncu --page source --csv` dump (SASS view).

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [N]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
body = rows[2:]
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", tot, " instructions", len(body))
agg = {k: sum(int(r[col[k]] or 0) for r in body) for k in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
ex = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print("warp instructions executed", ex)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = {k[6:]: int(r[col[k]] or 0) for k in stall_cols if int(r[col[k]] or 0)}
    print(f"{i:5d} {r[col['# Samples']]:>5} ex={r[col['Instructions Executed']]:>7} {r[col['Source']].strip()[:70]:70s} {st}")
