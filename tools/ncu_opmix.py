#!/usr/bin/env python
"""Executed warp-instructions per opcode from an `ncu --page source --csv` dump."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {k: i for i, k in enumerate(hdr)}
cnt = collections.Counter()
for r in rows[2:]:
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", src)
    if not m: continue
    op = m.group(2)
    mods = m.group(3)
    if op == "IMAD" and (".MOV" in mods or ".IADD" in mods or ".SHL" in mods or ".HI" in mods or ".WIDE" in mods):
        op += "." + [x for x in ("MOV", "IADD", "SHL", "HI", "WIDE") if "." + x in mods][0]
    if op == "LEA" and ".HI" in mods: op = "LEA.HI"
    cnt[op] += int(r[col["Instructions Executed"]] or 0)
tot = sum(cnt.values())
FMA = {"IMAD", "IMAD.MOV", "IMAD.IADD", "IMAD.SHL", "IMAD.HI", "IMAD.WIDE", "IDP", "FFMA", "HFMA2", "FMUL", "FADD"}
fma = sum(v for k, v in cnt.items() if k in FMA)
mem = sum(v for k, v in cnt.items() if k in {"LDS", "STS", "LDG", "STG", "ST", "LD", "LDL", "STL", "LDC", "LDCU", "SYNCS", "UTMALDG"})
print(f"total {tot}  fma-pipe {fma} ({100*fma/tot:.1f}%)  mem/sync {mem} ({100*mem/tot:.1f}%)  other(alu etc) {tot-fma-mem} ({100*(tot-fma-mem)/tot:.1f}%)")
for k, v in cnt.most_common(30):
    print(f"{k:12s} {v:10d} {100*v/tot:5.1f}%")
