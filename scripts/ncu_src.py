#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top stall sites per kernel.  usage: ncu_src.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None and "Source" in r:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for s in sections[:1] if len(sys.argv) <= 3 else sections:
    hdr, data = s["hdr"], s["data"]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data)
    print("==", s["name"][:90]); print("total samples", tot, "instructions", len(data))
    agg = {}
    for r in data:
        for c in stall_cols:
            if r[c] not in ("", "0"):
                agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c])
    print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:topn]
    for i in sorted(top):
        r = data[i]
        st = {hdr[c][6:]: int(r[c]) for c in stall_cols if r[c] not in ("", "0")}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(i, r[isamp], r[iex], r[ia].strip()[:80], st)
