#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: top stalled instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# the dump may contain several kernels; take sections starting at a header row
sections, cur = [], None
for r in rows:
    if r and r[0] == "Address":
        cur = {"hdr": r, "data": []}; sections.append(cur)
    elif r and r[0] == "Kernel Name":
        name = r[1]
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[idx['# Samples']]) for r in data)
    print("total samples", tot, "instructions", len(data), "warp-instr executed", sum(int(r[idx['Instructions Executed']]) for r in data))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for k, r in enumerate(data):
        r.append(k)
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:topn]:
        s = int(r[idx['# Samples']])
        main = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{s:6d} {100*s/tot:5.1f}% #{r[-1]:4d} {r[idx['Source']].strip()[:64]:64s} {main}")
    print({h[6:]: sum(int(r[idx[h]]) for r in data) for h in stalls})
