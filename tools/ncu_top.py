#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        hdr = rows[k + 1]
        j = k + 2
        data = []
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            if len(rows[j]) == len(hdr):
                data.append(rows[j])
            j += 1
        isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ismp]) for r in data)
        print(f"== {name}: {len(data)} SASS lines, {tot} samples")
        agg = {}
        for r in data:
            for c in stall:
                agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
        print("   stall totals:", sorted(((v, n) for n, v in agg.items() if v), reverse=True)[:8])
        for i in sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:top]:
            r = data[i]
            st = sorted([(int(r[c]), hdr[c]) for c in stall], reverse=True)[:2]
            print(f"   {i:5d} smp={r[ismp]:>6} exec={r[iex]:>8}  {r[isrc].strip()[:80]:<80} {st}")
        k = j
        break           # first kernel instance is enough
    k += 1
