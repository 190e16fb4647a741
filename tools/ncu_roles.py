#!/usr/bin/env python
"""Per-role instruction / stall budget of the warp-specialised GEMM from an ncu capture taken with --import-source on.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_roles.py sass.csv [group_tiles]

The kernel's role branches each start with a setmaxnreg (USETMAXREG in SASS), so the SASS between two of them is one
role; a region is named after the instructions only that role uses.  "polling" = instructions of mbarrier retry loops
(the second SYNCS...TRYWAIT of a wait and the loop around it: same execution count as that retry).
group_tiles (default: executions of the first UTCIMMA) turns totals into warp-instructions per 128x192x128 group-tile."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

cuts = [0] + [i for i, r in enumerate(data) if "USETMAXREG" in r[isrc]] + [len(data)]
regions = []
for a, b in zip(cuts[:-1], cuts[1:]):
    seg = data[a:b]
    text = " ".join(r[isrc] for r in seg)
    if a == 0:
        name = "prologue"
    elif "UTCIMMA" in text:
        name = "MMA issuers"
    elif "UBLKCP" in text or ("UTMALDG" in text and "STTM" not in text and "LDTM" not in text and text.count("UTMALDG") <= 2):
        name = "W producer"
    elif "UTMALDG" in text:
        name = "X/scale producer"
    elif "STTM" in text:
        name = "expanders"
    elif "LDTM" in text:
        name = "epilogue"
    else:
        name = "other"
    regions.append((name, seg))

mma = [int(r[iex]) for r in data if "UTCIMMA" in r[isrc]]
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else (max(mma) if mma else 1)
tot_inst = sum(int(r[iex]) for r in data)
tot_smp = sum(int(r[ismp]) for r in data)
print(f"{len(data)} SASS instructions, {tot_inst} warp-instructions executed, {tot_smp} stall samples, {tiles} group-tiles")
print(f"{'role':18s} {'warp-inst':>12s} {'share':>6s} {'per tile':>9s} {'polling':>11s} {'poll/tile':>9s} {'samples':>8s}  top stall reasons")
agg = {}
for name, seg in regions:
    inst = sum(int(r[iex]) for r in seg)
    smp = sum(int(r[ismp]) for r in seg)
    # retry loops: a TRYWAIT directly preceded (within 6 instructions) by another TRYWAIT on the same barrier operand
    poll = 0
    for i, r in enumerate(seg):
        if "TRYWAIT" not in r[isrc]:
            continue
        op = r[isrc].split("TRYWAIT")[1]
        prev = [q for q in seg[max(0, i - 6):i] if "TRYWAIT" in q[isrc] and q[isrc].split("TRYWAIT")[1] == op]
        if not prev:
            continue
        n = int(r[iex])
        # the loop body: neighbours executed as often as the retry itself (+- 1 %)
        body = [q for q in seg[max(0, i - 4):i + 8] if n and abs(int(q[iex]) - n) <= 0.01 * n]
        poll += n * max(1, len(body))
    st = {}
    for r in seg:
        for c in stall_cols:
            st[hdr[c]] = st.get(hdr[c], 0) + int(r[c])
    top = sorted(((v, k) for k, v in st.items() if v), reverse=True)[:4]
    a = agg.setdefault(name, [0, 0, 0, {}])
    a[0] += inst; a[1] += poll; a[2] += smp
    for v, k in top:
        a[3][k] = a[3].get(k, 0) + v
for name, (inst, poll, smp, st) in agg.items():
    top = ", ".join(f"{k[6:]} {100 * v / max(smp, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{name:18s} {inst:12d} {100 * inst / tot_inst:5.1f}% {inst / tiles:9.1f} {poll:11d} {poll / tiles:9.1f} {smp:8d}  {top}")
print(f"{'total':18s} {tot_inst:12d} {100.0:5.1f}% {tot_inst / tiles:9.1f} {sum(a[1] for a in agg.values()):11d} "
      f"{sum(a[1] for a in agg.values()) / tiles:9.1f} {tot_smp:8d}")
