#!/usr/bin/env python
"""Text summaries of ncu outputs for profiles/:
  ncu_summary.py launches <launches.csv>      per-kernel totals / shares of a gpu__time_duration launch list
  ncu_summary.py report <file.ncu-rep>        key raw metrics of a --set full capture (runs `ncu -i`)"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]

if sys.argv[1] == "launches":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * (1e3 if r[iu] in ("ms", "msecond") else 1.0)   # -> us
        name = r[ik].split("(")[0][:70]
        c = tot.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += v
    total = sum(c[1] for c in tot.values())
    print(f"total {total:.1f} us over {sum(c[0] for c in tot.values())} launches (ncu serialised, cold cache: compare shares)")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.2f}%  {t:12.1f} us  {n:5d} launches  avg {t / n:10.1f} us  {name}")
else:
    out = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print("kernel:", d[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:80s} {d[i]:>16s} {units[i]}")
