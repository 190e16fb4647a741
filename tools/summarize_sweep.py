#!/usr/bin/env python
"""Markdown summary of a tools/sweep.py JSONL file (for profiles/)."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print("| model | layer (NxK) | xb | M | GEMM us | TOPS | HBM frac | INT8 frac | fused us | cuBLAS f16 us | cuBLAS i8 us | vs f16 |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    f16 = r.get("cublas_f16_us")
    i8 = r.get("cublas_i8_us")
    print(f"| {r['model']} | {r['layer']} | {r['x_bits']} | {r['M']} | {r['gemm_us']:.1f} | {r['gemm_tops']:.1f} | {r['hbm_frac_gemm']:.3f} | "
          f"{r['i8_frac_gemm']:.3f} | {r['fused_us']:.1f} | {f16 and f'{f16:.1f}' or '-'} | {i8 and f'{i8:.1f}' or '-'} | "
          f"{f16 and f'{f16 / r['gemm_us']:.2f}x' or '-'} |")
