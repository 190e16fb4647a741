#!/usr/bin/env bash
# Decode-size sweep of the W6Ax GEMM harness over the linear-layer shapes of the model families the reference sweeps
# (role of /root/reference/engine/test_flexq_kernel.sh:7-39): one result file per (model, M, N, K) under flexq_results/,
# named like the reference's so its plotting scripts find them.  Run from compat/ after `make`.
#   BS="1 2 4 8 16" ./test_flexq_kernel.sh        # override the batch sizes
set -euo pipefail
cd "$(dirname "$0")"
out=flexq_results
mkdir -p "$out"
# model : hidden : ffn   (qkv = 3*hidden x hidden, o = hidden x hidden, up = ffn x hidden on A6; down = hidden x ffn on A8)
models="llama_7b:4096:11008 llama_30b:6656:17920 llama_2_13b:5120:13824 llama_2_70b:8192:28672 opt_30b:7168:28672"
for M in ${BS:-1 2 4 8}; do
  for spec in $models; do
    IFS=: read -r name h f <<< "$spec"
    for layer in "$((3 * h)) $h 6" "$h $h 6" "$f $h 6" "$h $f 8"; do
      read -r N K xb <<< "$layer"
      ./bin/test_bgemm_kernel "$M" "$N" "$K" "$xb" 6 > "$out/${name}_${M}x${N}x${K}_w6a${xb}.txt"
    done
  done
done
