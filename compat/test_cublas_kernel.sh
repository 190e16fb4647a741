#!/usr/bin/env bash
# The same shape sweep through the cuBLAS FP16 comparator (role of /root/reference/engine/test_cublas_kernel.sh);
# results under cublas_results/ with the file names test_flexq_kernel.sh uses, so the two directories pair up.
set -euo pipefail
cd "$(dirname "$0")"
out=cublas_results
mkdir -p "$out"
models="llama_7b:4096:11008 llama_30b:6656:17920 llama_2_13b:5120:13824 llama_2_70b:8192:28672 opt_30b:7168:28672"
for M in ${BS:-1 2 4 8}; do
  for spec in $models; do
    IFS=: read -r name h f <<< "$spec"
    for layer in "$((3 * h)) $h 6" "$h $h 6" "$f $h 6" "$h $f 8"; do
      read -r N K xb <<< "$layer"
      ./bin/test_cublas_kernel "$M" "$N" "$K" > "$out/${name}_${M}x${N}x${K}_w6a${xb}.txt"
    done
  done
done
