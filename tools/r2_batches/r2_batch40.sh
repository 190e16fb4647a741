set -x
mkdir -p gpurun_out/r2
for s in "2048 8192 1024" "2048 3584 8192" "2048 8192 3584" "2048 8192 2048" "2048 7168 8192" "2048 8192 7168" "1152 8192 1024" "896 8192 3584"; do
  set -- $s
  timeout 120 python tools/run_case.py --m $1 --n $2 --k $3 --iters 6 >> gpurun_out/r2/tp_shapes_b40.txt 2>&1
done
timeout 120 python tools/trace.py --m 2048 --n 8192 --k 1024 --units 48 --cta -1 > gpurun_out/r2/trace_2048_k1024_b40.txt 2>&1
echo done
