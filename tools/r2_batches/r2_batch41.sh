set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b41.txt
timeout 300 python tools/sweep.py --models 70b --ms 512,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b41.jsonl > gpurun_out/r2/sweep_b41.log 2>&1
for s in "2048 8192 1024" "2048 3584 8192" "2048 8192 3584"; do set -- $s; timeout 120 python tools/run_case.py --m $1 --n $2 --k $3 --iters 6 >> gpurun_out/r2/tp_shapes_b41.txt 2>&1; done
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2/bench_b41.json 2> gpurun_out/r2/bench_b41.err
echo done
