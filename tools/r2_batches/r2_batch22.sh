set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2/tests_gpu_b22.txt
timeout 600 python tools/mlp_bench.py > gpurun_out/r2/mlp_b22.txt 2>&1
timeout 600 python bench.py --no-extra > gpurun_out/r2/bench_b22.json 2> gpurun_out/r2/bench_b22.err
echo done
