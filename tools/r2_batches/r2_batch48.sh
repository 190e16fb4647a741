set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b48.txt
timeout 600 python bench.py --no-extra > gpurun_out/r2/bench_b48.json 2> gpurun_out/r2/bench_b48.err
timeout 300 python tools/mlp_bench.py > gpurun_out/r2/mlp_b48.txt 2>&1
echo done
