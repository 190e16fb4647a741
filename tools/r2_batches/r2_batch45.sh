set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b45.txt
timeout 900 python bench.py > gpurun_out/r2/bench_b45.json 2> gpurun_out/r2/bench_b45.err
echo done
