set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b21.txt
timeout 900 python bench.py > gpurun_out/r2/bench_b21.json 2> gpurun_out/r2/bench_b21.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_ref_b21.json 2> gpurun_out/r2/bench_ref_b21.err
timeout 300 python bench.py --config c1 > gpurun_out/r2/c1_b21.json 2> gpurun_out/r2/c1_b21.err
echo done
