set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b49.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_b49.txt 2>&1
timeout 600 python bench.py --no-extra > gpurun_out/r2/bench_b49.json 2> gpurun_out/r2/bench_b49.err
echo done
