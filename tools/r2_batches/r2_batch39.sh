set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster" 2>&1 | tail -8 > gpurun_out/r2/tests_gpu_b39a.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b39.txt
echo done
