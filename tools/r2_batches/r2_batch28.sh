set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b28.txt
for c in 1 0; do FLEXQ_TMAP_CACHE=$c python tools/host_overhead.py >> gpurun_out/r2/host_overhead_b28.txt 2>&1; done
timeout 600 python tools/producer_bench.py > gpurun_out/r2/producers_b28.jsonl 2> gpurun_out/r2/producers_b28.err
echo done
