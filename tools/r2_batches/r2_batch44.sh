set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decomposition or baseline_shapes or repeatab" 2>&1 | tail -4 > gpurun_out/r2/tests_gpu_b44.txt
for rep in 1 2; do
for v in 1 0; do
  FLEXQ_SPARE_MODEL=$v timeout 600 python tools/sweep.py --models 70b,7b,l3-8b --ms 512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b44_sm${v}_$rep.jsonl > gpurun_out/r2/sweep_b44_sm${v}_$rep.log 2>&1
done
done
echo done
