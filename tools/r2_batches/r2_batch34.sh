set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster or decomposition or repeatab" 2>&1 | tail -8 > gpurun_out/r2/tests_gpu_b34a.txt
grep -q "passed" gpurun_out/r2/tests_gpu_b34a.txt && ! grep -q "failed" gpurun_out/r2/tests_gpu_b34a.txt || { echo "cluster tests failed, stopping"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b34.txt
for rep in 1 2; do
for v in main nocluster; do
  rt=1; [ $v = nocluster ] && rt=0
  FLEXQ_CLUSTER_RUNTIME=$rt timeout 300 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16 --no-cublas --out gpurun_out/r2/sweep_b34_${v}_$rep.jsonl > gpurun_out/r2/sweep_b34_${v}_$rep.log 2>&1
done
done
timeout 120 python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 --noflush > gpurun_out/r2/trace_16_4096_hot_b34.txt 2>&1
echo done
