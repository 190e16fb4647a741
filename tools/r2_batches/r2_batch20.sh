set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b20.txt
for rep in 1 2; do
for v in cluster nocluster; do
  rt=1; [ $v = nocluster ] && rt=0
  FLEXQ_CLUSTER_RUNTIME=$rt timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,32,64 --no-cublas --out gpurun_out/r2/sweep_b20_${v}_$rep.jsonl > gpurun_out/r2/sweep_b20_${v}_$rep.log 2>&1
done
done
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 > gpurun_out/r2/trace_16_4096_b20.txt 2>&1
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 --cta -1 > gpurun_out/r2/trace_16_8192_b20.txt 2>&1
echo done
