set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b27.txt
for rep in 1 2; do
for v in auto hand; do
  h=-1; [ $v = hand ] && h=1
  FLEXQ_HANDOFF=$h timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 192,384,576,768 --no-cublas --out gpurun_out/r2/sweep_b27_${v}_$rep.jsonl > gpurun_out/r2/sweep_b27_${v}_$rep.log 2>&1
done
done
echo done
