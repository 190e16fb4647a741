set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b26.txt
for rep in 1 2; do
for v in main pre; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = pre ] && lib=$PWD/tools/ubench/ab/lib_pre.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b --ms 512,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b26_${v}_$rep.jsonl > gpurun_out/r2/sweep_b26_${v}_$rep.log 2>&1
done
done
echo done
