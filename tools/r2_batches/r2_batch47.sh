set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b47.txt
for rep in 1 2; do
for v in main shfl; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = shfl ] && lib=$PWD/tools/ubench/ab/lib_shfl.so
  FLEXQ_B200_LIB=$lib timeout 600 python tools/sweep.py --models 70b,7b --ms 512,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b47_${v}_$rep.jsonl > gpurun_out/r2/sweep_b47_${v}_$rep.log 2>&1
  FLEXQ_B200_LIB=$lib timeout 120 python tools/run_case.py --m 2048 --n 8192 --k 1024 --iters 6 >> gpurun_out/r2/k1024_b47_$v.txt 2>&1
done
done
echo done
