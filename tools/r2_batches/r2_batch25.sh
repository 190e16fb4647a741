set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b25.txt
for rep in 1 2; do
for v in main exp4; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = exp4 ] && lib=$PWD/tools/ubench/ab/lib_exp4.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,32 --no-cublas --out gpurun_out/r2/sweep_b25_${v}_$rep.jsonl > gpurun_out/r2/sweep_b25_${v}_$rep.log 2>&1
done
done
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 --cta -1 > gpurun_out/r2/trace_16_8192_b25.txt 2>&1
echo done
