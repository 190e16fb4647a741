set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b4.txt
python tools/trace.py --m 2048 --n 28672 --k 8192 --units 40 > gpurun_out/r2/trace_2048_b4.txt 2>&1
python tools/trace.py --m 2048 --n 28672 --k 8192 --units 40 --cta 145 > gpurun_out/r2/trace_2048_b4_spare.txt 2>&1
for v in main pre nosx skipbias noepi nold; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = main ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 600 python tools/sweep.py --models 70b --ms 2048,4096 --no-cublas --out gpurun_out/r2/sweep_b4_$v.jsonl > gpurun_out/r2/sweep_b4_$v.log 2>&1
done
echo done
