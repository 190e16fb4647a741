set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b13.txt
for v in main b9; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = b9 ] && lib=$PWD/tools/ubench/ab/lib_b9.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,64,128,256,512,1024,2048 --no-cublas --out gpurun_out/r2/sweep_b13_$v.jsonl > gpurun_out/r2/sweep_b13_$v.log 2>&1
done
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 > gpurun_out/r2/trace_16_4096_b13.txt 2>&1
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 --cta -1 > gpurun_out/r2/trace_16_8192_b13.txt 2>&1
python tools/trace.py --m 512 --n 4096 --k 4096 --units 30 --cta -1 > gpurun_out/r2/trace_512_4096_b13.txt 2>&1
echo done
