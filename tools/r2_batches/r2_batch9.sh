set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b9.txt
for v in main nofrag; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = main ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b --ms 16,128,256,512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b9_$v.jsonl > gpurun_out/r2/sweep_b9_$v.log 2>&1
done
timeout 600 python bench.py > gpurun_out/r2/bench_b9.json 2> gpurun_out/r2/bench_b9.err
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 > gpurun_out/r2/trace_16_b9.txt 2>&1
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 > gpurun_out/r2/trace_16_4096_b9.txt 2>&1
echo done
