set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b8.txt
for v in main nofrag; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = main ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b --ms 16,128,256,512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b8_$v.jsonl > gpurun_out/r2/sweep_b8_$v.log 2>&1
done
timeout 600 python bench.py > gpurun_out/r2/bench_b8.json 2> gpurun_out/r2/bench_b8.err
timeout 300 python bench.py --config c1 > gpurun_out/r2/c1_b8.json 2> gpurun_out/r2/c1_b8.err
echo done
