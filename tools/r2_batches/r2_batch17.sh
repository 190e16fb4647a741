set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b17.txt
for rep in 1 2; do
for v in main notstore; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = notstore ] && lib=$PWD/tools/ubench/ab/lib_notstore.so
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b --ms 128,256,512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b17_${v}_$rep.jsonl > gpurun_out/r2/sweep_b17_${v}_$rep.log 2>&1
done
done
timeout 600 python bench.py --no-extra > gpurun_out/r2/bench_b17.json 2> gpurun_out/r2/bench_b17.err
echo done
