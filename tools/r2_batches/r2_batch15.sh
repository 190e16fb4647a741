set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b15.txt
for rep in 1 2; do
for v in atom atomsleep hand; do
  FLEXQ_B200_LIB=$PWD/tools/ubench/ab/lib_$v.so timeout 900 python tools/sweep.py --models 70b,7b --ms 16,256,512,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b15_${v}_$rep.jsonl > gpurun_out/r2/sweep_b15_${v}_$rep.log 2>&1
done
done
echo done
