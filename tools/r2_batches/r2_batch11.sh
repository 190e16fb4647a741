set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b11.txt
for v in main b9 main2 b9b; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  case $v in b9*) lib=$PWD/tools/ubench/ab/lib_b9.so;; esac
  FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b --ms 16,128,256,512,1024,2048 --no-cublas --out gpurun_out/r2/sweep_b11_$v.jsonl > gpurun_out/r2/sweep_b11_$v.log 2>&1
done
timeout 600 python bench.py > gpurun_out/r2/bench_b11.json 2> gpurun_out/r2/bench_b11.err
for m in 16 300 2048; do
  timeout 600 compute-sanitizer --tool memcheck python tools/run_case.py --m $m --n 8192 --k 8192 --iters 2 > gpurun_out/r2/sanitizer_memcheck_m$m.txt 2>&1
done
for m in 16 300; do
  timeout 900 compute-sanitizer --tool racecheck python tools/run_case.py --m $m --n 4096 --k 4096 --iters 1 > gpurun_out/r2/sanitizer_racecheck_m$m.txt 2>&1
done
echo done
