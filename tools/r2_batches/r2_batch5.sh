set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b5.txt
for v in main nofrag; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = main ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 600 python tools/sweep.py --models 70b --ms 512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b5_$v.jsonl > gpurun_out/r2/sweep_b5_$v.log 2>&1
done
python tools/trace.py --m 2048 --n 28672 --k 8192 --units 40 > gpurun_out/r2/trace_2048_b5.txt 2>&1
python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/case_b5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/r2/prof_prefill_b5 -f python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/ncu_b5.log 2>&1
echo done
