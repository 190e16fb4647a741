# round 2, batch 2: full parity run, pipeline trace, A/B of epilogue variants, ncu capture of the bias-MMA kernel
set -x
mkdir -p gpurun_out/r2
./tools/ubench/denorm_ffma2 > gpurun_out/r2/denorm_ffma2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b2.txt
python tools/trace.py --m 2048 --n 28672 --k 8192 --units 48 > gpurun_out/r2/trace_2048_b2.txt 2>&1
for v in main wg2 wg4 epiwait; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = main ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 600 python tools/sweep.py --models 70b --ms 1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b2_$v.jsonl > gpurun_out/r2/sweep_b2_$v.log 2>&1
done
python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/case_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/r2/prof_prefill_b2 -f python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/ncu_b2.log 2>&1
echo done
