# round 2, batch 1: subnormal-FFMA2 ubench, parity of the bias-MMA kernel, A/B of the prefill variants
set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/smi.txt
./tools/ubench/denorm_ffma2 > gpurun_out/r2/denorm_ffma2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b1.txt
for v in base bias rearm; do
  lib=$PWD/tools/ubench/ab/lib_$v.so
  [ $v = bias ] && lib=$PWD/flexq_b200/libflexq_b200.so
  FLEXQ_B200_LIB=$lib timeout 600 python tools/sweep.py --models 70b --ms 256,512,1024,2048,4096 --no-cublas --out gpurun_out/r2/sweep_b1_$v.jsonl > gpurun_out/r2/sweep_b1_$v.log 2>&1
done
echo done
