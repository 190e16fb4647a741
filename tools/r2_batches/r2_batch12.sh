set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b12.txt
python tools/llama_diag.py > gpurun_out/r2/llama_diag_b12.txt 2>&1
for v in main noalign b9; do
  lib=$PWD/flexq_b200/libflexq_b200.so
  [ $v = b9 ] && lib=$PWD/tools/ubench/ab/lib_b9.so
  pct=80; [ $v = noalign ] && pct=1000
  FLEXQ_ALIGN_PCT=$pct FLEXQ_B200_LIB=$lib timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,64,128,256,512 --no-cublas --out gpurun_out/r2/sweep_b12_$v.jsonl > gpurun_out/r2/sweep_b12_$v.log 2>&1
done
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 > gpurun_out/r2/trace_16_4096_b12.txt 2>&1
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 --cta -1 > gpurun_out/r2/trace_16_8192_b12.txt 2>&1
FLEXQ_ALIGN_PCT=1000 python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 > gpurun_out/r2/trace_16_4096_b12_noalign.txt 2>&1
python tools/trace.py --m 256 --n 4096 --k 4096 --units 40 --cta -1 > gpurun_out/r2/trace_256_4096_b12.txt 2>&1
echo done
