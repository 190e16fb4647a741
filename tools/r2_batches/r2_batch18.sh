set -x
mkdir -p gpurun_out/r2
for t in 128 192; do
  FLEXQ_MTILE=$t timeout 900 python tools/sweep.py --models 70b,7b,l3-8b --ms 192,256,320,384,512,640,768,1024 --no-cublas --out gpurun_out/r2/sweep_b18_mtile$t.jsonl > gpurun_out/r2/sweep_b18_mtile$t.log 2>&1
done
python tools/trace.py --m 2048 --n 8192 --k 8192 --units 40 --cta 0 > gpurun_out/r2/trace_2048_8192_b18.txt 2>&1
python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/case_b18.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/r2/prof_prefill_b18 -f python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/ncu_b18.log 2>&1
echo done
