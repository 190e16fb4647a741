set -x
mkdir -p gpurun_out/r2
python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/case_b46.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/r2/prof_prefill_b46 -f python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/ncu_b46.log 2>&1
echo done
