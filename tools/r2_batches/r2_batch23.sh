set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2/tests_gpu_b23.txt
timeout 600 python tools/producer_bench.py > gpurun_out/r2/producers_b23.jsonl 2> gpurun_out/r2/producers_b23.err
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2/bench_short_b23.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_b23.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2/ncu_launches_b23.log 2>&1
python tools/run_case.py --m 16 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/case_decode_b23.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/r2/prof_decode_b23 -f python tools/run_case.py --m 16 --n 28672 --k 8192 --iters 3 > gpurun_out/r2/ncu_decode_b23.log 2>&1
echo done
