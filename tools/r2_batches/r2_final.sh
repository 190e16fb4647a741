set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_final.txt
timeout 900 python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err
timeout 1500 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,64,128,192,256,384,512,768,1024,2048,4096 --out gpurun_out/r2/sweep_final.jsonl > gpurun_out/r2/sweep_final.log 2>&1
timeout 300 python bench.py --config c1 > gpurun_out/r2/c1_final.json 2> gpurun_out/r2/c1_final.err
echo done
