set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b19.txt
timeout 1500 python tools/sweep.py --models 70b,7b,l3-8b --ms 1,16,64,128,192,256,384,512,768,1024,2048,4096 --out gpurun_out/r2/sweep_b19.jsonl > gpurun_out/r2/sweep_b19.log 2>&1
echo done
