set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b10.txt
timeout 900 python tools/sweep.py --models 70b,7b --ms 1,16,64,128,256,512,2048 --no-cublas --out gpurun_out/r2/sweep_b10_main.jsonl > gpurun_out/r2/sweep_b10_main.log 2>&1
python tools/trace.py --m 16 --n 8192 --k 8192 --units 12 > gpurun_out/r2/trace_16_b10.txt 2>&1
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta 101 > gpurun_out/r2/trace_16_4096_b10.txt 2>&1
echo done
