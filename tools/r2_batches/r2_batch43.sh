set -x
mkdir -p gpurun_out/r2
timeout 120 python tools/trace.py --m 2048 --n 4096 --k 4096 --units 80 --cta -1 > gpurun_out/r2/trace_2048_4096_b43.txt 2>&1
echo done
