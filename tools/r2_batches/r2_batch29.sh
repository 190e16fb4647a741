set -x
mkdir -p gpurun_out/r2
python tools/trace.py --m 16 --n 4096 --k 4096 --units 12 --cta -1 --noflush > gpurun_out/r2/trace_16_4096_hot_b29.txt 2>&1
python tools/trace.py --m 256 --n 4096 --k 4096 --units 40 --cta -1 > gpurun_out/r2/trace_256_4096_b29.txt 2>&1
python tools/trace.py --m 128 --n 4096 --k 4096 --units 40 --cta -1 > gpurun_out/r2/trace_128_4096_b29.txt 2>&1
echo done
