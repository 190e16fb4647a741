set -x
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/tp2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2/tests_gpu_multi_tp2.txt
for xb in 6 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --xbits $xb > gpurun_out/r2/bench_tp2_a$xb.json 2> gpurun_out/r2/bench_tp2_a$xb.err
done
FLEXQ_BENCH_AR_CHUNKS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --xbits 8 > gpurun_out/r2/bench_tp2_a8_c2.json 2> gpurun_out/r2/bench_tp2_a8_c2.err
echo done
