set -x
mkdir -p gpurun_out/r2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench_tp8_final.json 2> gpurun_out/r2/bench_tp8_final.err
echo done
