set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2/tests_gpu_all_tp2_final.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_tp2_final2.json 2> gpurun_out/r2/bench_tp2_final2.err
echo done
