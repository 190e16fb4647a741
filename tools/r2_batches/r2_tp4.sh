set -x
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2/bench_tp4_final.json 2> gpurun_out/r2/bench_tp4_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_tp2_final.json 2> gpurun_out/r2/bench_tp2_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 4 --steps 3 --warmup 1 --impl reference > gpurun_out/r2/bench_ref_tp4.json 2> gpurun_out/r2/bench_ref_tp4.err
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2/tests_gpu_multi_final.txt
echo done
