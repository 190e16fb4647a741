set -x
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/tp8_gpus.txt
run() { # name, env..., extra args
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 5 --xbits 8 --no-extra > gpurun_out/r2/bench_tp8_$name.json 2> gpurun_out/r2/bench_tp8_$name.err
}
run c2 FLEXQ_BENCH_AR_CHUNKS=2
run c3 FLEXQ_BENCH_AR_CHUNKS=3
run c4 FLEXQ_BENCH_AR_CHUNKS=4
run c1 FLEXQ_BENCH_AR_CHUNKS=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 10 --warmup 5 --xbits 6 --no-extra > gpurun_out/r2/bench_tp8_a6.json 2> gpurun_out/r2/bench_tp8_a6.err
echo done
