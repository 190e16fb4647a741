set -x
mkdir -p gpurun_out/r2
FLEXQ_AR_UNROLL=4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 tools/ar_probe.py > gpurun_out/r2/ar_probe_tp4.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29652 tools/tp_probe.py > gpurun_out/r2/tp_probe_tp4.txt 2>&1
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29653 bench.py --gpus 4 --steps 10 --warmup 5 --no-extra > gpurun_out/r2/bench_tp4b_$name.json 2> gpurun_out/r2/bench_tp4b_$name.err
}
run mc_c1 FLEXQ_BENCH_AR_MC=1 FLEXQ_BENCH_AR_CHUNKS=1
run mc_c2 FLEXQ_BENCH_AR_MC=1 FLEXQ_BENCH_AR_CHUNKS=2
run p2p_c2 FLEXQ_BENCH_AR_MC=0 FLEXQ_BENCH_AR_CHUNKS=2
echo done
