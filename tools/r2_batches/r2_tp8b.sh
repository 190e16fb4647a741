set -x
mkdir -p gpurun_out/r2
for u in 4 8; do
FLEXQ_AR_UNROLL=$u timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/ar_probe.py > gpurun_out/r2/ar_probe_tp8_u$u.txt 2>&1
done
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 5 --xbits 8 --no-extra > gpurun_out/r2/bench_tp8b_$name.json 2> gpurun_out/r2/bench_tp8b_$name.err
}
run u8r8 FLEXQ_AR_UNROLL=8 FLEXQ_BENCH_AR_RESERVE=8
run u8r16 FLEXQ_AR_UNROLL=8 FLEXQ_BENCH_AR_RESERVE=16
echo done
