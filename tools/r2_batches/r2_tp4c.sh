set -x
mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/topo_tp4.txt 2>&1
for nm in 1 0; do
FLEXQ_BENCH_NUMA=$nm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2968$nm bench.py --gpus 4 --steps 10 --warmup 5 > gpurun_out/r2/bench_tp4c_numa$nm.json 2> gpurun_out/r2/bench_tp4c_numa$nm.err
done
echo done
