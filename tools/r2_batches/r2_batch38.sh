set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rmsnorm or silu or fused_producers or llama" 2>&1 | tail -6 > gpurun_out/r2/tests_gpu_b38.txt
timeout 600 python tools/producer_bench.py > gpurun_out/r2/producers_b38.jsonl 2> gpurun_out/r2/producers_b38.err
timeout 600 python tools/decode_stack.py --model llama3-8b --batches 1,16 --chain --fuse-gate-up --clone-layers --out gpurun_out/r2/decode_l3_8b_chain_b38.jsonl > gpurun_out/r2/decode_l3_b38.log 2>&1
echo done
