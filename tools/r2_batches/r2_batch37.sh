set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster or decomposition or repeatab" 2>&1 | tail -8 > gpurun_out/r2/tests_gpu_b37a.txt
grep -q "passed" gpurun_out/r2/tests_gpu_b37a.txt && ! grep -q "failed" gpurun_out/r2/tests_gpu_b37a.txt || { echo "cluster tests failed, stopping"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/tests_gpu_b37.txt
timeout 300 python tools/sweep.py --models 70b,7b --ms 1,16 --no-cublas --out gpurun_out/r2/sweep_b37.jsonl > gpurun_out/r2/sweep_b37.log 2>&1
timeout 600 python tools/decode_stack.py --model llama3-8b --batches 1,16 --chain --fuse-gate-up --clone-layers --out gpurun_out/r2/decode_l3_8b_chain_b37.jsonl > gpurun_out/r2/decode_l3_b37.log 2>&1
echo done
