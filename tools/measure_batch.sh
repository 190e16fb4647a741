set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/tests_gpu.txt
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
python tools/sweep.py --models 70b,7b,l3-8b --out gpurun_out/sweep_r1.jsonl > gpurun_out/sweep_r1.log 2>&1
python tools/decode_stack.py --model llama2-70b --layers 40 --fuse-gate-up --batches 1,4,16 > gpurun_out/decode_70b.jsonl 2> gpurun_out/decode_70b.err
python tools/decode_stack.py --model llama3-8b > gpurun_out/decode_l3_8b.jsonl 2> gpurun_out/decode_l3_8b.err
python tools/decode_stack.py --model llama3-8b --fuse-gate-up > gpurun_out/decode_l3_8b_fused.jsonl 2> gpurun_out/decode_l3_8b_fused.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/prof_prefill_r1 -f python tools/run_case.py --m 2048 --n 28672 --k 8192 --iters 3 > gpurun_out/ncu_p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:w6ax_gemm -c 1 -s 2 -o gpurun_out/prof_decode_r1 -f python tools/run_case.py --m 16 --n 28672 --k 8192 --iters 3 > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quant_act_native -c 1 -s 2 -o gpurun_out/prof_quant_r1 -f python tools/run_case.py --m 2048 --n 8192 --k 28672 --xb 8 --iters 3 --fused > gpurun_out/ncu_q.log 2>&1
echo done
for v in NOEPI NOMATH NOLD; do echo "VARIANT $v"; FLEXQ_B200_LIB=$PWD/tools/ubench/ab/lib_$v.so python tools/sweep.py --models 70b --ms 2048 --no-cublas --out gpurun_out/sweep_exp_$v.jsonl 2>&1 | tail -3; done > gpurun_out/experiments_r1.log 2>&1
python tools/producer_bench.py > gpurun_out/producers_r1.jsonl 2>&1
python tools/decode_stack.py --model llama2-70b --layers 40 --fuse-gate-up --chain --batches 1,4,16 > gpurun_out/decode_70b_chain.jsonl 2> gpurun_out/decode_70b_chain.err
python tools/decode_stack.py --model llama3-8b --fuse-gate-up --chain --batches 1,4,16 > gpurun_out/decode_l3_8b_chain.jsonl 2> gpurun_out/decode_l3_8b_chain.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1
