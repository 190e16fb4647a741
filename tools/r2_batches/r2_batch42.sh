set -x
mkdir -p gpurun_out/r2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_b42.txt 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_ref_b42.json 2> gpurun_out/r2/bench_ref_b42.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_b42.json 2> gpurun_out/r2/bench_b42.err
echo done
