set -x
mkdir -p gpurun_out/r2
for rep in 1 2; do
for pct in 70 55; do
  FLEXQ_ALIGN_SM_PCT=$pct timeout 300 python tools/sweep.py --models 7b,l3-8b --ms 1,16,32 --no-cublas --out gpurun_out/r2/sweep_b36_pct${pct}_$rep.jsonl > gpurun_out/r2/sweep_b36_pct${pct}_$rep.log 2>&1
done
done
echo done
