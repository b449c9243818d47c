#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in test_gpu_conv_tc test_gpu_parity test_gpu_diffusion; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -12 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b256.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/bench_b256.log
cat gpurun_out/summary.txt
