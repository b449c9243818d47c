#!/bin/bash
# first GPU pass: each test file in its own process (a trapped kernel poisons the CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_conv_tc test_gpu_parity test_gpu_diffusion; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -25 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 1 --warmup 3 --batch 32 --ddim-steps 10 > gpurun_out/bench_small.log 2>&1; echo "bench_small exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/bench_small.log
cat gpurun_out/summary.txt
