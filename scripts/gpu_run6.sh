#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_conv_tc test_gpu_parity test_gpu_diffusion test_gpu_training; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
grep -h "\[bf16\]" gpurun_out/test_gpu_parity.log | sort -t' ' -k5 -g | tail -12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 400 python scripts/profile_ops.py celeba64 256 12 2>&1 | sed -n 1,28p
PDAE_STREAM_BF16=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "oracle_agrees" 2>&1 | grep "bf16\]\|passed\|failed"
cat gpurun_out/summary.txt
