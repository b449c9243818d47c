#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_conv_tc test_gpu_parity test_gpu_diffusion test_gpu_training; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
grep -h "largest\|worst" gpurun_out/test_gpu_training.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_2gpu.log | cut -c1-700
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_ref.log | cut -c1-500
cat gpurun_out/summary.txt
