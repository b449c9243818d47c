#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in test_gpu_conv_tc test_gpu_parity test_gpu_diffusion; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -15 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_b256.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/bench_b256.log
# launch list of 2 decoder steps (second one is warm) + full capture of the dominant kernel
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?" >> gpurun_out/summary.txt
PDAE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 140 -c 6 -o gpurun_out/prof_conv_tc_r1 -f python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
