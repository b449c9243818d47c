#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc2 --csv --log-file gpurun_out/conv_tc2_traffic.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 400 python scripts/profile_ops.py celeba64 256 40 > gpurun_out/profile_ops_celeba64.txt 2>&1; sed -n 1,12p gpurun_out/profile_ops_celeba64.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_celeba64.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_celeba64.log | cut -c1-330
