#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 > gpurun_out/test_gpu_parity.log 2>&1; echo "parity exit $?"; tail -6 gpurun_out/test_gpu_parity.log
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
PDAE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 150 -c 8 -o gpurun_out/prof_conv_tc2_r1 -f python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_celeba64.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_celeba64.log | cut -c1-600
timeout 900 python bench.py --steps 1 --warmup 3 --workload ffhq128 --batch 32 --no-cpu-baseline > gpurun_out/bench_ffhq128.log 2>&1; echo "bench ffhq exit $?"; tail -1 gpurun_out/bench_ffhq128.log | cut -c1-400
