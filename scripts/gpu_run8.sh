#!/bin/bash
# Full validation + evidence capture (one GPU): tests, smoke, bench (both arms), ncu launch list + DRAM traffic, per-op profile.
mkdir -p gpurun_out
for f in conv_tc parity diffusion training fullsize callers; do
  echo "== tests/test_gpu_$f.py"; timeout 900 python -m pytest tests/test_gpu_$f.py -x -q -m gpu 2>&1 | tail -2
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log > gpurun_out/bench_default.json; cut -c1-400 gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_reference.log > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc2 --csv --log-file gpurun_out/conv_tc2_traffic.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
PDAE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python scripts/ncu_step.py celeba64 256 bf16 2 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 400 python scripts/profile_ops.py celeba64 256 60 > gpurun_out/profile_ops_final.txt 2>&1; sed -n 1,14p gpurun_out/profile_ops_final.txt
