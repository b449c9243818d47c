#!/bin/bash
# Round-2 validation + evidence capture on one B200: tests, smoke, bench (both arms, other workloads), ncu launch lists,
# `--set full` captures of representative conv_tc3 shapes, DRAM traffic of the conv_tc3 launches of a decoder step.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_smoke.log
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"; head -c 250 gpurun_out/r02_bench_final.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_reference_final.json; echo
timeout 900 python bench.py --workload ffhq128 --steps 1 --warmup 3 --no-extras > gpurun_out/r02_bench_ffhq128_b64.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_ffhq128_b64.json; echo
timeout 900 python bench.py --workload ffhq256 --steps 1 --warmup 3 --no-extras > gpurun_out/r02_bench_ffhq256_b8.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_ffhq256_b8.json; echo
KREGEX='regex:conv_|gn_|split3|softmax|gemm_batched|ch_stats|timestep_embedding|stem_conv|transpose_v|ddim|zero_kernel|attention'
for m in bf16x3 bf16; do
  PDAE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r02_launches_celeba64_b256_$m.csv python scripts/ncu_step.py celeba64 256 $m 2 > gpurun_out/ncu_list_$m.log 2>&1; echo "ncu list $m exit $? $(wc -l < gpurun_out/r02_launches_celeba64_b256_$m.csv) lines"
  PDAE_NO_GRAPH=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc3 --csv --log-file gpurun_out/r02_conv_tc3_traffic_$m.csv python scripts/ncu_step.py celeba64 256 $m 2 > gpurun_out/ncu_traffic_$m.log 2>&1; echo "ncu traffic $m exit $?"
done
# --set full on representative shapes (one launch each)
i=0
for spec in "256 64 64 64 0 64 0 0 0" "256 64 64 128 0 128 0 1 0" "256 32 32 256 0 256 0 1 0" "256 64 64 64 64 64 1 0 0" "256 64 64 128 0 128 1 1 0" "256 16 16 512 0 256 1 0 0"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 4 -c 1 -o gpurun_out/r02_tc3_full_$i python scripts/conv3_bench.py $spec > gpurun_out/ncu_full_$i.log 2>&1; echo "ncu full $i ($spec) exit $?"
done
for m in bf16 bf16x3; do timeout 300 python scripts/profile_ops.py celeba64 256 70 $m > gpurun_out/r02_ops_$m.txt 2>&1; head -8 gpurun_out/r02_ops_$m.txt; done
