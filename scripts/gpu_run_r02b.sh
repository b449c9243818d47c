#!/bin/bash
# Round-2 final validation + evidence on one B200: tests, smoke, bench (both arms), training step (bench, per-plan profile,
# ncu launch list), wgrad_tc micro-benchmark + one `--set full` capture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_smoke.log
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"; head -c 250 gpurun_out/r02_bench_final.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; head -c 200 gpurun_out/r02_bench_reference_final.json; echo
for b in 32 64; do timeout 300 python scripts/train_bench.py --batch $b 2>/dev/null | tail -1 > gpurun_out/r02_train_step_n1_b$b.json; head -c 220 gpurun_out/r02_train_step_n1_b$b.json; echo; done
timeout 300 python scripts/train_profile.py > gpurun_out/r02_train_profile.txt 2>&1; tail -5 gpurun_out/r02_train_profile.txt
timeout 200 python scripts/wgrad_bench.py > gpurun_out/r02_wgrad_bench.txt 2>&1; cat gpurun_out/r02_wgrad_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step_b32.csv python scripts/ncu_train_step.py > gpurun_out/ncu_train.log 2>&1; echo "ncu train list exit $? $(wc -l < gpurun_out/r02_launches_train_step_b32.csv) lines"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 40 -c 1 -o gpurun_out/r02_wgrad_tc_full python scripts/wgrad_bench.py > gpurun_out/ncu_wgrad_full.log 2>&1; echo "ncu wgrad full exit $?"
for m in bf16x3; do timeout 300 python scripts/profile_ops.py celeba64 256 70 $m > gpurun_out/r02_ops_$m.txt 2>&1; head -4 gpurun_out/r02_ops_$m.txt; done
