mkdir -p gpurun_out
for m in bf16 bf16x3; do timeout 300 python scripts/profile_ops.py celeba64 256 70 $m > gpurun_out/r02_ops_$m.txt 2>&1; head -12 gpurun_out/r02_ops_$m.txt; done
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02_pytest3.log 2>&1; tail -4 gpurun_out/r02_pytest3.log
