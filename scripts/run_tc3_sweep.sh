mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc3.py tests/test_gpu_parity.py -q -x -p no:cacheprovider 2>&1 | tail -3
for m in bf16 bf16x3; do timeout 300 python scripts/profile_ops.py celeba64 256 70 $m > gpurun_out/r02_ops_$m.txt 2>&1; head -12 gpurun_out/r02_ops_$m.txt; done
timeout 600 python scripts/train_bench.py --batch 32 --steps 5 2>&1 | tail -2 | tee gpurun_out/r02_train_n1.json
