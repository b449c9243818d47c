mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_staleness.py -q -p no:cacheprovider 2>&1 | tail -12
timeout 600 python scripts/train_bench.py --batch 32 --steps 5 2>&1 | tail -1 | tee gpurun_out/r02_train_n1_tc.json
PDAE_TRAIN_TC_FWD=0 timeout 600 python scripts/train_bench.py --batch 32 --steps 5 2>&1 | tail -1
