timeout 400 python -m pytest tests/test_gpu_conv_tc3.py -x -q 2>&1 | tail -3
for spec in "256 64 64 128 0 128 1 1 0" "256 32 32 256 0 256 1 1 0" "256 16 16 512 0 512 1 1 0" "256 32 32 256 128 128 1 0 0" "256 32 32 128 0 128 1 0 1"; do
  for m in 0 1; do echo "MRG128=$m $spec: $(PDAE_TC3_MRG128=$m timeout 100 python scripts/conv3_bench.py $spec 2>&1 | tail -1)"; done
done
