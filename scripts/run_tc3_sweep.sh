mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc3.py tests/test_gpu_parity.py -q -x -p no:cacheprovider 2>&1 | tail -3
python scripts/conv3_bench.py 256 64 64 64 0 64 1 0 0 2>&1 | tail -1
python scripts/conv3_bench.py 256 64 64 64 64 64 1 0 0 2>&1 | tail -1
python scripts/conv3_bench.py 256 64 64 64 0 64 1 0 1 2>&1 | tail -1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; tail -c 300 gpurun_out/r02_bench4.err; head -c 200 gpurun_out/r02_bench4.json
