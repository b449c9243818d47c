mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; tail -6 gpurun_out/r02_pytest4.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -c 300 gpurun_out/r02_bench3.err; head -c 300 gpurun_out/r02_bench3.json
