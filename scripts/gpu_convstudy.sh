#!/bin/bash
mkdir -p gpurun_out
for args in "256 64 64 64 64 3 0 1 1" "256 64 64 64 64 3 0 1 0" "256 64 64 64 64 3 0 0 0" "256 64 64 64 64 3 1 0 1" "256 64 64 64 64 3 1 0 0" "256 64 64 64 64 3 0 0 0 0 v1" "256 64 64 64 64 3 1 0 0 0 v1" "256 16 16 256 256 1 1 0 1" "256 16 16 256 256 1 1 0 0" "256 16 16 256 256 1 0 0 0" "256 16 16 256 256 1 1 0 0 128" "256 16 16 256 256 1 1 0 0 0 v1" "256 32 32 128 128 3 1 0 1" "256 32 32 128 128 3 1 0 0 0 v1"; do
  timeout 120 python scripts/conv_bench.py $args 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 8 -c 2 -o gpurun_out/prof_tc2_small -f python scripts/conv_bench.py 256 64 64 64 64 3 0 1 1 > gpurun_out/ncu_tc2_small.log 2>&1
echo "ncu exit $?"
