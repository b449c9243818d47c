# torchrun N=2 sanity of bench.py under different conv_tc3 settings (each: exit code + first JSON chars or the first error lines)
i=0
for envs in "PDAE_TC3_DUAL=1" "PDAE_TC3_DUAL=0" "PDAE_TC3_DUAL=1"; do
  i=$((i+1))
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus 2 --steps 1 --warmup 3 --no-extras > gpurun_out/n2_$i.json 2> gpurun_out/n2_$i.err
  echo "run $i ($envs): exit $? $(head -c 160 gpurun_out/n2_$i.json)"
  grep -m3 "rank[01]\]:.*Error\|unspecified\|illegal" gpurun_out/n2_$i.err | head -3
done
nvidia-smi --query-gpu=index,ecc.errors.uncorrected.volatile.total,ecc.errors.corrected.volatile.total --format=csv 2>&1 | head -4
dmesg 2>/dev/null | grep -i xid | tail -3
