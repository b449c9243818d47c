# A/B: conv_tc3 64-wide split mode: two MMA-issuer warps (PDAE_TC3_DUAL) x halo pipeline depth (PDAE_TC3_SA)
for spec in "256 64 64 64 0 64 1 0 0" "256 64 64 64 64 64 1 0 0" "256 64 64 64 0 64 1 0 1" "256 64 64 128 64 64 1 0 0"; do
  for m in 0 1; do for sa in 2 3; do echo "DUAL=$m SA=$sa $spec: $(PDAE_TC3_DUAL=$m PDAE_TC3_SA=$sa timeout 100 python scripts/conv3_bench.py $spec 2>&1 | tail -1 | sed 's/.*SB=[-0-9]*: //')"; done; done
done
PDAE_TC3_DBG=1 PDAE_TC3_DUAL=1 PDAE_TC3_SA=3 timeout 100 python scripts/conv3_bench.py 256 64 64 64 0 64 1 0 0 2>&1 | grep "tc3 dbg" | tail -1
