#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/r2g_variants_small.log
for n in 32768 65536 131072; do
  echo "== variants at n=$n (32 sets)" | tee -a $OUT/r2g_variants_small.log; G2048_VARIANT_SETS=32 timeout 600 python scripts/kernel_variants.py run $n 6000 2>&1 | tee -a $OUT/r2g_variants_small.log
done
