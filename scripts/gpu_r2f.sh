#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for n in 131072 262144 524288; do
  echo "== variants at n=$n (32 sets)"; G2048_VARIANT_SETS=32 timeout 600 python scripts/kernel_variants.py run $n 6000 2>&1 | tee -a $OUT/r2f_variants_small.log
done
echo "== bench 131072 (strong N=8 shard) on one GPU"; timeout 600 python bench.py --envs 131072 --steps 20 --warmup 5 --no-cpu-baseline --no-config4 --e2e-steps 20 2>$OUT/r2f_bench_131k.err | tee $OUT/r2f_bench_131k.json | cut -c1-250
