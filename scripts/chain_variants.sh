#!/bin/bash
# Time every library build under gym-2048_b200/variants/ (scripts/kernel_variants.py build ...) with scripts/bench_chain.py.
set -u
mkdir -p gpurun_out
for so in gym-2048_b200/variants/libg2048_*.so; do
  name=$(basename $so .so); name=${name#libg2048_}
  echo "== $name"
  G2048_SO=$PWD/$so G2048_CHAIN_SIZES=${G2048_CHAIN_SIZES:-1048576,262144,131072} timeout 300 python scripts/bench_chain.py 32 3000 2>&1
done | tee gpurun_out/chain_variants.log
