#!/bin/bash
# Round 2, session 2: what would a finer-grained dependency between consecutive step launches gain?  (G2048_NOWAIT
# experiment builds: no griddepcontrol.wait — valid here only because the 8 env sets are independent.)
set -u
mkdir -p gpurun_out
echo "== variants 1Mi"; G2048_VARIANT_SETS=8 timeout 600 python scripts/kernel_variants.py run 2>&1 | tee gpurun_out/r3a_variants.log
echo "== variants 262144"; G2048_VARIANT_SETS=32 timeout 600 python scripts/kernel_variants.py run 262144 6000 2>&1 | tee gpurun_out/r3a_variants_262k.log
