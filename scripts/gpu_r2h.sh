#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== variants 1Mi"; G2048_VARIANT_SETS=8 timeout 600 python scripts/kernel_variants.py run 2>&1 | tee $OUT/r2h_variants.log
