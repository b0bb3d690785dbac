#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pipes2"; timeout 600 ./scripts/micro/pipes2 2>&1 | tee $OUT/r2j_pipes2.log
