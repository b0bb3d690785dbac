#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== chained launches"; timeout 600 python scripts/bench_chain.py 32 4000 2>&1 | tee gpurun_out/r3b_chain.log
