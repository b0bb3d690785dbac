#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for sel in "two_host_threads" "survives" "raw_abi"; do
  echo "== racecheck: $sel"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all \
      python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "$sel" 2>&1 | tail -25
done 2>&1 | tee $OUT/r3m_racecheck.log
