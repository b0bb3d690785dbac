#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== timeline with the prologue prefetch"; timeout 300 ./scripts/micro/timeline 2>&1 | cut -c1-2000 | tee gpurun_out/r2p_timeline.log
echo "== timeline without it"; timeout 300 ./scripts/micro/timeline_nopf 2>&1 | cut -c1-2000 | tee gpurun_out/r2p_timeline_nopf.log
