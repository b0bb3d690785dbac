#!/bin/bash
# compute-sanitizer passes over a small, kernel-covering subset of the GPU tests + e2e chunking sweep.
set -u
mkdir -p gpurun_out
SEL='special_boards or add_tile or crossing or symmetries or sample_actions or gae_kernel or out_of_place or status_and_move or step_many or observe_all or converters'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
      python -m pytest tests -x -q -m gpu -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/sanitizer_$tool.log
done
for c in 1 2 3 4; do
  python bench.py --steps 2000 --warmup 50 --e2e-steps 100 --e2e-chunks $c --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('chunks $c  e2e %.4g steps/s  %.3f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))" | tee -a gpurun_out/e2e_chunks.log
done
