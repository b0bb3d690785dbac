#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the step kernel.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench" ; timeout 600 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 100 --warmup 20 --e2e-steps 3 --fused-steps 0 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
  echo "== ncu full capture of the step kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:g2048_step_kernel -s 60 -c 3 \
      -f -o $OUT/step_full python bench.py --steps 100 --warmup 20 --e2e-steps 3 --fused-steps 0 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
  echo "== ncu full capture of the multi-step kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:g2048_step_many_kernel -s 2 -c 1 \
      -f -o $OUT/step_many_full python bench.py --steps 100 --warmup 20 --e2e-steps 3 --fused-steps 32 --no-cpu-baseline > $OUT/ncu_many.log 2>&1
  ls -la $OUT
fi
