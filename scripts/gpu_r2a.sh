#!/bin/bash
# Round 2, GPU visit A: parity suite, driver-style short bench (N=1), long bench, kernel variants.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee $OUT/r2a_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5 | tee $OUT/r2a_smoke.log
echo "== bench (driver flags: --steps 20 --warmup 5)"
timeout 600 python bench.py --steps 20 --warmup 5 2>$OUT/r2a_bench_short.err | tee $OUT/r2a_bench_short.json | cut -c1-1500
tail -3 $OUT/r2a_bench_short.err
echo "== bench (defaults)"
timeout 600 python bench.py --no-cpu-baseline 2>$OUT/r2a_bench_default.err | tee $OUT/r2a_bench_default.json | cut -c1-600
echo "== bench 131072 boards (strong N=8 shard) on one GPU"
timeout 600 python bench.py --envs 131072 --no-cpu-baseline --e2e-steps 20 2>$OUT/r2a_bench_131k.err | tee $OUT/r2a_bench_131k.json | cut -c1-600
echo "== variants"
timeout 900 python scripts/kernel_variants.py run 2>&1 | tee $OUT/r2a_variants.log
