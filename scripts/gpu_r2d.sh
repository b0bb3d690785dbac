#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee $OUT/r2d_pytest.log
echo "== variants"; timeout 900 python scripts/kernel_variants.py run 2>&1 | tee $OUT/r2d_variants.log
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 2>$OUT/r2d_bench.err | tee $OUT/r2d_bench.json | cut -c1-200; tail -3 $OUT/r2d_bench.err
echo "== extras"; timeout 600 python scripts/bench_extras.py 2>&1 | grep -v Warn | tee $OUT/r2d_extras.log
