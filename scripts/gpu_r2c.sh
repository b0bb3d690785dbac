#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest (new tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/r2c_pytest.log
echo "== variants"; timeout 900 python scripts/kernel_variants.py run 2>&1 | tee $OUT/r2c_variants.log
echo "== bench (driver flags)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r2c_bench.err | tee $OUT/r2c_bench.json | cut -c1-200; tail -3 $OUT/r2c_bench.err
