#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== host env tests"; timeout 600 python -m pytest tests/test_gpu_env_api.py -x -q -m gpu -k "host_stepped" 2>&1 | tail -3
echo "== e2e sweep: split copies (default build)"; timeout 600 python scripts/e2e_sweep.py 2>&1 | grep -v Warn | tee $OUT/r2o_e2e_split.log
echo "== e2e sweep: all copies of a slice on its own stream"; G2048_SO=$PWD/gym-2048_b200/variants/libg2048_nosplit.so timeout 600 python scripts/e2e_sweep.py 2>&1 | grep -v Warn | tee $OUT/r2o_e2e_nosplit.log
