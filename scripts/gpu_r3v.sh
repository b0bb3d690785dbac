#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== host env tests"; timeout 900 python -m pytest tests/test_gpu_env_api.py -x -q -m gpu -k "host_stepped" 2>&1 | tail -3
for v in tail0 tail12 tail8 tail6 tail4; do
  echo "== $v"; G2048_SO=$PWD/gym-2048_b200/variants/libg2048_$v.so timeout 600 python scripts/e2e_wire_sweep.py 1048576 60 2>&1 | grep -v "cpus of"
done | tee gpurun_out/r3v_e2e_tail.log
