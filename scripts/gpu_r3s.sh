#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== host env tests"; timeout 900 python -m pytest tests/test_gpu_env_api.py -x -q -m gpu -k "host_stepped" 2>&1 | tail -4
echo "== e2e wire sweep"; timeout 900 python scripts/e2e_wire_sweep.py 1048576 60 2>&1 | tee gpurun_out/r3s_e2e_wire.log
