#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/single_env_latency.py 2>&1 | grep -v Warn | tee gpurun_out/r2m_single_env.log
