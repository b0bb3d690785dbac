#!/bin/bash
# Quick GPU visit: parity tests + a short device-only bench line (no CPU baseline).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10000 --warmup 200 --e2e-steps 50 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_quick.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('value %.4g steps/s  %.2f us/step  frac %.3f  e2e %.4g  clocks %s' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value'], d['clocks']))
"
