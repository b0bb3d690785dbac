#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_env_api.py -x -q -m gpu 2>&1 | tail -4
for n in 131072 262144 524288; do
echo "== bench $n on one GPU (driver flags)"; timeout 600 python bench.py --envs $n --steps 20 --warmup 5 --no-cpu-baseline --no-config4 --e2e-steps 20 --fused-steps 0 2>$OUT/r2l_bench_$n.err | tee $OUT/r2l_bench_$n.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('us/step %.3f (min %.3f max %.3f) value %.4g  %s' % (d['ms_per_step']*1e3, d['timing']['ms_per_step_min']*1e3, d['timing']['ms_per_step_max']*1e3, d['value'], d['timing']['issue'][:40]))"
done
echo "== bench default flags"; timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('us/step %.3f frac %.3f config4 %s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['config4']))"
