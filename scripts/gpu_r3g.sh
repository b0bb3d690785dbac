#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for v in n1 n2 n3; do
for T in 2; do
  G2048_SO=$PWD/gym-2048_b200/variants/libg2048_$v.so timeout 900 python bench.py --gpus 1 --envs 1048576 --steps 20 --warmup 5 --issue-threads $T --small-below 2097152 --no-config4 --fused-steps 0 --e2e-steps 3 --no-cpu-baseline 2>$OUT/r3g.err > $OUT/r3g.json || tail -5 $OUT/r3g.err
  python - <<PY
import json
d = json.loads(open("$OUT/r3g.json").read().strip().splitlines()[-1])
p = d.get("plain_launches") or {}
l = d.get("long_region") or {}
print("$v T=$T: us/step %.3f (min %.3f max %.3f) frac %.3f  long region %.3f us  plain %.3f us  checksum %s" % (
    d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"],
    l.get("ms_per_step", 0) * 1e3, p.get("ms_per_step", 0) * 1e3, d["state_checksum"]))
PY
done; done 2>&1 | tee $OUT/r3g_shapes_t2.log
