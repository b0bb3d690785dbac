#!/bin/bash
# bench.py headline, driver flags, 1 / 2 / 3 issuing threads per shard size
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for envs in 131072 262144 524288 1048576; do
for T in 1 2 4; do
  timeout 900 python bench.py --gpus 1 --envs $envs --steps 20 --warmup 5 --issue-threads $T --small-below 2097152 --no-config4 --fused-steps 0 --e2e-steps 3 --no-cpu-baseline 2>$OUT/r3f.err > $OUT/r3f.json || tail -5 $OUT/r3f.err
  python - <<PY
import json
d = json.loads(open("$OUT/r3f.json").read().strip().splitlines()[-1])
p = d.get("plain_launches") or {}
l = d.get("long_region") or {}
print("$envs T=$T: us/step %.3f (min %.3f max %.3f) frac %.3f  long region %.3f us  plain %.3f us  checksum %s" % (
    d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"],
    l.get("ms_per_step", 0) * 1e3, p.get("ms_per_step", 0) * 1e3, d["state_checksum"]))
PY
done; done 2>&1 | tee $OUT/r3f_issue_threads.log
