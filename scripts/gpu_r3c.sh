#!/bin/bash
# bench.py with chained launches, driver flags; Python-loop issue vs step-list issue at 1 Mi boards
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for mode in py list; do
  extra=""; [ $mode = list ] && extra="--small-below 2097152"
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 $extra 2>$OUT/r3c_bench_$mode.err > $OUT/r3c_bench_$mode.json
  python - <<PY
import json
d = json.loads(open("$OUT/r3c_bench_$mode.json").read().strip().splitlines()[-1])
p = d.get("plain_launches") or {}
c4 = d.get("config4") or {}
print("$mode: value %.4g  us/step %.3f (min %.3f max %.3f)  frac %.3f  checksum %s  plain %.3f us  c4 %.2f us (%.3f)  e2e %.4g  fused %.4g\n  issue: %s" % (
    d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    p.get("ms_per_step", 0) * 1e3, c4.get("us_per_step", 0), c4.get("roofline_frac", 0), d["e2e"]["value"], (d.get("fused") or {}).get("value", 0), d["timing"]["issue"]))
PY
done 2>&1 | tee $OUT/r3c_bench.log
