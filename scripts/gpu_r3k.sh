#!/bin/bash
# 8-GPU visit: BASELINE config 3 (1,048,576 boards in total) at N = 4 and 8, driver flags, chained launches + two issuing threads
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for N in 8 4; do
  echo "== bench N=$N (driver flags)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r3k_bench_n$N.err > $OUT/r3k_bench_n$N.json
  tail -2 $OUT/r3k_bench_n$N.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open("$OUT/r3k_bench_n$N.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N=%d value %.4g  us/step %.3f (min %.3f max %.3f)  frac %.3f  checksum %s  long %.3f  plain %.3f  e2e %.4g (%.3f ms)  compact %.4g  weak %.4g (%.2f us)  fused %.4g\n   issue: %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    d["long_region"]["ms_per_step"] * 1e3, d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_compact"]["value"], w.get("value", 0), w.get("ms_per_step", 0) * 1e3, (d.get("fused") or {}).get("value", 0), d["timing"]["issue"][:100]))
PY
done 2>&1 | tee $OUT/r3k_scaling.log
nproc | tee -a $OUT/r3k_scaling.log
