#!/bin/bash
# bench.py at N = 1 and N = 2 (driver flags) with chained launches and two issuing threads
set -u
mkdir -p gpurun_out
OUT=gpurun_out
NG=${1:-2}
for N in 1 $NG; do
  echo "== bench N=$N (driver flags)"
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r3j_bench_n$N.err > $OUT/r3j_bench_n$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r3j_bench_n$N.err > $OUT/r3j_bench_n$N.json
  fi
  tail -2 $OUT/r3j_bench_n$N.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open("$OUT/r3j_bench_n$N.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
c4 = d.get("config4") or {}
print("N=%d value %.4g  us/step %.3f (min %.3f max %.3f)  frac %.3f  checksum %s  long %.3f  plain %.3f  e2e %.4g (%.3f ms)  compact %.4g  weak %.4g (%.2f us)  fused %.4g c4 %.2f us (%.3f)\n   issue: %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    d["long_region"]["ms_per_step"] * 1e3, d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_compact"]["value"], w.get("value", 0), w.get("ms_per_step", 0) * 1e3, (d.get("fused") or {}).get("value", 0), c4.get("us_per_step", 0), c4.get("roofline_frac", 0), d["timing"]["issue"][:100]))
PY
done 2>&1 | tee $OUT/r3j_scaling.log
