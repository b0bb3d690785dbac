#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== chained launches"; timeout 600 python scripts/bench_chain.py 32 4000 2>&1 | tee $OUT/r3i_chain.log
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r3i_bench.err | tee $OUT/r3i_bench.json | cut -c1-200; tail -3 $OUT/r3i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3i_bench.json").read().strip().splitlines()[-1])
print("us/step %.3f (span %.3f, min %.3f max %.3f) frac %.3f  long %.3f  plain %.3f  e2e %.4g (%.3f ms) compact %.4g config4 %.2f us (%.3f) fused %.4g checksum %s" % (
    d["ms_per_step"] * 1e3, (d["timing"]["ms_per_step_all_regions_span"] or 0) * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["long_region"]["ms_per_step"] * 1e3,
    d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_compact"]["value"],
    d["config4"]["us_per_step"], d["config4"]["roofline_frac"], d["fused"]["value"], d["state_checksum"]))
PY
