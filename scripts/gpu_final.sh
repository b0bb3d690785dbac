#!/bin/bash
# Final validation visit of a round: full GPU suite, smoke, both bench arms with the driver's flags, default flags.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/final_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee $OUT/final_smoke.log
echo "== bench reference arm (driver flags)"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>$OUT/final_bench_ref.err | tee $OUT/final_bench_ref.json | cut -c1-160
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 2>$OUT/final_bench.err | tee $OUT/final_bench.json | cut -c1-160; tail -3 $OUT/final_bench.err
echo "== bench (default flags)"; timeout 900 python bench.py --no-cpu-baseline 2>$OUT/final_bench_default.err | tee $OUT/final_bench_default.json | cut -c1-160
python - <<'PY'
import json
for f in ("final_bench", "final_bench_default"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "us/step %.3f frac %.3f e2e %.4g (%.3f ms) compact %.4g config4 %.2f us fused %.4g checksum %s clocks %s" % (
        d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_compact"]["value"],
        d["config4"]["us_per_step"], d["fused"]["value"], d["state_checksum"], d["clocks"]))
PY
