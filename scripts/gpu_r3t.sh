#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest (env api + bench test)"; timeout 1700 python -m pytest tests/test_gpu_env_api.py -x -q -m gpu 2>&1 | tail -4
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r3t_bench.err | tee $OUT/r3t_bench.json | cut -c1-200; tail -3 $OUT/r3t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3t_bench.json").read().strip().splitlines()[-1])
print("us/step %.3f frac %.3f | e2e %.4g (%.3f ms) plain wire %.4g (%.3f ms) compact %.4g (%.3f ms) | config4 %.2f us checksum %s" % (
    d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_plain_wire"]["value"], d["e2e_plain_wire"]["ms_per_step"],
    d["e2e_compact"]["value"], d["e2e_compact"]["ms_per_step"], d["config4"]["us_per_step"], d["state_checksum"]))
print(d["e2e"])
PY
echo "== memcheck host env"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_env_api.py -x -q -m gpu -k "packed_wire and not 1048576" 2>&1 | tail -4
