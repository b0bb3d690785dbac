#!/bin/bash
# Session 2 validation visit: full parity suite, smoke, both bench arms (driver flags), sanitizer over the chained-launch tests
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest -m gpu"; timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/r3l_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee $OUT/r3l_smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>$OUT/r3l_bench_ref.err | tee $OUT/r3l_bench_ref.json | cut -c1-200
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 2>$OUT/r3l_bench.err | tee $OUT/r3l_bench.json | cut -c1-200; tail -3 $OUT/r3l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3l_bench.json").read().strip().splitlines()[-1])
print("us/step %.3f (span %.3f, min %.3f max %.3f) frac %.3f  long %.3f (max %.3f)  plain %.3f  e2e %.4g (%.3f ms) compact %.4g config4 %.2f us (%.3f) fused %.4g checksum %s clocks %s cpu %s" % (
    d["ms_per_step"] * 1e3, (d["timing"]["ms_per_step_all_regions_span"] or 0) * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"],
    d["long_region"]["ms_per_step"] * 1e3, d["long_region"]["ms_per_step_max"] * 1e3,
    d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_compact"]["value"],
    d["config4"]["us_per_step"], d["config4"]["roofline_frac"], d["fused"]["value"], d["state_checksum"], d["clocks"], d["cpu_baseline"]["value"]))
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool (chained launches)"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
      python -m pytest tests/test_gpu_chain.py -x -q -m gpu -k "1000 or 70001 or raw_abi or survives or in_kernel_policy or two_host_threads" 2>&1 | tail -6 | tee $OUT/r3l_sanitizer_chain_$tool.log
done
