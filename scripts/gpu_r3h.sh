#!/bin/bash
# Session 2 of round 2, validation visit: full parity suite, smoke, both bench arms with the driver's flags, ncu of the
# chained lean kernel, ncu launch list of the bench.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest -m gpu"; timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/r3h_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee $OUT/r3h_smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>$OUT/r3h_bench_ref.err | tee $OUT/r3h_bench_ref.json | cut -c1-200
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 2>$OUT/r3h_bench.err | tee $OUT/r3h_bench.json | cut -c1-200; tail -3 $OUT/r3h_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3h_bench.json").read().strip().splitlines()[-1])
print("us/step %.3f (span %.3f) frac %.3f  long %.3f  plain %.3f  e2e %.4g (%.3f ms) compact %.4g config4 %.2f us (%.3f) fused %.4g checksum %s clocks %s cpu %s" % (
    d["ms_per_step"] * 1e3, (d["timing"]["ms_per_step_all_regions_span"] or 0) * 1e3, d["roofline"]["frac"], d["long_region"]["ms_per_step"] * 1e3,
    d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_compact"]["value"],
    d["config4"]["us_per_step"], d["config4"]["roofline_frac"], d["fused"]["value"], d["state_checksum"], d["clocks"], d["cpu_baseline"]["value"]))
PY
for k in chain chain_direct; do
  echo "== ncu full: $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:g2048_step_kernel -s 40 -c 2 \
      -f -o $OUT/r3h_step_$k python scripts/profile_kernels.py $k > $OUT/r3h_ncu_$k.log 2>&1
  tail -1 $OUT/r3h_ncu_$k.log
done
echo "== ncu launch list of the bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 600 --csv \
    --log-file $OUT/r3h_launches.csv python bench.py --steps 20 --warmup 5 --repeats 25 --spinup 2048 --e2e-steps 3 --fused-steps 4 --no-cpu-baseline --no-config4 --no-plain --long-region 0 > $OUT/r3h_bench_under_ncu.log 2>&1
wc -l $OUT/r3h_launches.csv
