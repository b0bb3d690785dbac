#!/bin/bash
# Round 2 validation visit: full parity suite, smoke, driver-style bench, extras, final ncu captures.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/r2i_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee $OUT/r2i_smoke.log
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 2>$OUT/r2i_bench.err | tee $OUT/r2i_bench.json | cut -c1-200; tail -3 $OUT/r2i_bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>$OUT/r2i_bench_ref.err | tee $OUT/r2i_bench_ref.json | cut -c1-200
echo "== extras"; timeout 900 python scripts/bench_extras.py 2>&1 | grep -v Warn | tee $OUT/r2i_extras.log
for k in lean mask all eprun; do
  echo "== ncu full: $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:g2048_step_kernel -s 40 -c 2 \
      -f -o $OUT/r2i_step_$k python scripts/profile_kernels.py $k > $OUT/r2i_ncu_$k.log 2>&1
  tail -1 $OUT/r2i_ncu_$k.log
done
echo "== ncu launch list of the bench (time share of the step kernel in the timed region)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 600 --csv \
    --log-file $OUT/r2i_launches.csv python bench.py --steps 20 --warmup 5 --repeats 25 --spinup 2048 --e2e-steps 3 --fused-steps 4 --no-cpu-baseline --no-config4 > $OUT/r2i_bench_under_ncu.log 2>&1
wc -l $OUT/r2i_launches.csv
