#!/bin/bash
# Round 2, GPU visit B: pipe microbenchmark, launch-rate diagnostics, ncu captures of the step-kernel specialisations.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pipes2" ; timeout 300 ./scripts/micro/pipes2 2>&1 | tee $OUT/r2b_pipes2.log
echo "== launch rate" ; timeout 600 python scripts/launch_rate.py 2>&1 | grep -v Warning | tee $OUT/r2b_launch_rate.log
echo "== bench (driver flags)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/r2b_bench.err | tee $OUT/r2b_bench.json | cut -c1-300
for k in lean mask all eprun; do
  echo "== ncu full: $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:g2048_step_kernel -s 40 -c 2 \
      -f -o $OUT/r2b_step_$k python scripts/profile_kernels.py $k > $OUT/r2b_ncu_$k.log 2>&1
  tail -2 $OUT/r2b_ncu_$k.log
done
echo "== ncu full: mask at 262144 boards (config 4)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:g2048_step_kernel -s 40 -c 2 \
    -f -o $OUT/r2b_step_mask_262k python scripts/profile_kernels.py mask 262144 > $OUT/r2b_ncu_mask262k.log 2>&1
echo "== ncu launch list of the bench"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 17000 -c 300 --csv \
    --log-file $OUT/r2b_launches.csv python bench.py --steps 20 --warmup 5 --repeats 5 --e2e-steps 3 --fused-steps 4 --no-cpu-baseline > $OUT/r2b_bench_under_ncu.log 2>&1
ls -la $OUT | tail -20
