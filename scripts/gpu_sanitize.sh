#!/bin/bash
# compute-sanitizer passes over a small, kernel-covering subset of the GPU tests (round 2: + the single-env kernel,
# the in-kernel policies, step_n / step_list, the compact host format, every compile-time output set).
set -u
mkdir -p gpurun_out
SEL='special_boards or add_tile or crossing or symmetries or sample_actions or gae_kernel or out_of_place or status_and_move or step_many or observe_all or converters or TestBoard or TestStep or in_kernel_policy or step_schedule or step_n_is or nibble or every_output_set or stack_function'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 --target-processes all \
      python -m pytest tests -x -q -m gpu -k "$SEL" 2>&1 | tail -8 | tee gpurun_out/r2_sanitizer_$tool.log
done
