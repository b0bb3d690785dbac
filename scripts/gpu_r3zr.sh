#!/bin/bash
# compute-sanitizer racecheck over the kernel-covering subset of the GPU suite with the FINAL library of round 2
set -u
mkdir -p gpurun_out
SEL='special_boards or add_tile or crossing or symmetries or sample_actions or gae_kernel or out_of_place or status_and_move or step_many or observe_all or converters or TestBoard or TestStep or in_kernel_policy or step_schedule or step_n_is or nibble or every_output_set or stack_function or (chained and (1000 or 70001)) or raw_abi or survives or (packed_wire and (255 or 70001))'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all \
    python -m pytest tests -x -q -m gpu -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/r3zz_sanitizer_racecheck.log
