#!/bin/bash
# bench.py with its DEFAULT flags (K = 200, W = 50): run time and the line
set -u
mkdir -p gpurun_out
time (timeout 900 python bench.py 2>gpurun_out/r3y_bench_default.err > gpurun_out/r3y_bench_default.json)
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3y_bench_default.json").read().strip().splitlines()[-1])
print("default flags: steps %d warmup %d repeats %d | us/step %.3f (min %.3f max %.3f) frac %.3f long %.3f plain %.3f | e2e %.4g (%.3f ms) plain wire %.4g compact %.4g | c4 %.2f us (%.3f) fused %.4g | cpu %.4g on %d cores | checksum %s launches %d" % (
    d["steps"], d["warmup"], d["repeats"], d["ms_per_step"]*1e3, d["timing"]["ms_per_step_min"]*1e3, d["timing"]["ms_per_step_max"]*1e3, d["roofline"]["frac"],
    d["long_region"]["ms_per_step"]*1e3, d["plain_launches"]["ms_per_step"]*1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_plain_wire"]["value"], d["e2e_compact"]["value"],
    d["config4"]["us_per_step"], d["config4"]["roofline_frac"], d["fused"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["state_checksum"], d["gpu_launches"]))
PY
