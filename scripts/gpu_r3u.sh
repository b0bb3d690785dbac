#!/bin/bash
# 8-GPU visit with the packed-wire host call: N = 8 (and the e2e legs), driver flags
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-weak --fused-steps 0 2>$OUT/r3u_bench_n$N.err > $OUT/r3u_bench_n$N.json
tail -2 $OUT/r3u_bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("$OUT/r3u_bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.4g  us/step %.3f (min %.3f max %.3f)  frac %.3f  checksum %s  long %.3f  plain %.3f | e2e %.4g (%.3f ms)  plain wire %.4g (%.3f ms)  compact %.4g (%.3f ms)" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    d["long_region"]["ms_per_step"] * 1e3, d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_plain_wire"]["value"], d["e2e_plain_wire"]["ms_per_step"], d["e2e_compact"]["value"], d["e2e_compact"]["ms_per_step"]))
PY
