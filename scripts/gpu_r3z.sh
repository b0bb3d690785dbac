#!/bin/bash
# N = 2: the packed-wire host call for several unpack-thread counts (each rank owns half of the host's CPUs)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
for T in 0 2 3 5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-weak --fused-steps 0 --no-plain --long-region 0 --spinup 2048 --unpack-threads $T 2>$OUT/r3z.err > $OUT/r3z.json
  python - <<PY
import json
d = json.loads(open("$OUT/r3z.json").read().strip().splitlines()[-1])
print("N=2 unpack threads $T: e2e %.4g (%.3f ms)  plain wire %.4g (%.3f ms)  compact %.4g (%.3f ms)" % (
    d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_plain_wire"]["value"], d["e2e_plain_wire"]["ms_per_step"], d["e2e_compact"]["value"], d["e2e_compact"]["ms_per_step"]))
PY
done 2>&1 | tee $OUT/r3z_e2e_n2.log
