#!/bin/bash
# 8-GPU visit: N = 8 only, topology-aware core pinning on / off
set -u
mkdir -p gpurun_out
OUT=gpurun_out
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|Model name|NUMA" | tee $OUT/r3o_scaling.log
cat /sys/devices/system/cpu/cpu0/topology/thread_siblings_list /sys/devices/system/cpu/cpu1/topology/thread_siblings_list | tr '\n' ' ' | tee -a $OUT/r3o_scaling.log; echo | tee -a $OUT/r3o_scaling.log
for pin in 1 0; do
  N=8
  echo "== bench N=$N pin=$pin (driver flags)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --pin-cores $pin --no-weak --fused-steps 0 --e2e-steps 20 2>$OUT/r3o_bench_n$N.err > $OUT/r3o_bench_n${N}_pin$pin.json
  tail -2 $OUT/r3o_bench_n$N.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open("$OUT/r3o_bench_n${N}_pin$pin.json").read().strip().splitlines()[-1])
print("N=%d pin=$pin value %.4g  us/step %.3f (min %.3f max %.3f mean %.3f)  frac %.3f  checksum %s  long %.3f (max %.3f)  plain %.3f  e2e %.4g (%.3f ms)" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["timing"]["ms_per_step_mean"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    d["long_region"]["ms_per_step"] * 1e3, d["long_region"]["ms_per_step_max"] * 1e3, d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done 2>&1 | tee -a $OUT/r3o_scaling.log
