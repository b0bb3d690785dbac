#!/bin/bash
# last multi-rank sanity check of the session: N = 2, driver flags, final library; plus the nibble / packed host tests
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_env_api.py tests/test_gpu_parity.py -x -q -m gpu -k "host_stepped or two_real_gpus" 2>&1 | tail -3
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r3w_bench_n$N.err > $OUT/r3w_bench_n$N.json
tail -2 $OUT/r3w_bench_n$N.err | cut -c1-200
python - <<PY
import json
d = json.loads(open("$OUT/r3w_bench_n$N.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N=%d value %.4g  us/step %.3f (min %.3f max %.3f)  frac %.3f  checksum %s  long %.3f  plain %.3f | e2e %.4g (%.3f ms)  plain wire %.4g (%.3f ms)  compact %.4g | weak %.4g (%.2f us) fused %.4g" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["timing"]["ms_per_step_min"] * 1e3, d["timing"]["ms_per_step_max"] * 1e3, d["roofline"]["frac"], d["state_checksum"],
    d["long_region"]["ms_per_step"] * 1e3, d["plain_launches"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_plain_wire"]["value"], d["e2e_plain_wire"]["ms_per_step"], d["e2e_compact"]["value"], w.get("value", 0), w.get("ms_per_step", 0) * 1e3, d["fused"]["value"]))
PY
