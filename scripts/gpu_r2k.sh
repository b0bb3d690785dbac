#!/bin/bash
# Round 2, second 8-GPU visit: BASELINE config 3 sharded over 1/2/4/8 GPUs with the final library.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== bench test"; timeout 900 python -m pytest tests/test_gpu_env_api.py -q -m gpu -k "bench_prints" 2>&1 | tail -3
for N in 1 2 4 8; do
  echo "== bench N=$N (driver flags)"
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/r2k_bench_n$N.err > $OUT/r2k_bench_n$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r2k_bench_n$N.err > $OUT/r2k_bench_n$N.json
  fi
  python - <<PY
import json
d = json.loads(open("$OUT/r2k_bench_n$N.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N=%d value %.4g  us/step %.3f  frac %.3f  checksum %s  e2e %.4g (%.3f ms)  e2e_compact %.4g  weak %.4g (%.2f us)  fused %.4g  issue: %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["state_checksum"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_compact"]["value"], w.get("value", 0), w.get("ms_per_step", 0) * 1e3, (d.get("fused") or {}).get("value", 0), d["timing"]["issue"][:28]))
PY
done 2>&1 | tee $OUT/r2k_scaling.log
echo "== reference arm N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>/dev/null | cut -c1-300 | tee $OUT/r2k_ref_n8.json
