#!/bin/bash
# Round 2, 8-GPU visit: BASELINE config 3 sharded over 1/2/4/8 GPUs (strong scaling + state checksums), the 2-GPU
# sharding test, the node's concurrent pinned-D2H ceiling.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi topo -m > $OUT/r2e_topo.log 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> $OUT/r2e_topo.log
nproc >> $OUT/r2e_topo.log
echo "== 2-GPU sharding test"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_real_gpus" 2>&1 | tail -3 | tee $OUT/r2e_pytest_2gpu.log
for N in 1 2 4 8; do
  echo "== bench N=$N (driver flags)"
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-config4 2>$OUT/r2e_bench_n$N.err > $OUT/r2e_bench_n$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>$OUT/r2e_bench_n$N.err > $OUT/r2e_bench_n$N.json
  fi
  python - <<PY
import json
d = json.loads(open("$OUT/r2e_bench_n$N.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N=%d value %.4g  us/step %.3f  frac %.3f  checksum %s  e2e %.4g (%.3f ms)  e2e_compact %.4g  weak %.4g  fused %.4g  issue: %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["state_checksum"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e_compact"]["value"], w.get("value", 0), (d.get("fused") or {}).get("value", 0), d["timing"]["issue"][:40]))
PY
done 2>&1 | tee $OUT/r2e_scaling.log
echo "== concurrent D2H ceiling, 8 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/pcie_concurrent.py 2>&1 | grep -v -i warn | tee $OUT/r2e_pcie8.log
echo "== concurrent D2H ceiling, 4 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_concurrent.py 2>&1 | grep -v -i warn | tee $OUT/r2e_pcie4.log
