#!/bin/bash
# Box topology + effect of NUMA-local CPU binding on the e2e (host-buffer) leg.
mkdir -p gpurun_out
{
nvidia-smi topo -m
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
  b=$(echo ${d#0000} | tr 'A-Z' 'a-z'); b="0000${b#0000}"
  p=/sys/bus/pci/devices/${b}
  [ -d $p ] || p=/sys/bus/pci/devices/$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//')
  echo "$d numa_node=$(cat $p/numa_node 2>/dev/null) local_cpulist=$(cat $p/local_cpulist 2>/dev/null)"
done
free -g | head -2
} 2>&1 | tee gpurun_out/topo.log
echo "== e2e, default affinity"
python bench.py --steps 2000 --warmup 50 --e2e-steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('e2e %.4g steps/s  %.3f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))" | tee -a gpurun_out/topo.log
