#!/bin/bash
# what the box exposes about GPU <-> NUMA locality (for bench.py's per-rank binding)
nvidia-smi topo -m 2>&1 | head -30
echo ---
ls /sys/devices/system/node/ 2>&1 | head
for d in /sys/devices/system/node/node*; do echo $d $(cat $d/cpulist 2>/dev/null); done
echo ---
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader
for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo $b $(cat /sys/bus/pci/devices/$b/numa_node 2>&1) $(cat /sys/bus/pci/devices/$b/local_cpulist 2>&1); done
echo ---
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"
nproc; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"
numactl --hardware 2>&1 | head -5
