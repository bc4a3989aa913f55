cd $GRAFT_REPO_ROOT
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L2|L3|MHz" 
echo "nproc=$(nproc) affinity=$(taskset -p $$)"
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag
nvidia-smi topo -m 2>/dev/null | head -8
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemFree $n/meminfo)"; done
python scripts/tail_times.py chr1 0.3 2>&1 | tail -3
