#!/bin/bash
# round 2: ncu --set full captures of the main kernels on the default workload (chr1 x 1.0), one kernel per run
# usage: gpu_r2_profile.sh [kernel-regex:skip ...]   (skip = launches of that kernel to let pass first)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import bench; bench.make_workload('chr1', None)" > /dev/null 2>&1   # build + cache the workload once
LIST=${@:-"dijkstra_thread_kernel:6 match_dataflow_kernel:3 chunk_scatter:3 chunk_scan_text:3 fill_text:6 radix_scatter:40 slot_hints:9 place_slots:3"}
for item in $LIST; do
  k=${item%%:*}; s=${item##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/r2_ncu_$k \
      python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -12
