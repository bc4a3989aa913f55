#!/bin/bash
# GPU tests, Dijkstra tier-0 variants on chr1 x 1.0 (label slots per thread), default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2j}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python -c "import bench; bench.make_workload('chr1', None)" > /dev/null 2>&1
for v in main; do
  if [ $v = main ]; then unset MTG_LIB_PATH; else export MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/$v.so; fi
  timeout 600 python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dj=d['dijkstra']
print('$v', d['byte_identical_to_oracle'], 'kernel_ms', round(dj['kernel_ms_per_step'],4), 'dj_step_ms', round(dj['ms_per_step'],4), 'ovf', dj['overflow_sources'], 'match_ms', round(dj['match_ms_per_step'],3), 'step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['tail_ms_rank0'])"
done
unset MTG_LIB_PATH
