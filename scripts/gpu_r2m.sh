#!/bin/bash
# A/B: record copy chunk (tail_ab), matching spin pause (variants)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2m}
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/${T}_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; grep -v "mtg " gpurun_out/${T}_tail_ab_chr1.txt | tail -6
for v in main sleep0 sleep30; do
  if [ $v = main ]; then unset MTG_LIB_PATH; else export MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/$v.so; fi
  timeout 600 python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dj=d['dijkstra']
print('$v', d['byte_identical_to_oracle'], 'dj kernel_ms', round(dj['kernel_ms_per_step'],4), 'match_ms', round(dj['match_ms_per_step'],3), 'match_kernel_ms', round(dj['match_kernel_ms_per_step'],3), 'retries', dj['match_blocked_retries'], 'step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
unset MTG_LIB_PATH
