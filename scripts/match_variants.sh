cd $GRAFT_REPO_ROOT
for v in "$@"; do
for w in "ecoli 1.0" "pangenome 0.25" "chr1 0.3"; do set -- $w
MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/$v.so python bench.py --steps 20 --warmup 3 --workload $1 --scale $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); dj=d['dijkstra']
print('$v', '$1', d['byte_identical_to_oracle'], 'match_kernel_ms', round(dj['match_kernel_ms_per_step'],4), 'match_ms', round(dj['match_ms_per_step'],4), 'retries', dj['match_blocked_retries'], 'step', round(d['ms_per_step'],3))"
done; done
