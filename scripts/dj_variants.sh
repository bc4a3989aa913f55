cd $GRAFT_REPO_ROOT
for v in "$@"; do
for w in "chr1 0.3" "pangenome 0.25"; do set -- $w
MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/$v.so python bench.py --steps 10 --warmup 3 --workload $1 --scale $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); dj=d['dijkstra']
print('$v', '$1', d['byte_identical_to_oracle'], 'kernel_ms', round(dj['kernel_ms_per_step'],4), 'dj_step_ms', round(dj['ms_per_step'],4), 'ovf', dj['overflow_sources'], 'step', round(d['ms_per_step'],2))"
done; done
