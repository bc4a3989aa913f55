cd $GRAFT_REPO_ROOT
for i in 1 2 3; do python bench.py --steps 10 --warmup 3 --workload chr1 --scale 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); dj=d['dijkstra']
print('chr1 0.3', d['byte_identical_to_oracle'], 'kernel_ms', round(dj['kernel_ms_per_step'],4), 'step_ms', round(dj['ms_per_step'],4), 'settled/s', round(dj['settled_nodes']/dj['kernel_ms_per_step']/1e6,2), 'G/s', 'ovf', dj['overflow_sources'])"; done
python bench.py --steps 50 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); dj=d['dijkstra']
print('ecoli', d['byte_identical_to_oracle'], 'kernel_ms', round(dj['kernel_ms_per_step'],4), d['ms_per_step'], d['e2e']['ms_per_step'])"
