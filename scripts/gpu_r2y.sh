#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --workload ecoli --steps 25 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ecoli', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['byte_identical_to_oracle'], d['tail_ms_rank0'])"
timeout 900 python bench.py --workload pangenome --scale 0.25 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pan0.25', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['byte_identical_to_oracle'], d['tail_ms_rank0'], d['workload_sizes']['unitigs'])"
