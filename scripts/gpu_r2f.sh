#!/bin/bash
# round 2 (session 2): GPU tests, tail A/B on chr1 (hint sources, cache hints), default bench with slow-call tracing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2f}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/${T}_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/${T}_tail_ab_chr1.txt
MTG_TRACE=1 timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_chr1.json 2> gpurun_out/${T}_chr1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${T}_chr1.err
python - <<PY
import json
for f in ["${T}_chr1"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["e2e"]["ms_per_step"], 2), d["byte_identical_to_oracle"], d["tail_ms_rank0"], {k: round(v, 2) for k, v in d["phases_ms_rank0"].items()}, d["ms_per_step_spread_rank0"], d["e2e"]["ms_per_step_spread_rank0"])
    except Exception as e:
        print(f, "no line", e)
PY
