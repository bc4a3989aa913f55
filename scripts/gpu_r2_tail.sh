#!/bin/bash
# round 2: GPU parity tests, tail A/B on chr1 x 1.0 and pangenome, default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2d_tests.log
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2d_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -9 gpurun_out/r2d_tail_ab_chr1.txt
timeout 900 python scripts/tail_ab.py pangenome 1.0 5 > gpurun_out/r2d_tail_ab_pan.txt 2>&1; echo "ab rc=$?"; tail -9 gpurun_out/r2d_tail_ab_pan.txt
timeout 1500 python bench.py --steps 8 --warmup 3 > gpurun_out/r2d_chr1.json 2> gpurun_out/r2d_chr1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_chr1.err
python - <<'PY'
import json
for f in ["r2d_chr1"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["e2e"]["ms_per_step"], 2), d["byte_identical_to_oracle"], d["tail_ms_rank0"], {k: round(v, 2) for k, v in d["phases_ms_rank0"].items()}, d["ms_per_step_spread_rank0"])
    except Exception as e:
        print(f, "no line", e)
PY
