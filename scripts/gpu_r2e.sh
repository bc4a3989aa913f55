#!/bin/bash
# round 2 (session 2): tail A/B on chr1, then the default bench with slow-call tracing, with and without the clock sampler
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2e_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -7 gpurun_out/r2e_tail_ab_chr1.txt
MTG_TRACE=1 timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_chr1.json 2> gpurun_out/r2e_chr1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2e_chr1.err
MTG_TRACE=1 MTG_BENCH_NOSAMPLER=1 timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_chr1_nosampler.json 2> gpurun_out/r2e_chr1_nosampler.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2e_chr1_nosampler.err
python - <<'PY'
import json
for f in ["r2e_chr1", "r2e_chr1_nosampler"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["e2e"]["ms_per_step"], 2), d["byte_identical_to_oracle"], d["tail_ms_rank0"], {k: round(v, 2) for k, v in d["phases_ms_rank0"].items()}, d["ms_per_step_spread_rank0"])
    except Exception as e:
        print(f, "no line", e)
PY
