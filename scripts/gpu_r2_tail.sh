#!/bin/bash
# round 2: GPU parity tests, host lookahead micro-benchmark, tail A/B on chr1 x 1.0 (walk depth 4 vs 5), default bench, launch lists
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2c_tests.log
g++ -O2 -o /tmp/host_lookahead scripts/micro/host_lookahead.cpp && /tmp/host_lookahead 512 > gpurun_out/r2c_host_lookahead.jsonl; cat gpurun_out/r2c_host_lookahead.jsonl
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2c_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2c_tail_ab_chr1.txt
for d in 5; do MTG_LIB_PATH=build_variants/d$d.so timeout 600 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2c_tail_ab_chr1_d$d.txt 2>&1; echo "ab d$d rc=$?"; tail -6 gpurun_out/r2c_tail_ab_chr1_d$d.txt; done
timeout 1500 python bench.py --steps 8 --warmup 3 > gpurun_out/r2c_chr1.json 2> gpurun_out/r2c_chr1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c_chr1.err
MTG_NO_PREEXTEND=1 timeout 600 python bench.py --profile --steps 5 --warmup 3 > gpurun_out/r2c_chr1_nopreextend.json 2>&1; echo "nopre rc=$?"; tail -c 700 gpurun_out/r2c_chr1_nopreextend.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_chr1.csv python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload pangenome --steps 8 --warmup 3 > gpurun_out/r2c_pangenome.json 2> gpurun_out/r2c_pangenome.err; echo "bench pan rc=$?"
timeout 600 python bench.py --workload ecoli --steps 20 --warmup 5 > gpurun_out/r2c_ecoli.json 2> gpurun_out/r2c_ecoli.err; echo "bench ecoli rc=$?"
python - <<'PY'
import json
for f in ["r2c_chr1", "r2c_pangenome", "r2c_ecoli"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["e2e"]["ms_per_step"], 2), d["byte_identical_to_oracle"], d["tail_ms_rank0"], {k: round(v, 2) for k, v in d["phases_ms_rank0"].items()}, d["ms_per_step_spread_rank0"])
    except Exception as e:
        print(f, "no line", e)
PY
