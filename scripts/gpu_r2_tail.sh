#!/bin/bash
# round 2: GPU parity tests, tail A/B on chr1 x 1.0 (walk depth 5 vs 4 vs 3, with / without the arena copy and the lookahead),
# default bench, launch list of the same command, other workloads, reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2b_tests.log
cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2b_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2b_tail_ab_chr1.txt
for d in 4 3; do MTG_LIB_PATH=build_variants/d$d.so timeout 600 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2b_tail_ab_chr1_d$d.txt 2>&1; echo "ab d$d rc=$?"; tail -6 gpurun_out/r2b_tail_ab_chr1_d$d.txt; done
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_chr1.json 2> gpurun_out/r2b_chr1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_chr1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_chr1.csv python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload pangenome --steps 5 --warmup 3 > gpurun_out/r2b_pangenome.json 2> gpurun_out/r2b_pangenome.err; echo "bench pan rc=$?"
timeout 600 python bench.py --workload ecoli --steps 20 --warmup 5 > gpurun_out/r2b_ecoli.json 2> gpurun_out/r2b_ecoli.err; echo "bench ecoli rc=$?"
MTG_TAIL_HOST=0 timeout 600 python bench.py --workload ecoli --steps 20 --warmup 5 > gpurun_out/r2b_ecoli_devprep.json 2> gpurun_out/r2b_ecoli_devprep.err; echo "bench ecoli devprep rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_ecoli.csv python bench.py --workload ecoli --profile --steps 2 --warmup 3 > gpurun_out/r2b_ncu_ecoli.log 2>&1; echo "ncu ecoli rc=$?"
python - <<'PY'
import json
for f in ["r2b_chr1", "r2b_pangenome", "r2b_ecoli", "r2b_ecoli_devprep"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["e2e"]["ms_per_step"], 2), d["byte_identical_to_oracle"], d["tail_ms_rank0"], {k: round(v, 2) for k, v in d["phases_ms_rank0"].items()})
    except Exception as e:
        print(f, "no line", e)
PY
