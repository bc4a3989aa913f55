#!/bin/bash
# round 2 measurement pass on one GPU: GPU tests, smoke(), the bench lines of every workload (both arms for the default one),
# the launch list of the default bench command.  Outputs: gpurun_out/${T}_*
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2final}
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L2|L3|MHz" > gpurun_out/${T}_box.txt; echo "nproc=$(nproc)" >> gpurun_out/${T}_box.txt; free -g >> gpurun_out/${T}_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 1500 python bench.py > gpurun_out/${T}_chr1.json 2> gpurun_out/${T}_chr1.err; echo "bench chr1 rc=$?"; tail -c 400 gpurun_out/${T}_chr1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_chr1_reference.json 2> gpurun_out/${T}_chr1_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --workload pangenome > gpurun_out/${T}_pangenome.json 2> gpurun_out/${T}_pangenome.err; echo "bench pangenome rc=$?"
timeout 900 python bench.py --workload ecoli --steps 25 > gpurun_out/${T}_ecoli.json 2> gpurun_out/${T}_ecoli.err; echo "bench ecoli rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_chr1.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ["${T}_chr1", "${T}_pangenome", "${T}_ecoli", "${T}_chr1_reference"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"].get("ms_per_step", 0), 2), d.get("byte_identical_to_oracle"),
              {k: round(v, 2) for k, v in d.get("phases_ms_rank0", {}).items()}, d.get("tail_ms_rank0"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "no line", e)
PY
