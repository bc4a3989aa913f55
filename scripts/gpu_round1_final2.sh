# Round-1 final measurement pass (1 GPU): tests, smoke, bench lines (both arms) for every workload.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench8_ecoli.json 2> gpurun_out/bench8_ecoli.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench8_ecoli_reference.json 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 --workload pangenome --scale 1.0 > gpurun_out/bench8_pangenome_full.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 --workload pangenome > gpurun_out/bench8_pangenome.json 2>/dev/null
timeout 1200 python bench.py --steps 5 --warmup 3 --workload chr1 --scale 1.0 > gpurun_out/bench8_chr1_full.json 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 --workload human --scale 0.05 > gpurun_out/bench8_human_0.05.json 2>/dev/null
for f in ecoli pangenome_full pangenome chr1_full human_0.05; do python - <<PY
import json
d=json.load(open("gpurun_out/bench8_$f.json"))
print("$f", d["config"]["unitigs"], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["byte_identical_to_oracle"], round(d["cpu_baseline"]["seconds"],3), round(d["roofline"]["frac"],4), d["dijkstra"]["kernel_ms_per_step"], d["dijkstra"]["settled_nodes"])
PY
done
python -c "
import json; d=json.load(open('gpurun_out/bench8_ecoli_reference.json')); print('reference arm', d['ms_per_step'], d['cpu_baseline'])"
