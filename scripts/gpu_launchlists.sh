cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for W in ecoli pangenome; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches2_$W.csv python bench.py --steps 2 --warmup 3 --workload $W > gpurun_out/ncu_launches2_$W.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench2_ecoli.json 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --workload pangenome > gpurun_out/bench2_pangenome.json 2>/dev/null
