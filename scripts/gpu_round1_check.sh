cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ecoli.json 2> gpurun_out/bench_ecoli.err; tail -3 gpurun_out/bench_ecoli.err; cat gpurun_out/bench_ecoli.json
timeout 600 python bench.py --steps 5 --warmup 3 --workload pangenome > gpurun_out/bench_pan.json 2> gpurun_out/bench_pan.err; tail -3 gpurun_out/bench_pan.err; cat gpurun_out/bench_pan.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pan.csv python bench.py --steps 1 --warmup 3 --workload pangenome > gpurun_out/ncu_pan.log 2>&1; tail -2 gpurun_out/ncu_pan.log | cut -c1-300
