# Round-1 measurement pass (1 GPU): tests, smoke, bench lines (both arms), launch lists and full ncu captures.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench3_ecoli.json 2> gpurun_out/bench3_ecoli.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench3_ecoli_reference.json 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --workload pangenome > gpurun_out/bench3_pangenome.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --workload pangenome > gpurun_out/bench3_pangenome_reference.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches3_ecoli.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu3_ecoli.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dijkstra_thread_kernel -s 3 -c 1 -o gpurun_out/prof3_dijkstra_thread -f python bench.py --steps 1 --warmup 3 --workload pangenome > gpurun_out/ncu3_full_dj.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_dataflow_kernel -s 3 -c 1 -o gpurun_out/prof3_match_dataflow -f python bench.py --steps 1 --warmup 3 --workload pangenome > gpurun_out/ncu3_full_match.log 2>&1
ls gpurun_out | grep 3
