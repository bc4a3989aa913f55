cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches7_ecoli.csv python bench.py --steps 2 --warmup 3 > gpurun_out/l7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dijkstra_thread_kernel -s 3 -c 1 -o gpurun_out/prof7_dj_thread_ecoli -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu7e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dijkstra_thread_kernel -s 6 -c 1 -o gpurun_out/prof7_dj_thread_chr1 -f python bench.py --steps 1 --warmup 3 --workload chr1 --scale 0.3 > gpurun_out/ncu7c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_dataflow_kernel -s 3 -c 1 -o gpurun_out/prof7_match_ecoli -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu7m.log 2>&1
ls -la gpurun_out/*7*
