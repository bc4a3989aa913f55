cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dijkstra_warp_kernel -s 3 -c 1 -o gpurun_out/prof_dijkstra -f python bench.py --steps 1 --warmup 3 --workload pangenome > gpurun_out/ncu_full_dj.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_rounds_kernel -s 6 -c 1 -o gpurun_out/prof_match -f python bench.py --steps 1 --warmup 3 --workload pangenome > gpurun_out/ncu_full_match.log 2>&1
(time timeout 900 python scripts/phase_times.py pangenome 1.0 16) 2>&1 | tail -9 | cut -c1-1000
(time timeout 1200 python scripts/phase_times.py chr1 1.0 16) 2>&1 | tail -9 | cut -c1-1000
