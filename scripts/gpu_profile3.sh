cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/phase_times.py human 0.01 16 2>&1 | tail -3 | cut -c1-600
python scripts/phase_times.py chr1 0.3 16 2>&1 | tail -3 | cut -c1-900
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_chr1.csv python bench.py --steps 1 --warmup 3 --workload chr1 --scale 0.3 > gpurun_out/ncu_launches_chr1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dijkstra_warp_kernel -s 6 -c 1 -o gpurun_out/prof_dijkstra_chr1 -f python bench.py --steps 1 --warmup 3 --workload chr1 --scale 0.3 > gpurun_out/ncu_full_dj2.log 2>&1
tail -3 gpurun_out/ncu_full_dj2.log | cut -c1-300
