# One-GPU measurement pass: bench lines (no profiler), then ncu launch list and full captures of the top kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
W=${1:-pangenome}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ecoli.json 2> gpurun_out/bench_ecoli.err; tail -2 gpurun_out/bench_ecoli.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload $W > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -2 gpurun_out/bench_$W.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --workload $W > gpurun_out/bench_${W}_reference.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ecoli_reference.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$W.csv python bench.py --steps 2 --warmup 3 --workload $W > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dijkstra_warp_kernel|match_rounds_kernel|radix_scatter' -c 6 -o gpurun_out/prof_$W -f python bench.py --steps 1 --warmup 3 --workload $W > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
nproc; lscpu | grep -E "Model name|^CPU\(s\)|L2|L3" 
