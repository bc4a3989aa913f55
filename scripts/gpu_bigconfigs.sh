cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --workload pangenome --scale 1.0 > gpurun_out/bench4_pangenome_full.json 2> gpurun_out/bench4_pangenome_full.err; tail -2 gpurun_out/bench4_pangenome_full.err
timeout 1500 python bench.py --steps 3 --warmup 3 --workload chr1 --scale 1.0 > gpurun_out/bench4_chr1_full.json 2> gpurun_out/bench4_chr1_full.err; tail -2 gpurun_out/bench4_chr1_full.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload human --scale 0.05 > gpurun_out/bench4_human_0.05.json 2> gpurun_out/bench4_human_0.05.err; tail -2 gpurun_out/bench4_human_0.05.err
timeout 600 python bench.py > gpurun_out/bench4_ecoli.json 2>/dev/null
