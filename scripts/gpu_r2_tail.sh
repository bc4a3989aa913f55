#!/bin/bash
# round 2: GPU parity tests, tail A/B on chr1 x 1.0 (walk depth 4 vs 3, with / without the arena copy and the lookahead),
# default bench, launch list of the same command, other workloads, reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L2|L3|MHz" > gpurun_out/r2_box.txt
echo "nproc=$(nproc)" >> gpurun_out/r2_box.txt
free -g >> gpurun_out/r2_box.txt
nvidia-smi -L >> gpurun_out/r2_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tail_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_tail_tests.log
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2_tail_ab_chr1.txt
MTG_LIB_PATH=build_variants/d3.so timeout 600 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/r2_tail_ab_chr1_d3.txt 2>&1; echo "ab d3 rc=$?"; tail -6 gpurun_out/r2_tail_ab_chr1_d3.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_tail_chr1.json 2> gpurun_out/r2_tail_chr1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_tail_chr1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_tail_launches_chr1.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2_tail_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python scripts/tail_ab.py pangenome 1.0 5 > gpurun_out/r2_tail_ab_pan.txt 2>&1; echo "ab pan rc=$?"; tail -6 gpurun_out/r2_tail_ab_pan.txt
timeout 600 python bench.py --workload pangenome --steps 5 --warmup 3 > gpurun_out/r2_tail_pangenome.json 2> gpurun_out/r2_tail_pangenome.err; echo "bench pan rc=$?"
timeout 600 python bench.py --workload ecoli --steps 20 --warmup 5 > gpurun_out/r2_tail_ecoli.json 2> gpurun_out/r2_tail_ecoli.err; echo "bench ecoli rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_tail_chr1_ref.json 2> gpurun_out/r2_tail_chr1_ref.err; echo "ref rc=$?"
