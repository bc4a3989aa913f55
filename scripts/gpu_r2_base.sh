#!/bin/bash
# round 2 baseline: tests, default bench (chr1 x 1.0, --bcalm-in), launch list of the same command, box facts
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|L2|L3|MHz" > gpurun_out/r2_box.txt
echo "nproc=$(nproc)" >> gpurun_out/r2_box.txt
free -g >> gpurun_out/r2_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_base_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/r2_base_tests.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_base_chr1.json 2> gpurun_out/r2_base_chr1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_base_chr1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_base_launches_chr1.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2_base_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_base_chr1_ref.json 2> gpurun_out/r2_base_chr1_ref.err; echo "ref rc=$?"
