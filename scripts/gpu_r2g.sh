#!/bin/bash
# tail A/B on chr1 (+ latency probes)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2g}
lscpu | grep -i "model name\|mhz\|L2\|L3" > gpurun_out/${T}_cpu.txt
timeout 900 python scripts/tail_ab.py chr1 1.0 5 > gpurun_out/${T}_tail_ab_chr1.txt 2>&1; echo "ab rc=$?"; tail -22 gpurun_out/${T}_tail_ab_chr1.txt
