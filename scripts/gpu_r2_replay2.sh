#!/bin/bash
# replay with the walk's own step-to-step dependence (mode 16) and with the prefetch first (mode 32)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2replay2}
TAIL_AB_REPLAY="1:4,17:4,49:4,19:4,51:4,21:4,29:4,61:4,17:5,49:5,17:8,49:8,1:4,17:4,49:4" TAIL_AB_ONLY=late,replay timeout 900 python scripts/tail_ab.py chr1 1.0 1 > gpurun_out/${T}_chr1.txt 2>&1; echo "rc=$?"
grep -v "mtg trace" gpurun_out/${T}_chr1.txt | tail -20
