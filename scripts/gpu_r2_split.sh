#!/bin/bash
# where the time of the tail's walk phase goes: set-up, run loops, start scan + re-roots, breaking (MTG_TRACE)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2split}
for w in chr1 pangenome; do
TAIL_AB_ONLY=default MTG_TRACE=1 timeout 900 python scripts/tail_ab.py $w 1.0 3 > gpurun_out/${T}_$w.txt 2>&1; echo "rc=$?"
grep "walk:\|^default" gpurun_out/${T}_$w.txt | tail -12
done
