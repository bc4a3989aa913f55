#!/bin/bash
# perfect-lookahead replay of the Euler walk on a resident chr1 graph (host-side diagnostic; see walk_replay in host_tail.cpp)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2replay}
TAIL_AB_ONLY=default,replay timeout 900 python scripts/tail_ab.py chr1 1.0 3 > gpurun_out/${T}_chr1.txt 2>&1; echo "rc=$?"
grep -v "^\[mtg trace\]" gpurun_out/${T}_chr1.txt | tail -60
