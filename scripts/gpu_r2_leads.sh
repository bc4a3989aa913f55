#!/bin/bash
# how many steps ahead the walk's records are asked for (MTG_TRACE histogram), depth 4 and the depth-5 variant
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2leads}
TAIL_AB_ONLY=default,carried MTG_TRACE=1 timeout 900 python scripts/tail_ab.py chr1 1.0 1 > gpurun_out/${T}_chr1.txt 2>&1; echo "rc=$?"
grep "walk:\|^default\|^carried" gpurun_out/${T}_chr1.txt | tail -12
MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/d5.so TAIL_AB_ONLY=default,carried MTG_TRACE=1 timeout 900 python scripts/tail_ab.py chr1 1.0 1 > gpurun_out/${T}_chr1_d5.txt 2>&1; echo "rc=$?"
grep "walk:\|^default\|^carried" gpurun_out/${T}_chr1_d5.txt | tail -12
