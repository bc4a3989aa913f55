#!/bin/bash
# depth-5 (128-byte) walk records: the walk itself next to the perfect-lookahead replay on the same records
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2replay5}
export MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/d5.so
TAIL_AB_REPLAY="1:4,1:5,1:6,1:8,3:4,3:5,3:6,3:8,1:5,3:5" MTG_TRACE=1 TAIL_AB_ONLY=default,replay timeout 900 python scripts/tail_ab.py chr1 1.0 3 > gpurun_out/${T}_chr1.txt 2>&1; echo "rc=$?"
grep -v "trace\] [a-z_]* *[0-9.]* ms" gpurun_out/${T}_chr1.txt | tail -40
