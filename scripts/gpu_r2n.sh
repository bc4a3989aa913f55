#!/bin/bash
# walk records of depth 5 (128 B) against depth 4 (64 B) with the round-2 run loop
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2n}
for v in main d5 main d5; do
  if [ $v = main ]; then unset MTG_LIB_PATH; else export MTG_LIB_PATH=$GRAFT_REPO_ROOT/build_variants/$v.so; fi
  echo "== $v" >> gpurun_out/${T}_tail_depth.txt
  TAIL_AB_ONLY=default,sources1 timeout 600 python scripts/tail_ab.py chr1 1.0 5 2>&1 | grep -v "mtg \|THP\|unitigs" >> gpurun_out/${T}_tail_depth.txt
done
cat gpurun_out/${T}_tail_depth.txt
