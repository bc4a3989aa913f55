#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAIL_AB_ONLY=default,store,nta timeout 900 python scripts/tail_ab.py chr1 1.0 5 2>&1 | grep "default\|store\|nta" | tee gpurun_out/r2y3_chr1.txt
TAIL_AB_ONLY=default,store,nta timeout 900 python scripts/tail_ab.py pangenome 1.0 5 2>&1 | grep "default\|store\|nta" | tee gpurun_out/r2y3_pan.txt
