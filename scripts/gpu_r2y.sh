#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAIL_AB_ONLY=default,sources1,fast0,spin0,spin20,spin40 timeout 900 python scripts/tail_ab.py chr1 1.0 5 2>&1 | grep "default\|spin\|sources\|fast" | tee gpurun_out/r2y2_chr1.txt
TAIL_AB_ONLY=default timeout 900 python scripts/tail_ab.py pangenome 1.0 5 2>&1 | grep "default" | tee gpurun_out/r2y2_pan.txt
