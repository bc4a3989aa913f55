#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAIL_AB_ONLY=probe timeout 600 python scripts/tail_ab.py chr1 1.0 2 2>&1 | grep "probe" | tail -6 | tee gpurun_out/r2o_probe_chr1.txt
TAIL_AB_ONLY=probe timeout 600 python scripts/tail_ab.py pangenome 1.0 2 2>&1 | grep "probe" | tail -6 | tee gpurun_out/r2o_probe_pan.txt
