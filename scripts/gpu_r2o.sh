#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2o}
MTG_TRACE=1 MTG_TRACE_ALL=1 TAIL_AB_ONLY=default timeout 600 python scripts/tail_ab.py chr1 1.0 4 2>&1 | grep "tail records\|default" | tail -12
OMP_WAIT_POLICY=active MTG_TRACE=1 MTG_TRACE_ALL=1 TAIL_AB_ONLY=default timeout 600 python scripts/tail_ab.py chr1 1.0 4 2>&1 | grep "tail records\|default" | tail -12
