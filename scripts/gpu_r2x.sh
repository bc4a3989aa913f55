#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log
MTG_TRACE=1 MTG_TRACE_ALL=1 TAIL_AB_ONLY=default,memcpy timeout 600 python scripts/tail_ab.py chr1 1.0 5 2>&1 | grep "tail records\|default\|memcpy" | tail -24 | tee gpurun_out/r2x_tail_copy.txt
