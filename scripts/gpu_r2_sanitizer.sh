#!/bin/bash
# compute-sanitizer over the smoke run (memcheck) and a slice of the GPU parity tests (racecheck: the shared-memory kernels)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r2_sanitizer_memcheck.log; tail -2 gpurun_out/r2_sanitizer_memcheck.out
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "degenerate or tier0_unpacked or parser_equals_host or overflow_tier" > gpurun_out/r2_sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2_sanitizer_racecheck.log; tail -3 gpurun_out/r2_sanitizer_racecheck.out
