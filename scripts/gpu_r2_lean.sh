#!/bin/bash
# the lean run loop (default) against the round's earlier carried-chain loop (MTG_WALK_CHAIN=carried), interleaved, on resident graphs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r2lean}
for w in chr1 pangenome; do
TAIL_AB_ONLY=default,carried,lean2,carried2,lean3 MTG_TRACE=${MTG_TRACE_AB:-} timeout 900 python scripts/tail_ab.py $w 1.0 5 > gpurun_out/${T}_$w.txt 2>&1; echo "rc=$?"
grep -v "mtg trace\] [a-z_(), +]* *[0-9.]* ms" gpurun_out/${T}_$w.txt | tail -9
done
