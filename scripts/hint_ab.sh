cd $GRAFT_REPO_ROOT
for rep in 1 2; do for mode in "MTG_TAIL_DEEP=0" "MTG_TAIL_DEEP=1" "MTG_TAIL_DEEP=2"; do
  for w in "chr1 0.3" "pangenome 1.0"; do echo "== $mode $w"; env $mode python scripts/phase_times.py $w 16 2>&1 | grep tail_ms | cut -c1-160; done
done; done
