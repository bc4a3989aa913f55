cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do for mode in "MTG_TAIL_NOHINT=1" "MTG_X=1"; do
  for w in "chr1 0.3" "pangenome 1.0"; do echo "== $mode $w"; env $mode python scripts/phase_times.py $w 16 2>&1 | grep tail_ms | cut -c1-160; done
done; done
