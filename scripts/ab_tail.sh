cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for mode in "" "MTG_TAIL_COLDCOPY=1" "MTG_TAIL_HOST=1" "MTG_TAIL_HOST=1 OMP_NUM_THREADS=1"; do
  echo "== $mode"; env $mode python scripts/phase_times.py pangenome 1.0 16 2>&1 | tail -1 | cut -c1-140
done; done
grep AnonHuge /proc/meminfo
