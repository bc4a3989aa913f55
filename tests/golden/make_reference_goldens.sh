#!/usr/bin/env bash
# Produces golden outputs with the REAL reference (algbio/matchtigs 2.1.9, `--threads 1`) for the inputs under
# tests/golden/reference/inputs.  Needs a Rust toolchain and network access -- neither exists in the build environment
# of this repository, which is why parity with the Rust binary is still unpinned.
#
#   cargo install matchtigs --version 2.1.9          # or: MATCHTIGS=/path/to/matchtigs
#   bash tests/golden/make_reference_goldens.sh
#   git add tests/golden/reference/outputs && git commit
#
# tests/test_golden.py picks the outputs up automatically: the CPU oracle (`-m "not gpu"`) and the CUDA path (`-m gpu`)
# must then reproduce them byte for byte.
set -euo pipefail
cd "$(dirname "$0")/reference"
BIN="${MATCHTIGS:-matchtigs}"
mkdir -p outputs
python3 - <<'PY' > /tmp/mtg_golden_jobs.txt
import json
for c in json.load(open("manifest.json")):
    flag = "--fa-in" if c["mode"] == "fasta" else "--bcalm-in"
    print(flag, c["input"], c["k"], c["gfa"], c["fasta"], c["bitvector"])
PY
while read -r flag input k gfa fasta bitvector; do
    echo "== $input ($flag, k=$k)"
    "$BIN" "$flag" "$input" -k "$k" --threads 1 \
        --greedytigs-gfa-out "$gfa" --greedytigs-fa-out "$fasta" --greedytigs-duplication-bitvector-out "$bitvector"
done < /tmp/mtg_golden_jobs.txt
"$BIN" --version > outputs/VERSION.txt 2>&1 || true
echo "done: $(ls outputs | wc -l) files in $(pwd)/outputs"
