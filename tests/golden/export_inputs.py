"""Writes the deterministic inputs that `make_reference_goldens.sh` feeds to the real matchtigs binary.

Run here (no Rust needed):  python tests/golden/export_inputs.py
It (re)creates tests/golden/reference/inputs/*.fa and manifest.json.  The inputs are small (a few hundred KB in total)
and are committed, so that whoever has a Rust toolchain needs nothing from this repo but the shell script.
"""
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import tools  # noqa: E402
from helpers import random_fasta  # noqa: E402

OUT = Path(__file__).resolve().parent / "reference"


def cases():
    # compacted de Bruijn graphs (valid for --fa-in and, with their L: links, for --bcalm-in)
    g = tools.genome(30_000, 11, families=6, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=5)
    yield "dbg_repeats_k21", tools.unitigs(g, 21)[0], 21, ["fasta", "bcalm"]
    anc = tools.genome(20_000, 51, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    strains = tools.pangenome(anc, 20, 7, snp_site_rate=0.04, indel_site_rate=0.003)
    yield "dbg_pangenome_k15", tools.unitigs(strains, 15)[0], 15, ["fasta", "bcalm"]
    g = tools.genome(4_000, 5, families=4, copies=8, min_len=9, max_len=60, divergence=0.0, tandem_arrays=3)
    yield "dbg_palindromes_k9", tools.unitigs(g, 9)[0], 9, ["fasta", "bcalm"]
    yield "ecoli_2pct_k31", tools.config_unitigs("ecoli", 0.02)[0], 31, ["fasta", "bcalm"]
    # arbitrary (not dBG-valid) FASTA: parallel edges, self loops, palindromic ends -- stresses every tie-break
    for seed, k, n, pool in [(1, 5, 120, 6), (2, 8, 300, 12), (3, 11, 200, 30), (4, 4, 80, 3)]:
        yield f"arbitrary_s{seed}_k{k}", random_fasta(random.Random(7700 + seed), n, k, max_extra=8, pool=pool), k, ["fasta"]


def main():
    inputs = OUT / "inputs"
    inputs.mkdir(parents=True, exist_ok=True)
    manifest = []
    for name, text, k, modes in cases():
        (inputs / f"{name}.fa").write_bytes(text)
        for mode in modes:
            manifest.append({"name": f"{name}.{mode}", "input": f"inputs/{name}.fa", "k": k, "mode": mode,
                             "gfa": f"outputs/{name}.{mode}.gfa", "fasta": f"outputs/{name}.{mode}.fa",
                             "bitvector": f"outputs/{name}.{mode}.bitvector"})
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(f"{len(manifest)} cases, inputs in {inputs}")


if __name__ == "__main__":
    main()
