"""Writes tests/golden/hand_derived.json.

The reference ships no golden vectors for this path and cannot be run here, so these are NOT reference
outputs.  They are tiny inputs whose expected results were derived by hand from the reference's rules
(see README.md in this directory for two fully written-out derivations); the script only records the
expected values typed in below -- it never asks the oracle.  tests/test_golden.py then checks the oracle
(CPU) and the CUDA path (GPU) against them.
"""
import json
from pathlib import Path

CASES = [
    {"name": "chain_k5", "k": 5, "mode": "fasta", "input": ">0\nAACCGGA\n>1\nCGGATT\n",
     "nodes": 6, "sources": [1, 4], "triples": [], "walks": [[0, 2]],
     "gfa": "H\tKL:Z:5\nS\t1\tAACCGGATT\n", "fasta": ">1\nAACCGGATT\n", "bitvector": "11111\n"},
    {"name": "palindromic_ends_k5", "k": 5, "mode": "fasta", "input": ">0\nACGTTGCA\n",
     "nodes": 2, "sources": [0, 1], "triples": [0, 1, 4], "walks": [[0]],
     "gfa": "H\tKL:Z:5\nS\t1\tACGTTGCA\n", "fasta": ">1\nACGTTGCA\n", "bitvector": "1111\n"},
    {"name": "branch_newest_edge_first_k5", "k": 5, "mode": "fasta",
     "input": ">0\nAAAACCCC\n>1\nCCCCG\n>2\nCCCCT\n>3\nCCCGAAT\n",
     "nodes": 10, "sources": [1, 3, 6, 8], "triples": [], "walks": [[2, 6], [0, 4]],
     "gfa": "H\tKL:Z:5\nS\t1\tCCCCGAAT\nS\t2\tAAAACCCCT\n", "fasta": ">1\nCCCCGAAT\n>2\nAAAACCCCT\n",
     "bitvector": "1111\n11111\n"},
    {"name": "snp_bubble_k5", "k": 5, "mode": "fasta", "input": ">0\nGGACT\n>1\nGACTAGCTT\n>2\nGACTCGCTT\n>3\nGCTTA\n",
     "nodes": 8, "sources": [1, 3, 4, 6], "triples": [], "walks": [[2, 6], [0, 4]],
     "gfa": "H\tKL:Z:5\nS\t1\tGACTAGCTTA\nS\t2\tGGACTCGCTT\n", "fasta": ">1\nGACTAGCTTA\n>2\nGGACTCGCTT\n",
     "bitvector": "111111\n111111\n"},
]

if __name__ == "__main__":
    Path(__file__).with_name("hand_derived.json").write_text(json.dumps(CASES, indent=1) + "\n")
