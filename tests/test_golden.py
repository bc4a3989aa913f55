"""Known-answer tests against tests/golden/hand_derived.json (hand-derived, see tests/golden/README.md):
the oracle on CPU, the CUDA path on GPU."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle

CASES = json.loads((Path(__file__).parent / "golden" / "hand_derived.json").read_text())


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_hand_derived(case):
    o = oracle.Oracle()
    (o.load_fasta if case["mode"] == "fasta" else o.load_bcalm)(case["input"].encode(), case["k"])
    o.run()
    assert o.num("nodes") == case["nodes"]
    assert list(o.array("out_nodes")) == case["sources"]
    assert list(o.array("triples")) == case["triples"]
    assert [list(map(int, w)) for w in o.walks()] == case["walks"]
    assert o.text("gfa").decode() == case["gfa"]
    assert o.text("fasta").decode() == case["fasta"]
    assert o.text("bitvector").decode() == case["bitvector"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_hand_derived(case):
    import matchtigs_b200 as mt
    ctx = mt.Context(0)
    try:
        reader = mt.read_bigraph_from_fasta_as_edge_centric if case["mode"] == "fasta" else mt.read_bigraph_from_bcalm2_as_edge_centric
        g = reader(case["input"].encode(), case["k"], ctx)
        walks = mt.GreedytigAlgorithm.compute_tigs(g, mt.GreedytigAlgorithmConfiguration(k=case["k"]))
        assert ctx.graph_info()["nodes"] == case["nodes"]
        assert list(ctx.graph_export()["sources"]) == case["sources"]
        assert [list(map(int, w)) for w in walks] == case["walks"]
        assert mt.write_walks_gfa(g).decode() == case["gfa"]
        assert mt.write_walks_fasta(g).decode() == case["fasta"]
        assert mt.write_duplication_bitvector(g).decode() == case["bitvector"]
    finally:
        ctx.close()
