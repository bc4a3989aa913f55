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


# ---- outputs of the real reference binary, if somebody generated them (tests/golden/make_reference_goldens.sh) ----
REF_DIR = Path(__file__).parent / "golden" / "reference"
REF_CASES = json.loads((REF_DIR / "manifest.json").read_text())


def _reference_outputs(case):
    paths = {kind: REF_DIR / case[kind] for kind in ("gfa", "fasta", "bitvector")}
    if not all(p.exists() for p in paths.values()):
        pytest.skip("no output of the real matchtigs binary for this case (needs a Rust toolchain: "
                    "tests/golden/make_reference_goldens.sh); parity with the reference stays unpinned")
    return {kind: p.read_bytes() for kind, p in paths.items()}


def test_reference_inputs_are_reproducible():
    """The committed inputs are exactly what tests/golden/export_inputs.py generates (so the goldens can be regenerated)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_inputs", REF_DIR.parent / "export_inputs.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    names = set()
    for name, text, k, modes in mod.cases():
        assert (REF_DIR / "inputs" / f"{name}.fa").read_bytes() == text, name
        names |= {f"{name}.{m}" for m in modes}
    assert names == {c["name"] for c in REF_CASES}


@pytest.mark.parametrize("case", REF_CASES, ids=[c["name"] for c in REF_CASES])
def test_oracle_matches_reference_binary(case):
    want = _reference_outputs(case)
    o = oracle.Oracle()
    text = (REF_DIR / case["input"]).read_bytes()
    (o.load_fasta if case["mode"] == "fasta" else o.load_bcalm)(text, case["k"])
    o.run()
    for kind in ("gfa", "fasta", "bitvector"):
        assert o.text(kind) == want[kind], kind


@pytest.mark.gpu
@pytest.mark.parametrize("case", REF_CASES, ids=[c["name"] for c in REF_CASES])
def test_cuda_matches_reference_binary(case):
    want = _reference_outputs(case)
    import matchtigs_b200 as mt
    ctx = mt.Context(0)
    try:
        reader = mt.read_bigraph_from_fasta_as_edge_centric if case["mode"] == "fasta" else mt.read_bigraph_from_bcalm2_as_edge_centric
        g = reader((REF_DIR / case["input"]).read_bytes(), case["k"], ctx)
        mt.GreedytigAlgorithm.compute_tigs(g, mt.GreedytigAlgorithmConfiguration(k=case["k"]))
        assert mt.write_walks_gfa(g) == want["gfa"]
        assert mt.write_walks_fasta(g) == want["fasta"]
        assert mt.write_duplication_bitvector(g) == want["bitvector"]
    finally:
        ctx.close()
