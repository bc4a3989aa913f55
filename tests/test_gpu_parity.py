"""GPU parity tests: every stage of the CUDA path against the CPU oracle, through the C ABI.

Run on the B200 box: ``python -m pytest tests -m gpu -x -q``.  Integer / byte work: bit-exact.
"""
import random

import numpy as np
import pytest

import oracle
import tools
from helpers import check_tig_invariants, random_fasta

pytestmark = pytest.mark.gpu

META_COUNT = 0x00FFFFFF
META_TRUNC = 0x80000000


@pytest.fixture(scope="module")
def mt():
    import matchtigs_b200
    return matchtigs_b200


@pytest.fixture(scope="module")
def ctx(mt):
    c = mt.Context(0)
    yield c
    c.close()


def run_oracle(text, k, mode, **opt):
    o = oracle.Oracle(**opt)
    (o.load_fasta if mode == "fasta" else o.load_bcalm)(text, k)
    return o.run()


def build(mt, ctx, text, k, mode, device_parse=True):
    if mode == "fasta":
        return mt.read_bigraph_from_fasta_as_edge_centric(text, k, ctx, device_parse=device_parse)
    return mt.read_bigraph_from_bcalm2_as_edge_centric(text, k, ctx, device_parse=device_parse)


def compare_all(mt, ctx, text, k, mode, cap=8, dbg_valid=False, check_props=False, opt=None):
    """`opt`: parity-assumption switches, applied to the oracle here; the caller has set the same ones on `ctx`."""
    o = run_oracle(text, k, mode, **(opt or {}))
    g = build(mt, ctx, text, k, mode)
    U = o.num("unitigs")
    # --- step 1: graph ---
    gi = ctx.graph_info()
    assert gi["unitigs"] == U
    assert gi["nodes"] == o.num("nodes"), "node count"
    ex = ctx.graph_export()
    assert np.array_equal(ex["edge_from"], o.array("edge_from")[:2 * U]), "edge_from"
    assert np.array_equal(ex["edge_to"], o.array("edge_to")[:2 * U]), "edge_to"
    assert np.array_equal(ex["mirror"], o.array("mirror")), "mirror"
    assert np.array_equal(ex["imbalance"].astype(np.int64), o.array("mult0")), "imbalance"
    assert np.array_equal(ex["sources"], o.array("out_nodes")), "sources"
    assert gi["targets"] == int(o.array("in_node_map0").sum())
    # --- step 2: candidate lists ---
    ctx.dijkstra_candidates(cap)
    nodes, dists, meta = ctx.candidates_export()
    on, od, ol = o.candidates(cap)
    cnt = (meta & META_COUNT).astype(np.int64)
    assert np.array_equal(cnt, np.minimum(ol, cap)), "candidate counts"
    trunc = (meta & META_TRUNC) != 0
    assert not np.any(trunc & (ol <= cap) & (cnt < cap)), "truncated flag on a complete short list"
    assert np.all(trunc[ol > cap]), "missing truncated flag"
    mask = np.arange(cap)[None, :] < cnt[:, None]
    assert np.array_equal(nodes[mask], on[mask]), "candidate nodes"
    assert np.array_equal(dists[mask], od[mask]), "candidate distances"
    # --- step 3: matching ---
    tr = ctx.greedy_match()
    assert np.array_equal(tr.reshape(-1), o.array("triples")), "matched triples"
    # --- tail + outputs ---
    ctx.finish_walks()
    gw, ow = ctx.walks(), o.walks()
    assert len(gw) == len(ow), "walk count"
    for a, b in zip(gw, ow):
        assert np.array_equal(a, b), "walk edges"
    gfa, fa, bv = mt.write_walks_gfa(g), mt.write_walks_fasta(g), mt.write_duplication_bitvector(g)
    assert gfa == o.text("gfa"), "GFA bytes"
    assert fa == o.text("fasta"), "FASTA bytes"
    assert bv == o.text("bitvector"), "bitvector bytes"
    eo, io_, lim = ctx.walks_capi(U)
    assert np.array_equal(eo, o.array("c_edge_out")) and np.array_equal(io_, o.array("c_insert_out"))
    assert np.array_equal(lim, o.array("c_limits"))
    if check_props:
        check_tig_invariants(text, k, gfa, fa, bv, dbg_valid)
    return o, ctx.search_stats()


@pytest.mark.parametrize("mode", ["fasta", "bcalm"])
@pytest.mark.parametrize("seed", range(3))
def test_small_dbg(mt, ctx, mode, seed):
    g = tools.genome(30_000, 10 + seed, families=6, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=5)
    text, _, _ = tools.unitigs(g, 21)
    compare_all(mt, ctx, text, 21, mode, dbg_valid=True, check_props=True)


@pytest.mark.parametrize("mode", ["fasta", "bcalm"])
def test_pangenome_small(mt, ctx, mode):
    anc = tools.genome(20_000, 51, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    strains = tools.pangenome(anc, 20, 7, snp_site_rate=0.04, indel_site_rate=0.003)
    text, _, _ = tools.unitigs(strains, 15)
    o, st = compare_all(mt, ctx, text, 15, mode, dbg_valid=True, check_props=True)
    assert st["matched"] > 0


@pytest.mark.parametrize("k", [2, 3, 4, 5, 8, 11, 16, 31, 32, 33, 34, 40, 51, 63, 64])
def test_arbitrary_fasta_all_k(mt, ctx, k):
    for seed in range(4):
        rng = random.Random(1000 * k + seed)
        text = random_fasta(rng, rng.randint(1, 80), k, max_extra=12, pool=rng.choice([None, 3, 8]) if k > 2 else None)
        compare_all(mt, ctx, text, k, "fasta", check_props=(k <= 11))


@pytest.mark.parametrize("block", range(4))
def test_fuzz_tiny_genomes_both_readers(mt, ctx, block):
    """Many tiny, repeat-rich genomes compacted at small odd k (the unitig builder needs odd k; palindromic (k-1)-mers,
    hairpins and self-loops are frequent at this size), through both reader semantics, every stage compared."""
    for i in range(12):
        rng = random.Random(90_000 + 100 * block + i)
        k = 2 * rng.randint(1, 8) + 1
        g = tools.genome(rng.randint(40, 2500), rng.randint(1, 10**6), families=rng.randint(0, 4), copies=rng.randint(2, 6),
                         min_len=k, max_len=rng.randint(k + 1, 120), divergence=rng.choice([0.0, 0.02, 0.1]),
                         tandem_arrays=rng.randint(0, 3))
        text, _, _ = tools.unitigs(g, k)
        cap = rng.choice([1, 2, 4, 8, 16])
        for mode in ("fasta", "bcalm"):
            compare_all(mt, ctx, text, k, mode, cap=cap, dbg_valid=True, check_props=True)


@pytest.mark.parametrize("block", range(4))
def test_fuzz_dense_arbitrary_fasta(mt, ctx, block):
    """Arbitrary FASTA with record ends drawn from a tiny pool: parallel edges, self-mirrors and long candidate lists."""
    for i in range(12):
        rng = random.Random(95_000 + 100 * block + i)
        k = rng.randint(3, 13)
        text = random_fasta(rng, rng.randint(1, 600), k, max_extra=rng.choice([0, 2, 8, 30]), pool=rng.choice([2, 3, 5, 9, 40]))
        compare_all(mt, ctx, text, k, "fasta", cap=rng.choice([1, 3, 16]))


@pytest.mark.parametrize("cap", [1, 2, 3, 8, 64])
def test_cap_independence(mt, ctx, cap):
    # dense graphs with many short edges: small caps force re-query phases, results must not change
    for seed in range(4):
        rng = random.Random(500 + seed)
        k = rng.choice([7, 9, 11])
        text = random_fasta(rng, rng.randint(100, 400), k, max_extra=6, pool=rng.choice([6, 12, 30]))
        compare_all(mt, ctx, text, k, "fasta", cap=cap)


@pytest.mark.parametrize("cap", [1, 2, 3, 5])
def test_targets_reachable_only_through_targets(mt, ctx, cap):
    """A chain s -> v0 -> v1 -> ... of single-k-mer unitigs where every v_i is a target (extra out-tip) and is the only
    way to v_{i+1}: a list cut at `cap` must be flagged truncated even though no unsettled label is left when the
    cap-th target is popped (regression: the per-thread search tier forgot to relax the last target first)."""
    k = 9
    for seed in range(6):
        rng = random.Random(7000 + seed)
        R = bytes(rng.choice(b"ACGT") for _ in range(k - 1 + 14))
        recs = [R[i:i + k] for i in range(len(R) - k + 1)]                      # the chain, weight 1 each
        for i in range(1, len(R) - k + 2):                                      # out-tips: every chain node becomes a target
            node = R[i:i + k - 1]
            alt = bytes([c for c in b"ACGT" if c != R[i + k - 1]][:1]) if i + k - 1 < len(R) else b"A"
            recs.append(node + alt + bytes(rng.choice(b"ACGT") for _ in range(12)))
        for _ in range(2):                                                      # in-tips: the chain head becomes a source
            recs.append(bytes(rng.choice(b"ACGT") for _ in range(12)) + R[:k - 1])
        text = b"".join(b">%d\n%s\n" % (i, s) for i, s in enumerate(recs))
        compare_all(mt, ctx, text, k, "fasta", cap=cap)


def test_requery_phase_is_exercised(mt, ctx):
    rng = random.Random(4242)
    hit = 0
    for _ in range(6):
        text = random_fasta(rng, 400, 9, max_extra=4, pool=10)
        _, st = compare_all(mt, ctx, text, 9, "fasta", cap=1)
        hit += st["requery_phases"]
    assert hit > 0, "cap=1 never ran out of candidates: the re-query path is untested"


def test_overflow_tier(mt, ctx):
    # single-k-mer unitigs (weight 1) drawn from a dense de Bruijn graph: balls of radius k-1 hold far more
    # than the 320 labelled nodes of the shared-memory table, so tier 2 must produce the same lists
    rng = random.Random(99)
    k = 8
    kmers = set()
    while len(kmers) < 40000:
        kmers.add(bytes(rng.choice(b"ACGT") for _ in range(k)))
    text = b"".join(b">%d\n%s\n" % (i, s) for i, s in enumerate(sorted(kmers)))
    _, st = compare_all(mt, ctx, text, k, "fasta", cap=16)
    assert st["overflow_sources"] > 0, "tier 2 not exercised"


@pytest.mark.parametrize("tie_desc", [0, 1])
def test_tier0_unpacked_labels(mt, ctx, monkeypatch, tie_desc):
    """Graphs with at most 2^26 nodes run the tier-0 search with one-word labels (distance << 26 | node id); bigger ones
    with separate id / distance arrays.  MTG_T0_UNPACKED=1 forces the general kernel on a small input: same candidate
    lists, same everything, under both tie orders of assumption P1."""
    anc = tools.genome(30_000, 17, families=4, copies=4, min_len=40, max_len=400, divergence=0.03)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 12, 5, snp_site_rate=0.05, indel_site_rate=0.004), 19)
    monkeypatch.setenv("MTG_T0_UNPACKED", "1")
    ctx.set_option("p1_tie_desc", tie_desc)
    try:
        compare_all(mt, ctx, text, 19, "fasta", cap=8, opt={"p1_tie_desc": tie_desc} if tie_desc else None)
        compare_all(mt, ctx, text, 19, "bcalm", cap=3, opt={"p1_tie_desc": tie_desc} if tie_desc else None)
    finally:
        ctx.set_option("p1_tie_desc", 0)


def test_empty_and_degenerate_inputs(mt, ctx):
    compare_all(mt, ctx, b"", 5, "fasta")
    compare_all(mt, ctx, b">0\nACGTA\n", 5, "fasta")                # a single k-mer
    compare_all(mt, ctx, b">0\nACGT\n", 5 - 1, "fasta")              # palindromic unitig, even k
    compare_all(mt, ctx, b">0\nAAAAAA\n>1\nAAAAAA\n", 5, "fasta")   # duplicate records, self loops
    compare_all(mt, ctx, b">0 LN:i:6\nACGTTT\n", 5, "bcalm")        # no links at all
    compare_all(mt, ctx, b">0\nACGTA\nCCG\n>1\nCCGTT\n", 4, "fasta")  # multi-line record


def test_input_errors(mt, ctx):
    with pytest.raises(mt.MatchtigsError) as e:
        mt.read_bigraph_from_fasta_as_edge_centric(b">0\nACGNT\n", 3, ctx)
    assert e.value.code == -3
    with pytest.raises(mt.MatchtigsError):
        mt.read_bigraph_from_fasta_as_edge_centric(b">0\nAC\n", 5, ctx)
    with pytest.raises(mt.MatchtigsError):
        mt.read_bigraph_from_bcalm2_as_edge_centric(b">1 LN:i:5\nACGTA\n", 5, ctx)
    with pytest.raises(mt.MatchtigsError):
        mt.read_bigraph_from_bcalm2_as_edge_centric(b">0 LN:i:5 L:+:7:+\nACGTA\n", 5, ctx)
    with pytest.raises(mt.MatchtigsError):
        ctx.build_graph_from_sequences(np.frombuffer(b"ACGT", np.uint8), np.array([0, 4], np.uint64), 99)
    # the context stays usable after errors
    compare_all(mt, ctx, b">0\nACGTA\n", 5, "fasta")


def graph_state(ctx):
    ex = ctx.graph_export()
    return ctx.graph_info(), {k: v.copy() for k, v in ex.items()}


@pytest.mark.parametrize("bcalm", [False, True])
def test_device_parser_equals_host_reader(mt, ctx, bcalm):
    """mtg_build_graph_from_text (records parsed on the device) == host reader + array builders, incl. outputs."""
    anc = tools.genome(20_000, 31, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 8, 4, snp_site_rate=0.04, indel_site_rate=0.003), 21)
    # multi-line records, CRLF line ends, blank lines, no trailing newline
    lines = text.split(b"\n")
    messy = []
    for ln in lines:
        if ln.startswith(b">") or len(ln) < 40:
            messy.append(ln + b"\r")
        else:
            messy.extend([ln[:17], ln[17:33] + b"\r", b"", ln[33:]])
    variants = [text, b"\n".join(messy).rstrip(b"\n")]
    for t in variants:
        o = run_oracle(t, 21, "bcalm" if bcalm else "fasta")
        build(mt, ctx, t, 21, "bcalm" if bcalm else "fasta", device_parse=False)
        gi_ref, ex_ref = graph_state(ctx)
        ctx.build_graph_from_text(t, 21, bcalm)
        gi, ex = graph_state(ctx)
        assert gi == gi_ref
        for name in ex_ref:
            assert np.array_equal(ex[name], ex_ref[name]), name
        ctx.dijkstra_candidates(8)
        ctx.greedy_match()
        ctx.finish_walks()
        assert ctx.assemble_tigs("gfa") == o.text("gfa") and ctx.dup_bitvector() == o.text("bitvector")


@pytest.mark.parametrize("bcalm", [False, True])
def test_device_parser_text_beyond_4_gib(mt, ctx, bcalm):
    """Text positions are 64-bit in the device parser (the 32-bit quantities are chunk tags and guarded sums): a file of
    4.3 GiB -- a small workload whose header lines carry megabytes of comment -- must give the graph and the outputs of
    the same records without the padding (bcalm2 files of human-scale inputs are of this size, src/bin.rs:905-911)."""
    text, k, info = tools.config_unitigs("ecoli", 0.02)
    mode = "bcalm" if bcalm else "fasta"
    o = run_oracle(text, k, mode)
    build(mt, ctx, text, k, mode, device_parse=True)
    gi_ref, ex_ref = graph_state(ctx)
    lines = text.split(b"\n")
    n_hdr = sum(ln.startswith(b">") for ln in lines)
    pad = b" " + b"x" * (int(4.3 * 2**30) // n_hdr)
    big = b"\n".join(ln + pad if ln.startswith(b">") else ln for ln in lines)
    del lines
    assert len(big) > 2**32 + 2**20
    ctx.build_graph_from_text(big, k, bcalm)
    del big
    gi, ex = graph_state(ctx)
    assert gi == gi_ref
    for name in ex_ref:
        assert np.array_equal(ex[name], ex_ref[name]), name
    ctx.dijkstra_candidates(8)
    ctx.greedy_match()
    ctx.finish_walks()
    assert ctx.assemble_tigs("gfa") == o.text("gfa") and ctx.dup_bitvector() == o.text("bitvector")


def test_device_parser_errors_and_edge_cases(mt, ctx):
    for bad, bcalm in ((b"ACGT\n>0\nACGTA\n", False), (b">1 LN:i:5\nACGTA\n", True), (b">0 L:+:x:+\nACGTA\n", True),
                       (b">0 L:*:1:+\nACGTA\n", True), (b">0 L:+:7:+\nACGTA\n", True), (b">0\nACGNA\n", False), (b">0\nAC\n", True)):
        with pytest.raises(mt.MatchtigsError) as e:
            ctx.build_graph_from_text(bad, 5, bcalm)
        assert e.value.code == -3, bad
    ctx.build_graph_from_text(b"", 5, False)
    assert ctx.graph_info()["unitigs"] == 0
    ctx.build_graph_from_text(b">0 LN:i:5 L:+:0:+ L:+:1\nAAAAA", 5, True)  # short `L:` token is ignored like the host reader does
    assert ctx.graph_info()["unitigs"] == 1
    compare_all(mt, ctx, b">0\nACGTA\n", 5, "fasta")


def test_sharded_candidates_equal_full(mt, ctx):
    anc = tools.genome(20_000, 3, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 10, 9, snp_site_rate=0.04), 15)
    mt.read_bigraph_from_bcalm2_as_edge_centric(text, 15, ctx)
    ctx.dijkstra_candidates(8)
    fn, fd, fm = ctx.candidates_export()
    for R in (2, 3):
        for r in range(R):
            ctx.dijkstra_candidates(8, r, R)
            n, d, m = ctx.candidates_export()
            assert np.array_equal(m, fm[r::R])
            mask = np.arange(8)[None, :] < (m & META_COUNT)[:, None]  # slots past `count` are unspecified
            assert np.array_equal(n[mask], fn[r::R][mask]) and np.array_equal(d[mask], fd[r::R][mask])


def test_gathered_match_equals_single(mt, ctx):
    """Matching from slices laid out [shard][local] (what the NCCL all-gather produces) == single-GPU result."""
    import torch
    anc = tools.genome(20_000, 4, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 10, 11, snp_site_rate=0.04), 15)
    mt.read_bigraph_from_bcalm2_as_edge_centric(text, 15, ctx)
    ctx.dijkstra_candidates(8)
    want = ctx.greedy_match().copy()
    S = ctx.graph_info()["sources"]
    R, cap = 3, 8
    padded = (S + R - 1) // R
    rec = torch.zeros((R, padded, cap), dtype=torch.int64, device="cuda")
    meta = torch.zeros((R, padded), dtype=torch.int32, device="cuda")
    for r in range(R):
        ctx.dijkstra_candidates(cap, r, R)
        prec, pmeta, n, c = ctx.candidates_local()
        rec[r].copy_(mt.api.device_tensor(prec, (padded, cap), "<i8"))
        meta[r].copy_(mt.api.device_tensor(pmeta, (padded,), "<i4"))
    torch.cuda.synchronize()
    got = ctx.greedy_match(rec.data_ptr(), meta.data_ptr(), R)
    assert np.array_equal(got, want)


def test_reference_c_api(mt):
    """The five matchtigs_* symbols (src/clib.rs) against the oracle's C-API flavour."""
    l = mt._lib.load() if hasattr(mt, "_lib") else None
    from matchtigs_b200 import _lib
    l = _lib.load()
    rng = random.Random(31)
    U, k = 60, 9
    weights = np.array([rng.randint(1, 14) for _ in range(U)], dtype=np.uint64)
    # consistent links: derive them from a random arbitrary FASTA via the oracle-independent rule "suffix == prefix"
    seqs = []
    text = random_fasta(rng, U, k, max_extra=6, pool=8)
    from helpers import parse_fasta_seqs, revcomp
    seqs = parse_fasta_seqs(text)
    weights = np.array([len(s) - k + 1 for s in seqs], dtype=np.uint64)
    links = []
    ori = lambda s, f: s if f else revcomp(s)
    for a in range(U):
        for fa in (True, False):
            for b in range(U):
                for fb in (True, False):
                    if ori(seqs[a], fa)[-(k - 1):] == ori(seqs[b], fb)[:k - 1]:
                        links.append((a, fa, b, fb))
    o = oracle.Oracle().load_links(weights, links, k).run()
    l.matchtigs_initialise()
    h = l.matchtigs_initialise_graph(U)
    for a, fa, b, fb in links:
        l.matchtigs_merge_nodes(h, a, fa, b, fb)
    l.matchtigs_build_graph(h, weights.ctypes.data)
    eo, io_ = np.zeros(4 * U, np.int64), np.zeros(4 * U, np.uint64)
    lim = np.zeros(2 * U, np.uint64)
    n = l.matchtigs_compute_tigs(h, 5, 1, k, b"unused", b"unused", eo.ctypes.data, io_.ctypes.data, lim.ctypes.data)
    assert n == o.num("walks")
    ne = int(lim[n - 1])
    assert np.array_equal(eo[:ne], o.array("c_edge_out")) and np.array_equal(io_[:ne], o.array("c_insert_out"))
    assert np.array_equal(lim[:n], o.array("c_limits"))
    # algorithm 1 = unitigs
    h = l.matchtigs_initialise_graph(3)
    l.matchtigs_build_graph(h, np.array([1, 2, 3], np.uint64).ctypes.data)
    n = l.matchtigs_compute_tigs(h, 1, 1, 5, b"", b"", eo.ctypes.data, io_.ctypes.data, lim.ctypes.data)
    assert n == 3 and list(eo[:3]) == [0, 1, 2] and list(lim[:3]) == [1, 2, 3]


@pytest.mark.parametrize("name,scale", [("chr1", 0.04), ("pangenome", 0.06), ("human", 0.004)])
def test_other_configs_scaled(mt, ctx, name, scale):
    """BASELINE configs 3-5 at a scale the oracle finishes in seconds (repeat-rich, high-branching, k=51)."""
    text, k, info = tools.config_unitigs(name, scale)
    _, st = compare_all(mt, ctx, text, k, "fasta", cap=16)
    if name == "chr1":  # a few percent of its lists are longer than 16: searched again with 8x cap before the matching
        assert 0 < st["preextended_sources"] < info["unitigs"]
    compare_all(mt, ctx, text, k, "bcalm", cap=4)


def test_alternating_job_sizes_in_one_context(mt):
    """One context, jobs of very different sizes back to back (host-prepared tail below 2^17 nodes, device-prepared above;
    text parsed on the device): grow-only buffers, staged tail inputs and cached work space must never leak between jobs."""
    small, ks, _ = tools.config_unitigs("ecoli", 0.05)
    big, kb, info = tools.config_unitigs("chr1", 0.06)
    expected = {}
    for name, (text, k) in {"small": (small, ks), "big": (big, kb)}.items():
        o = oracle.Oracle(euler_fast=True)
        o.load_fasta(text, k)
        o.run()
        expected[name] = (o.text("gfa"), o.text("bitvector"), o.num("nodes"))
    assert expected["small"][2] < (1 << 17) <= expected["big"][2]
    c = mt.Context(0)
    try:
        for name in ["small", "big", "small", "small", "big", "small"]:
            text, k = (small, ks) if name == "small" else (big, kb)
            g = mt.read_bigraph_from_fasta_as_edge_centric(text, k, c, device_parse=True)
            mt.GreedytigAlgorithm.compute_tigs(g, mt.GreedytigAlgorithmConfiguration(k=k))
            assert mt.write_walks_gfa(g) == expected[name][0], name
            assert mt.write_duplication_bitvector(g) == expected[name][1], name
    finally:
        c.close()


def test_ecoli_scale_properties(mt, ctx):
    """BASELINE config 2 at 1/4 scale: byte identity + the size-independent properties."""
    text, k, info = tools.config_unitigs("ecoli", 0.25)
    compare_all(mt, ctx, text, k, "bcalm", dbg_valid=True, check_props=False)
    compare_all(mt, ctx, text, k, "fasta", dbg_valid=True, check_props=False)


def test_cli_end_to_end(mt, tmp_path):
    """`matchtigs`-compatible CLI: files in, files out, bytes equal to the oracle (incl. a gzipped output)."""
    import gzip
    from matchtigs_b200 import cli
    g = tools.genome(20_000, 77, families=4, copies=4, min_len=40, max_len=300, divergence=0.03)
    text, _, _ = tools.unitigs(g, 21)
    (tmp_path / "u.fa").write_bytes(text)
    for flag, mode in (("--bcalm-in", "bcalm"), ("--fa-in", "fasta")):
        o = run_oracle(text, 21, mode)
        rc = cli.main([flag, str(tmp_path / "u.fa"), "-k", "21", "--greedytigs-gfa-out", str(tmp_path / "o.gfa.gz"),
                       "--greedytigs-fa-out", str(tmp_path / "o.fa"), "--greedytigs-duplication-bitvector-out", str(tmp_path / "o.bv")])
        assert rc == 0
        assert gzip.open(tmp_path / "o.gfa.gz", "rb").read() == o.text("gfa")
        assert (tmp_path / "o.fa").read_bytes() == o.text("fasta")
        assert (tmp_path / "o.bv").read_bytes() == o.text("bitvector")


@pytest.mark.parametrize("name", ["p1_tie_desc", "p1_exclusive_bound", "p2_self_mirror_zero", "p3_oldest_first",
                                  "p6_bcalm_kmer_numbering", "p7_first_root_wins"])
def test_assumption_switches_flip_both_sides(mt, name):
    """every assumption about the un-vendored crates is a switch of the oracle AND of the product;
    flipped on both sides the CUDA path must again equal the oracle at every stage, on both readers, all search tiers."""
    c = mt.Context(0)
    try:
        c.set_option(name, 1)
        for seed in range(3):
            rng = random.Random(31_000 + seed)
            k = 2 * rng.randint(2, 8) + 1
            g = tools.genome(rng.randint(500, 6000), 70 + seed, families=rng.randint(1, 5), copies=rng.randint(2, 6), min_len=k,
                             max_len=rng.randint(k + 1, 150), divergence=rng.choice([0.0, 0.03, 0.1]), tandem_arrays=rng.randint(0, 3))
            text, _, _ = tools.unitigs(g, k)
            for mode in ("fasta", "bcalm"):
                compare_all(mt, c, text, k, mode, cap=rng.choice([2, 8, 16]), dbg_valid=True, check_props=True, opt={name: 1})
            # dense arbitrary records: long candidate lists, parallel edges, self-mirrors, the warp and CTA search tiers
            kk = rng.randint(3, 11)
            text = random_fasta(rng, rng.randint(50, 500), kk, max_extra=rng.choice([0, 4, 20]), pool=rng.choice([2, 4, 9]))
            compare_all(mt, c, text, kk, "fasta", cap=rng.choice([1, 4, 16]), opt={name: 1})
        with pytest.raises(mt.MatchtigsError):
            c.set_option("no_such_assumption", 1)
    finally:
        c.close()


def test_performance_counters_and_cli_lines(mt, ctx):
    """--dijkstra-performance-data-type Complete (greedytigs/mod.rs:647-673): the GPU path's counters in the reference's lines."""
    from matchtigs_b200 import cli
    g = tools.genome(30_000, 3, families=6, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=5)
    text, _, _ = tools.unitigs(g, 21)
    o, st = compare_all(mt, ctx, text, 21, "fasta", cap=16)
    # sources without a traversable out-edge settle themselves without a label table, truncated searches leave labels open
    assert st["labelled_nodes"] > 0 and st["settled_nodes"] > 0
    assert 1 <= st["max_open_nodes"] <= st["max_labelled_nodes"] <= st["labelled_nodes"]
    # the reference's search stops after m+1 targets; the GPU searches to the cap, so it labels at least as much per search
    assert st["max_labelled_nodes"] >= 1 and o.num("max_max_distance_array_size") >= 1
    lines = cli.performance_lines(st)
    assert lines[0] == "Dijkstras had a factor of 0.000 unnecessary heap elements"
    assert lines[1] == f"Dijktras had a maximum maximum heap size of {st['max_open_nodes']}"
    assert lines[2] == f"Dijktras had a maximum maximum distance array size of {st['max_labelled_nodes']}"


def test_library_nccl_entry_points_and_sharded_emission(mt):
    """The multi-GPU entry points of the C ABI on a one-rank communicator (the driver's test box has one GPU; N > 1 runs
    through bench.py): rendezvous, sliced upload + all-gather, candidate all-gather, walk broadcast, and the output texts
    assembled in shares -- every share at the offset it reports, the concatenation byte-identical to the oracle."""
    from matchtigs_b200 import sharding
    g = tools.genome(60_000, 8, families=8, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=6)
    text, _, _ = tools.unitigs(g, 21)
    c = mt.Context(0)
    try:
        c.comm_init(mt.Context.comm_unique_id(), 0, 1)
        for mode, bcalm in (("bcalm", True), ("fasta", False)):
            lo, hi = sharding.text_slice(len(text), 0, 1)
            c.build_graph_from_text_slices(np.frombuffer(text, np.uint8)[lo:hi], len(text), 21, bcalm)
            c.dijkstra_candidates(8, 0, 1)
            rec, meta = c.allgather_candidates()
            c.greedy_match(rec, meta, 1)
            c.finish_walks()
            c.broadcast_walks(0)
            o = run_oracle(text, 21, mode)
            T = c.walk_count()
            assert T == o.num("walks")
            for fmt, ref in (("gfa", o.text("gfa")), ("fasta", o.text("fasta")), (None, o.text("bitvector"))):
                for world in (1, 3, 7):
                    pieces, expect = [], 0
                    for r in range(world):
                        w_lo, w_hi = sharding.walk_share(T, r, world)
                        v, off, tot = (c.dup_bitvector_range_view(w_lo, w_hi) if fmt is None else c.assemble_tigs_range_view(fmt, w_lo, w_hi))
                        assert off == expect and tot == len(ref), (fmt, world, r)
                        pieces.append(v.tobytes())
                        expect += len(pieces[-1])
                    assert b"".join(pieces) == ref, (fmt, world)
        c.comm_destroy()
    finally:
        c.close()



@pytest.mark.parametrize("force_host", ["0", "1"])
def test_both_tail_preparations_on_the_same_inputs(mt, ctx, force_host, monkeypatch):
    """The walk records are built on the device for big graphs and on the host for small ones (< 2^17 nodes); MTG_TAIL_HOST
    forces either, so both builders see the inputs that stress them: nodes with more than four out-edges (header
    records), odd degrees (padding slots), self-mirrors, parallel edges, many components."""
    monkeypatch.setenv("MTG_TAIL_HOST", force_host)
    for seed in range(10):
        rng = random.Random(61_000 + seed)
        k = rng.randint(3, 11)
        text = random_fasta(rng, rng.randint(1, 700), k, max_extra=rng.choice([0, 2, 8]), pool=rng.choice([2, 3, 5, 9, 40, None]))
        compare_all(mt, ctx, text, k, "fasta", cap=rng.choice([2, 8, 16]))
    g = tools.genome(120_000, 77, families=10, copies=8, min_len=40, max_len=600, divergence=0.05, tandem_arrays=10)
    text, _, _ = tools.unitigs(g, 21)
    for mode in ("fasta", "bcalm"):
        compare_all(mt, ctx, text, 21, mode, cap=16, dbg_valid=True, check_props=True)
