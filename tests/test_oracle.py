"""CPU tests of the oracle itself: invariants of the reference's algorithm on
synthetic inputs, hand-derived micro cases, and agreement of its two Euler-decomposition variants.
The reference holds no golden vectors for this path, so these properties are what pins the oracle."""
import random

import numpy as np
import pytest

import oracle
import tools
from helpers import check_tig_invariants, parse_gfa_seqs, random_fasta, revcomp


def run(text, k, mode="fasta", fast=False, threads=1):
    o = oracle.Oracle(euler_fast=fast)
    (o.load_fasta if mode == "fasta" else o.load_bcalm)(text, k)
    return o.run(threads)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("mode", ["fasta", "bcalm"])
def test_dbg_unitigs_invariants(seed, mode):
    g = tools.genome(30_000, 10 + seed, families=6, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=5)
    text, nk, nu = tools.unitigs(g, 21)
    o = run(text, 21, mode)
    check_tig_invariants(text, 21, o.text("gfa"), o.text("fasta"), o.text("bitvector"), dbg_valid=True)
    assert o.num("walks") < nu  # compression happened


@pytest.mark.parametrize("seed", range(3))
def test_pangenome_invariants(seed):
    anc = tools.genome(8_000, 50 + seed, families=2, copies=3, min_len=50, max_len=200, divergence=0.02)
    strains = tools.pangenome(anc, 12, 7 + seed, snp_site_rate=0.03, indel_site_rate=0.002)
    text, nk, nu = tools.unitigs(strains, 15)
    for mode in ("fasta", "bcalm"):
        o = run(text, 15, mode)
        check_tig_invariants(text, 15, o.text("gfa"), o.text("fasta"), o.text("bitvector"), dbg_valid=True)
        assert len(o.array("triples")) > 0


@pytest.mark.parametrize("k", [4, 5, 7, 8, 11])
@pytest.mark.parametrize("seed", range(6))
def test_arbitrary_fasta_invariants(k, seed):
    rng = random.Random(1000 * k + seed)
    text = random_fasta(rng, rng.randint(1, 60), k, max_extra=12, pool=rng.choice([None, 3, 8]))
    o = run(text, k)
    check_tig_invariants(text, k, o.text("gfa"), o.text("fasta"), o.text("bitvector"), dbg_valid=False)
    # walk invariants (greedytigs/mod.rs:794-798): originals at both ends, every biedge exactly once
    dummy = o.array("edge_dummy_id")
    seen = []
    for w in o.walks():
        assert dummy[w[0]] == 0 and dummy[w[-1]] == 0
        seen.extend(int(e) >> 1 for e in w if dummy[e] == 0)
    assert sorted(seen) == list(range(o.num("unitigs")))


@pytest.mark.parametrize("seed", range(8))
def test_euler_fast_equals_faithful(seed):
    rng = random.Random(77 + seed)
    k = rng.choice([5, 7, 9])
    text = random_fasta(rng, rng.randint(20, 300), k, max_extra=10, pool=rng.choice([4, 10, 25]))
    a, b = run(text, k, fast=False), run(text, k, fast=True)
    ca, cb = a.cycles(), b.cycles()
    assert len(ca) == len(cb)
    for x, y in zip(ca, cb):
        assert np.array_equal(x, y)
    assert a.text("gfa") == b.text("gfa")


def test_micro_single_unitig():
    # one unitig, k=5: two tips; the greedy search finds nothing, eulerise joins the ends with a breaking edge.
    text = b">0\nACGTTGCA\n"  # palindromic as a whole: prefix ACGT is its own reverse complement
    o = run(text, 5)
    assert parse_gfa_seqs(o.text("gfa")) in ([b"ACGTTGCA"], [revcomp(b"ACGTTGCA")])
    assert o.text("bitvector") == b"1111\n"


def test_micro_two_overlapping_unitigs_merge():
    # u0 ends with the 4-mer u1 starts with: one walk covers both, no dummy needed (k=5).
    text = b">0\nAACCGGA\n>1\nCGGATT\n"
    o = run(text, 5)
    tigs = parse_gfa_seqs(o.text("gfa"))
    assert len(tigs) == 1 and tigs[0] in (b"AACCGGATT", revcomp(b"AACCGGATT"))
    assert o.text("bitvector") == b"11111\n"


def test_micro_tip_is_bridged_by_short_unitig():
    # A branch: X -> {Y (short, weight 1), Z}; greedy matching joins the out-imbalanced end of one
    # branch to the in-imbalanced node through the short unitig, repeating its k-mers ('0' bits).
    k = 5
    text = b">0\nAAAACCCC\n>1\nCCCCG\n>2\nCCCCT\n>3\nCCCGAAT\n"
    o = run(text, k)
    check_tig_invariants(text, k, o.text("gfa"), o.text("fasta"), o.text("bitvector"), dbg_valid=False)


def test_threads_variant_same_kmers():
    g = tools.genome(40_000, 5, families=8, copies=6, min_len=40, max_len=300, divergence=0.03)
    text, _, _ = tools.unitigs(g, 21)
    o = run(text, 21, "bcalm", threads=4)
    check_tig_invariants(text, 21, o.text("gfa"), o.text("fasta"), o.text("bitvector"), dbg_valid=True)


def test_errors_are_reported():
    o = oracle.Oracle()
    with pytest.raises(oracle.OracleError):
        o.load_fasta(b">0\nACGNT\n", 3)
    with pytest.raises(oracle.OracleError):
        o.load_fasta(b">0\nAC\n", 5)
    with pytest.raises(oracle.OracleError):
        o.load_bcalm(b">1 LN:i:5\nACGTA\n", 5)


def test_capi_encoding_matches_walks():
    # src/clib.rs:393-407
    rng = random.Random(5)
    text = random_fasta(rng, 40, 7, pool=6)
    o = run(text, 7)
    eo, io_, lim = o.array("c_edge_out"), o.array("c_insert_out"), o.array("c_limits")
    walks = o.walks()
    assert len(lim) == len(walks) and lim[-1] == len(eo) == len(io_)
    dummy, fwd, uni, w = o.array("edge_dummy_id"), o.array("edge_forward"), o.array("edge_unitig"), o.array("edge_weight")
    pos = 0
    for i, wk in enumerate(walks):
        for e in wk:
            if dummy[e]:
                assert eo[pos] == 0 and io_[pos] == w[e]
            else:
                assert eo[pos] == (int(uni[e]) if fwd[e] else -int(uni[e])) and io_[pos] == 0
            pos += 1
        assert lim[i] == pos


# ---- parity-assumption switches and the reference's Dijkstra performance counters ----
ASSUMPTIONS = ("p1_tie_desc", "p1_exclusive_bound", "p2_self_mirror_zero", "p3_oldest_first", "p6_bcalm_kmer_numbering",
               "p7_first_root_wins")


def run_opt(text, k, mode, fast=True, **opt):
    o = oracle.Oracle(euler_fast=fast, **opt)
    (o.load_fasta if mode == "fasta" else o.load_bcalm)(text, k)
    return o.run()


@pytest.mark.parametrize("name", ASSUMPTIONS)
def test_every_assumption_switch_gives_a_valid_tig_set(name):
    """Whichever way an assumption about the un-vendored crates falls, the result must still be a correct tig set
    (spectrum preserved, bitvector property, walk invariants); and each switch must be able to change the bytes."""
    changed = False
    for seed in range(6):
        rng = random.Random(4200 + seed)
        k = rng.choice([5, 7, 9])
        text = random_fasta(rng, rng.randint(30, 200), k, max_extra=8, pool=rng.choice([3, 6, 12]))
        mode = "bcalm" if name.startswith(("p6", "p7")) else "fasta"
        if mode == "bcalm":  # links need a real compacted graph
            g = tools.genome(3000 + 500 * seed, 900 + seed, families=4, copies=5, min_len=k, max_len=80, divergence=0.05, tandem_arrays=2)
            text, _, _ = tools.unitigs(g, k)
        base, alt = run_opt(text, k, mode), run_opt(text, k, mode, **{name: 1})
        check_tig_invariants(text, k, alt.text("gfa"), alt.text("fasta"), alt.text("bitvector"), dbg_valid=(mode == "bcalm"))
        changed |= alt.text("gfa") != base.text("gfa")
    assert changed, f"{name} never changed an output: the switch is not wired"


def test_p3_fast_euler_equals_faithful_under_the_switch():
    for seed in range(4):
        rng = random.Random(5100 + seed)
        text = random_fasta(rng, rng.randint(20, 200), 7, max_extra=10, pool=rng.choice([4, 10]))
        a, b = run_opt(text, 7, "fasta", fast=False, p3_oldest_first=1), run_opt(text, 7, "fasta", fast=True, p3_oldest_first=1)
        assert a.text("gfa") == b.text("gfa")


def test_registered_assumptions_without_alternative_refuse_nonzero():
    o = oracle.Oracle()
    o.set_option("p4_euler_policy", 0)
    o.set_option("p5_fasta_numbering", 0)
    with pytest.raises(KeyError):
        o.set_option("p4_euler_policy", 1)
    with pytest.raises(KeyError):
        o.set_option("no_such_option", 1)


def test_reference_performance_counters():
    """DijkstraPerformanceCounter semantics (greedytigs/mod.rs:647-673): iterations = heap pops, unnecessary heap elements =
    stale pops, heap / distance-array sizes as per-search maxima."""
    g = tools.genome(30_000, 3, families=6, copies=6, min_len=40, max_len=400, divergence=0.03, tandem_arrays=5)
    text, _, _ = tools.unitigs(g, 21)
    o = run_opt(text, 21, "fasta")
    it, un, calls = o.num("iterations"), o.num("unnecessary_heap_elements"), o.num("dijkstra_calls")
    assert it == o.num("settled") + un or it >= o.num("settled")  # the pop that exceeds the bound is an iteration, not a settle
    assert 0 <= un < it and calls > 0
    assert 1 <= o.num("max_max_heap_size") <= o.num("sum_max_heap_size") <= calls * o.num("max_max_heap_size")
    assert 1 <= o.num("max_max_distance_array_size") <= o.num("sum_max_distance_array_size")
    assert o.num("max_max_heap_size") <= o.num("sum_max_distance_array_size")
