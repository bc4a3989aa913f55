"""CPU tests of the product's host-side logic (no GPU): the C ABI loads and exports every declared symbol,
the record reader, and the host-sequential tail against the oracle on the oracle's own graph + triples."""
import os
import random
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
import tools
from helpers import random_fasta

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def mt():
    from matchtigs_b200 import _build
    _build.build_product()
    import matchtigs_b200
    return matchtigs_b200


def test_library_exports_every_declared_symbol(mt):
    from matchtigs_b200 import _lib
    lib = _lib.load()
    header = (ROOT / "include" / "matchtigs_b200.h").read_text()
    declared = set(re.findall(r"\b((?:mtg|matchtigs)_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared but not exported: {missing}"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)


def test_context_creation_fails_loudly_without_gpu(mt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mt.MatchtigsError) as e:
        mt.Context(0)
    assert e.value.code == -2  # MTG_ERR_CUDA: there is no CPU fallback


def test_reader_fasta_and_bcalm(mt):
    u = mt.Unitigs(b">0 LN:i:5 L:+:1:- L:-:0:+\nACGTA\n>1 LN:i:6\r\nACG\nTAC\n\n", True)
    assert u.count == 2 and list(u.offsets) == [0, 5, 11] and u.seq.tobytes() == b"ACGTAACGTAC"
    assert list(u.link_a) == [0, 0] and list(u.strand_a) == [1, 0] and list(u.link_b) == [1, 0] and list(u.strand_b) == [0, 1]
    assert mt.Unitigs(b"", False).count == 0
    for bad in (b"ACGT\n", b">1\nACGT\n", b">0 L:+:x:+\nACGT\n", b">0 L:*:1:+\nACGT\n"):
        with pytest.raises(mt.MatchtigsError):
            mt.Unitigs(bad, True)
    # the reader agrees with the oracle's reader on generated bcalm files
    g = tools.genome(20_000, 2, families=4, copies=4, min_len=40, max_len=200, divergence=0.03)
    text, _, nu = tools.unitigs(g, 21)
    u = mt.Unitigs(text, True)
    o = oracle.Oracle().load_bcalm(text, 21)
    assert u.count == nu == o.num("unitigs")
    assert np.array_equal((np.diff(u.offsets.astype(np.int64)) + 1 - 21).astype(np.uint64), o.array("edge_weight")[::2])


def oracle_inputs(text, k, mode, euler_fast=False):
    o = oracle.Oracle(euler_fast=euler_fast)
    (o.load_fasta if mode == "fasta" else o.load_bcalm)(text, k)
    o.run()
    U = o.num("unitigs")
    return o, (o.array("edge_from")[:2 * U], o.array("edge_to")[:2 * U], o.array("edge_weight")[:2 * U:2].astype(np.uint32),
               o.array("mirror"), o.array("triples"))


def check_tail(mt, text, k, mode, euler_fast=False):
    o, (ef, et, uw, mi, tr) = oracle_inputs(text, k, mode, euler_fast)
    walks, dummy_w, ms = mt.api.host_tail(k, ef, et, uw, mi, tr)
    ow = o.walks()
    assert len(walks) == len(ow)
    for a, b in zip(walks, ow):
        assert np.array_equal(a, b)
    U = o.num("unitigs")
    assert np.array_equal(dummy_w.astype(np.uint64), o.array("edge_weight")[2 * U:])


@pytest.mark.parametrize("seed", range(12))
def test_host_tail_equals_oracle_arbitrary(mt, seed):
    rng = random.Random(900 + seed)
    k = rng.choice([4, 5, 7, 8, 11])
    text = random_fasta(rng, rng.randint(1, 300), k, max_extra=10, pool=rng.choice([None, 3, 8, 25]))
    check_tail(mt, text, k, "fasta")


@pytest.mark.parametrize("mode", ["fasta", "bcalm"])
def test_host_tail_equals_oracle_dbg(mt, mode):
    anc = tools.genome(15_000, 8, families=3, copies=3, min_len=50, max_len=300, divergence=0.02)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 10, 5, snp_site_rate=0.04, indel_site_rate=0.003), 15)
    check_tail(mt, text, 15, mode)


def test_host_tail_threaded_preparation_with_hints(mt):
    """Above 2^18 edges / 2^17 nodes the host preparation runs on all cores (atomic placement + per-row sort) and builds
    the walk's prefetch hints; the walks must not change.  (The oracle's linear-time Euler variant is used: its faithful
    Vec::rotate_left one is quadratic; the two are compared with each other in test_oracle.py.)"""
    text, k, info = tools.config_unitigs("chr1", 0.05)
    assert info["unitigs"] > (1 << 17)
    check_tail(mt, text, k, "fasta", euler_fast=True)


WALK_SWITCHES = [{}, {"MTG_WALK_FAST": "0"}, {"MTG_WALK_SOURCES": "1"}, {"MTG_WALK_SOURCES": "3"}, {"MTG_WALK_NTSTORE": "0"},
                 {"MTG_WALK_PREFETCH": "t0"}, {"MTG_WALK_PREFETCH": "2"}, {"MTG_WALK_PROBE": "1"}, {"MTG_HOST_THREADS": "3"},
                 {"MTG_WALK_CHAIN": "carried"}, {"MTG_WALK_CHAIN": "carried", "MTG_WALK_SOURCES": "1"}, {"MTG_TRACE": "1"},
                 {"MTG_WALK_REPLAY": "0:4,19:5,13:2,49:3"}]


@pytest.mark.parametrize("seed", range(8))
def test_host_tail_with_lookahead_hints_on_small_graphs(mt, seed, monkeypatch):
    """The walk's lookahead machinery (hint levels, the written-out steady state, the second hint source, chain resets at
    nodes with three or four out-edges and at big nodes) normally only runs on graphs whose records outgrow the caches.
    MTG_TAIL_FORCEHINT=1 builds the hint levels for any graph, so that every variant of the run loop (the A/B switches
    select template instantiations and code paths) is compared with the oracle on small, repeat-rich, branching inputs."""
    rng = random.Random(7700 + seed)
    k = rng.choice([4, 5, 6, 8])  # small k: many nodes with three and four out-edges, palindromes, self-mirrors
    text = random_fasta(rng, rng.randint(30, 400), k, max_extra=10, pool=rng.choice([None, 3, 8, 25]))
    o, args = oracle_inputs(text, k, "fasta")
    ow = o.walks()
    monkeypatch.setenv("MTG_TAIL_FORCEHINT", "1")
    for env in WALK_SWITCHES:
        with monkeypatch.context() as m:
            for name, value in env.items():
                m.setenv(name, value)
            walks, _, _ = mt.api.host_tail(k, *args)
        assert len(walks) == len(ow), env
        assert all(np.array_equal(a, b) for a, b in zip(walks, ow)), env


def test_walk_lookahead_reaches_its_depth(mt):
    """The hints never change a result, so a broken lookahead would only show as a slower walk.  MTG_TRACE prints how many
    steps ahead every record of the walk was asked for; on a chr1-like graph (few nodes with more than two out-edges) the
    lean run loop must ask for most records a full WALK_DEPTH = 4 steps ahead and for next to none of them too late.
    (Own process: the library reads MTG_TRACE once.)"""
    import re
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, tools, matchtigs_b200 as mt\n"
        "from test_host_logic import oracle_inputs\n"
        "text, k, info = tools.config_unitigs('chr1', 0.05)\n"
        "o, args = oracle_inputs(text, k, 'fasta', euler_fast=True)\n"
        "walks, _, _ = mt.api.host_tail(k, *args)\n"
        "ow = o.walks()\n"
        "assert len(walks) == len(ow) and all(np.array_equal(a, b) for a, b in zip(walks, ow))\n"
    ) % (str(root), str(root / "tests"))
    env = dict(os.environ, MTG_TAIL_FORCEHINT="1", MTG_TRACE="1")
    for name in ("MTG_WALK_CHAIN", "MTG_WALK_PROBE", "MTG_WALK_SPIN", "MTG_TAIL_NOHINT"):
        env.pop(name, None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"steps by how many steps earlier their record was asked for \(0 = never\):((?: \d+:\d+)+)", r.stderr)
    assert m, r.stderr[-2000:]
    hist = {int(a): int(b) for a, b in (x.split(":") for x in m.group(1).split())}
    steps = sum(hist.values())
    assert steps > 100_000
    assert hist[4] >= 0.70 * steps, hist          # measured: 76 % on this graph, 75.5 % on chr1 x 1.0
    assert hist[0] + hist[1] <= 0.005 * steps, hist  # run starts and the rare slot used up in between


@pytest.mark.parametrize("seed", range(6))
def test_host_tail_heavy_matching_dummies_take_the_generic_breaking_path(mt, seed):
    """Behind the GPU matching a matching dummy always weighs less than k, and the tail recognises breaking dummies by edge
    id.  The host-only entry accepts arbitrary triples: with weights >= k every dummy breaks (greedytigs/mod.rs:767-777),
    found through the weight lookup.  The Euler cycle itself does not depend on the weights, so the walks must be exactly
    the dummy-free runs of the light result (same multiset: only the rotation start may move)."""
    rng = random.Random(4200 + seed)
    k = rng.choice([5, 7, 8])
    text = random_fasta(rng, rng.randint(40, 250), k, max_extra=8, pool=rng.choice([4, 9, 20]))
    o, (ef, et, uw, mi, tr) = oracle_inputs(text, k, "fasta")
    E0 = len(ef)
    light, _, _ = mt.api.host_tail(k, ef, et, uw, mi, tr)
    runs = []
    for w in light:
        cur = []
        for e in w.tolist():
            if e >= E0:
                if cur:
                    runs.append(tuple(cur))
                cur = []
            else:
                cur.append(e)
        if cur:
            runs.append(tuple(cur))
    heavy_tr = np.array(tr, dtype=np.uint32).reshape(-1, 3).copy()
    if len(heavy_tr):
        heavy_tr[:, 2] = k + rng.randint(0, 3)
    heavy, dummy_w, _ = mt.api.host_tail(k, ef, et, uw, mi, heavy_tr.reshape(-1))
    assert all((w < E0).all() for w in heavy), "a dummy of weight >= k survived inside a walk"
    assert sorted(tuple(w.tolist()) for w in heavy) == sorted(runs)
    assert (dummy_w >= k).all()


def test_host_tail_empty(mt):
    walks, dummy_w, _ = mt.api.host_tail(5, [], [], [], [], [])
    assert walks == [] and len(dummy_w) == 0


def test_cli_flag_surface(mt, tmp_path):
    """The stand-in CLI accepts the reference's flag names (src/bin.rs:56-205) and rejects reference-only work."""
    from matchtigs_b200 import cli
    p = cli.build_parser()
    a = p.parse_args(["--bcalm-in", "x.fa", "-k", "31", "-t", "4", "--greedytigs-gfa-out", "o.gfa", "--greedytigs-fa-out", "o.fa",
                      "--greedytigs-duplication-bitvector-out", "o.bv", "--dijkstra-node-weight-array-type", "EpochNodeWeightArray",
                      "--dijkstra-heap-type", "StdBinaryHeap", "--dijkstra-performance-data-type", "Complete",
                      "--dijkstra-staged-parallelism-divisor", "2.0", "--dijkstra-resource-limit-factor", "3",
                      "--log-level", "Debug", "--compression-level", "3", "--debug-print-walks"])
    assert a.k == 31 and a.threads == 4 and a.greedytigs_gfa_out == "o.gfa"
    for bad in (["--fa-in", "a", "--bcalm-in", "b", "-k", "5", "--greedytigs-fa-out", "o"],
                ["--gfa-in", "a", "--greedytigs-fa-out", "o"],
                ["--fa-in", "a", "-k", "5", "--matchtigs-fa-out", "o"],
                ["--fa-in", "a", "-k", "5", "--eulertigs-gfa-out", "o"],
                ["--fa-in", "a", "--greedytigs-fa-out", "o"]):
        with pytest.raises(SystemExit) as e:
            cli.main(bad)
        assert e.value.code not in (0, None)


def test_rust_shim_declares_exported_symbols():
    """rust/b200.rs (the binding a maintainer of the Rust host would add) and the built library agree on symbol names."""
    import ctypes
    import re
    from pathlib import Path
    from matchtigs_b200 import _lib
    root = Path(__file__).resolve().parent.parent
    src = (root / "rust" / "b200.rs").read_text()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    names = re.findall(r"pub fn (mtg_\w+)\(", block)
    assert len(names) >= 30
    lib = ctypes.CDLL(str(_lib.SO_PATH))
    for n in names:
        assert hasattr(lib, n), f"rust shim declares {n}, the library does not export it"
    header = (root / "include" / "matchtigs_b200.h").read_text()
    for n in names:
        assert re.search(rf"\b{n}\(", header), f"{n} is not declared in include/matchtigs_b200.h"
    # the #[repr(C)] stats struct has the fields of the C struct, in order
    fields_rs = re.findall(r"pub (\w+): [uf]\d+,", src[src.index("pub struct mtg_search_stats"):src.index('#[link(name')])
    assert fields_rs == [n for n, _ in _lib.SearchStats._fields_]
