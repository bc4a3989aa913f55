"""`matchtigs`-flag-compatible command line for the greedy path (reference: ``src/bin.rs:56-205``, main ``:850-1218``).

Stand-in for the Rust CLI (which cannot be compiled in this image): same flag names and meaning for
everything the greedy path uses; flags that select reference-only algorithms are rejected.

    python -m matchtigs_b200.cli --bcalm-in unitigs.fa -k 31 --greedytigs-gfa-out out.gfa \
        --greedytigs-duplication-bitvector-out out.bv
"""
from __future__ import annotations

import argparse
import gzip
import sys
import time


def _read(path: str) -> bytes:
    if path.endswith(".gz"):  # src/bin.rs:894, :905
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def _write(path: str, data, level: int) -> None:
    if path.endswith(".gz"):  # src/bin.rs:442-446, :634-638
        with gzip.open(path, "wb", compresslevel=level) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="matchtigs", description="Matchtigs: minimum plain text representation of kmer sets "
                                                             "(B200 greedy path).")
    p.add_argument("--gfa-in")
    p.add_argument("--fa-in")
    p.add_argument("--bcalm-in")
    for name in ("pathtigs", "eulertigs", "matchtigs"):
        p.add_argument(f"--{name}-gfa-out")
        p.add_argument(f"--{name}-fa-out")
    p.add_argument("--greedytigs-gfa-out")
    p.add_argument("--greedytigs-fa-out")
    p.add_argument("--greedytigs-duplication-bitvector-out")
    p.add_argument("--matchtigs-duplication-bitvector-out")
    p.add_argument("-k", type=int)
    p.add_argument("-t", "--threads", type=int, default=1)
    p.add_argument("--blossom5-command", default="blossom5")
    p.add_argument("--dijkstra-node-weight-array-type", default="HashbrownHashMap",
                   choices=["EpochNodeWeightArray", "HashbrownHashMap"])
    p.add_argument("--dijkstra-heap-type", default="StdBinaryHeap", choices=["StdBinaryHeap"])
    p.add_argument("--dijkstra-performance-data-type", default="None", choices=["None", "Complete"])
    p.add_argument("--dijkstra-staged-parallelism-divisor", type=float)
    p.add_argument("--dijkstra-resource-limit-factor", type=int, default=1)
    p.add_argument("--debug-print-graph", action="store_true")
    p.add_argument("--debug-print-walks", action="store_true")
    p.add_argument("--log-level", default="Info")
    p.add_argument("--compression-level", type=int, default=6)
    p.add_argument("--device", type=int, default=0, help="CUDA device (not a reference flag)")
    p.add_argument("--candidate-cap", type=int, default=16, help="candidate list depth per source (not a reference flag; "
                                                                 "results do not depend on it)")
    return p


def performance_lines(st: dict) -> list[str]:
    """The lines the reference logs under ``--dijkstra-performance-data-type Complete`` (greedytigs/mod.rs:647-673, typos
    included), fed with what the counters mean on the GPU path: a search extracts every label once (no lazy deletion, so
    iterations == settled nodes and no unnecessary heap elements), its open labels are the heap, its label table the
    distance array.  Averages are over the searches that ran."""
    searches = max(st["sources_searched"], 1)
    iterations = max(st["settled_nodes"], 1)
    return [
        f"Dijkstras had a factor of {0 / iterations:.3f} unnecessary heap elements",
        f"Dijktras had a maximum maximum heap size of {st['max_open_nodes']}",
        f"Dijktras had a maximum maximum distance array size of {st['max_labelled_nodes']}",
        f"Dijktras had an average maximum distance array size of {st['labelled_nodes'] / searches:.0f}",
        f"Dijkstras settled {st['settled_nodes']} nodes and relaxed {st['relaxed_edges']} edges "
        f"in {st['dijkstra_ms']:.3f} ms on the device",
    ]


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    inputs = [x for x in (args.fa_in, args.gfa_in, args.bcalm_in) if x]
    if len(inputs) != 1:  # src/bin.rs:852-862
        sys.exit("Specify exactly one of --fa-in, --gfa-in or --bcalm-in")
    if args.gfa_in:
        sys.exit("--gfa-in is served by the reference only (out of scope of the B200 greedy path)")
    for name in ("pathtigs", "eulertigs", "matchtigs"):
        if getattr(args, f"{name}_gfa_out") or getattr(args, f"{name}_fa_out") or (name == "matchtigs" and args.matchtigs_duplication_bitvector_out):
            sys.exit(f"{name} are computed by the reference only (out of scope of the B200 greedy path)")
    if args.k is None:
        sys.exit("-k is required with --fa-in / --bcalm-in")  # src/bin.rs:74, :82
    if not 0 <= args.compression_level <= 9:
        sys.exit("valid compression levels are 0 to 9")  # src/bin.rs:207-217
    if not (args.greedytigs_gfa_out or args.greedytigs_fa_out or args.greedytigs_duplication_bitvector_out):
        print("Nothing to compute: no greedytigs output requested", file=sys.stderr)
        return 0

    import matchtigs_b200 as mt
    t0 = time.time()
    ctx = mt.Context(args.device)
    text = _read(inputs[0])
    if args.fa_in:
        graph = mt.read_bigraph_from_fasta_as_edge_centric(text, args.k, ctx)
    else:
        graph = mt.read_bigraph_from_bcalm2_as_edge_centric(text, args.k, ctx)
    t1 = time.time()
    info = ctx.graph_info()
    print(f"Loading took {t1 - t0:.1f} seconds", file=sys.stderr)
    print(f"k = {args.k}\nGraph has {info['nodes']} nodes and {info['edges']} edges", file=sys.stderr)
    if args.debug_print_graph:
        ex = ctx.graph_export()
        for e, (a, b) in enumerate(zip(ex["edge_from"], ex["edge_to"])):
            print(f"{e} ({a} -> {b})")
    cfg = mt.GreedytigAlgorithmConfiguration(k=args.k, threads=args.threads,
                                             staged_parallelism_divisor=args.dijkstra_staged_parallelism_divisor,
                                             resource_limit_factor=args.dijkstra_resource_limit_factor,
                                             heap_type=args.dijkstra_heap_type,
                                             node_weight_array_type=args.dijkstra_node_weight_array_type,
                                             performance_data_type=args.dijkstra_performance_data_type,
                                             candidate_cap=args.candidate_cap)
    walks = mt.GreedytigAlgorithm.compute_tigs(graph, cfg)
    t2 = time.time()
    st = ctx.search_stats()
    print(f"Found {st['matched']} shortest paths\nFound {len(walks)} greedytigs", file=sys.stderr)
    if args.dijkstra_performance_data_type == "Complete":
        print("\n".join(performance_lines(st)), file=sys.stderr)
    if args.debug_print_walks:
        for w in walks:
            print(" ".join(str(int(e)) for e in w))
    if args.greedytigs_fa_out:
        _write(args.greedytigs_fa_out, mt.write_walks_fasta(graph), args.compression_level)
    if args.greedytigs_gfa_out:
        _write(args.greedytigs_gfa_out, mt.write_walks_gfa(graph), args.compression_level)
    if args.greedytigs_duplication_bitvector_out:
        # never gzipped, whatever the name (src/implementation/mod.rs:665)
        with open(args.greedytigs_duplication_bitvector_out, "wb") as f:
            f.write(mt.write_duplication_bitvector(graph))
    t3 = time.time()
    print(f"Computing greedytigs took {t2 - t1:.1f}s and writing took {t3 - t2:.1f}s\nDone", file=sys.stderr)
    ctx.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
