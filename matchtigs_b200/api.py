"""Host-side mirror of the reference's greedy-matchtig interface on top of the C ABI.

Names follow the reference so that call sites read alike:

* ``read_bigraph_from_fasta_as_edge_centric`` / ``read_bigraph_from_bcalm2_as_edge_centric``
  (reference call sites ``src/bin.rs:896-899``, ``:907-910``)
* ``GreedytigAlgorithmConfiguration`` / ``GreedytigAlgorithm.compute_tigs``
  (``src/implementation/greedytigs/mod.rs:40-90``)
* ``write_walks_gfa`` / ``write_walks_fasta`` / ``write_duplication_bitvector``
  (``src/bin.rs:667-818``, ``:466-606``, ``src/implementation/mod.rs:671-702``)

All computation happens in ``libmatchtigs_b200.so`` on the GPU; this module only moves bytes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class MatchtigsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_lib.STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class _DeviceArray:
    """Minimal ``__cuda_array_interface__`` holder so torch can wrap library-owned device memory."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_tensor(ptr: int, shape: tuple, typestr: str):
    """Zero-copy torch view of device memory owned by a ``Context`` (used for the NCCL all-gather)."""
    import torch
    return torch.as_tensor(_DeviceArray(ptr, shape, typestr), device="cuda")


def _ptr(a: np.ndarray | None):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Unitigs:
    """Parsed unitig records (host memory): concatenated sequence, offsets, bcalm links."""

    def __init__(self, text: bytes, bcalm: bool):
        l = _lib.load()
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = l.mtg_unitigs_parse(text, len(text), int(bcalm), C.byref(h), err, len(err))
        if rc != 0:
            raise MatchtigsError(rc, err.value.decode())
        self._l, self._h = l, h
        seq, off, la, sa, lb, sb = (C.c_void_p() for _ in range(6))
        u, nl = C.c_uint64(), C.c_uint64()
        l.mtg_unitigs_view(h, C.byref(seq), C.byref(off), C.byref(u), C.byref(la), C.byref(sa), C.byref(lb), C.byref(sb),
                           C.byref(nl))
        self.count = u.value
        self.n_links = nl.value

        def view(p, n, dt):
            if n == 0 or not p.value:
                return np.zeros(0, dtype=dt)
            return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt)

        self.offsets = view(off, self.count + 1, np.uint64)
        self.seq = view(seq, int(self.offsets[-1]) if self.count else 0, np.uint8)
        self.link_a, self.link_b = view(la, self.n_links, np.uint64), view(lb, self.n_links, np.uint64)
        self.strand_a, self.strand_b = view(sa, self.n_links, np.uint8), view(sb, self.n_links, np.uint8)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.mtg_unitigs_free(self._h)
            self._h = None


class Context:
    """One GPU context (``mtg_ctx``): the resident graph and every later step live here."""

    def __init__(self, device: int = 0):
        self._l = _lib.load()
        h = C.c_void_p()
        rc = self._l.mtg_ctx_create(C.byref(h), device)
        if rc != 0:
            raise MatchtigsError(rc, "mtg_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._l.mtg_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc: int):
        if rc != 0:
            raise MatchtigsError(rc, self._l.mtg_last_error(self._h).decode())

    ASSUMPTIONS = ("p1_tie_desc", "p1_exclusive_bound", "p2_self_mirror_zero", "p3_oldest_first", "p6_bcalm_kmer_numbering",
                   "p7_first_root_wins")

    def set_option(self, name: str, value: int):
        """Parity-assumption switches (same names as the oracle's ``set_option``)."""
        self._check(self._l.mtg_ctx_set_option(self._h, name.encode(), int(value)))

    # ---- step 1 ----
    def build_graph_from_sequences(self, seq: np.ndarray, offsets: np.ndarray, k: int):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._l.mtg_build_graph_from_sequences(self._h, _ptr(seq), _ptr(offsets), len(offsets) - 1, k, 0))

    def build_graph_from_device_sequences(self, seq_ptr: int, offsets_ptr: int, unitigs: int, k: int):
        self._check(self._l.mtg_build_graph_from_sequences(self._h, C.c_void_p(seq_ptr), C.c_void_p(offsets_ptr), unitigs, k, 1))

    def build_graph_from_links(self, weights, link_a, strand_a, link_b, strand_b, k: int, seq=None, offsets=None):
        w = np.ascontiguousarray(weights, dtype=np.uint64)
        la, lb = np.ascontiguousarray(link_a, dtype=np.uint64), np.ascontiguousarray(link_b, dtype=np.uint64)
        sa, sb = np.ascontiguousarray(strand_a, dtype=np.uint8), np.ascontiguousarray(strand_b, dtype=np.uint8)
        if seq is not None:
            seq = np.ascontiguousarray(seq, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._l.mtg_build_graph_from_links(self._h, len(w), _ptr(w), len(la), _ptr(la), _ptr(sa), _ptr(lb), _ptr(sb),
                                                       k, _ptr(seq), _ptr(offsets)))

    def build_graph_from_text(self, text, k: int, bcalm: bool, device_ptr: int | None = None, length: int | None = None):
        """Device-side record parsing + graph build.  `text`: bytes / uint8 array on the host, or pass device_ptr + length."""
        if device_ptr is not None:
            self._check(self._l.mtg_build_graph_from_text(self._h, C.c_void_p(device_ptr), length, int(bcalm), k, 1))
            return
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        self._check(self._l.mtg_build_graph_from_text(self._h, _ptr(buf), len(buf), int(bcalm), k, 0))

    def graph_info(self) -> dict:
        info = _lib.GraphInfo()
        self._check(self._l.mtg_graph_get_info(self._h, C.byref(info)))
        return info.as_dict()

    def graph_export(self) -> dict:
        gi = self.graph_info()
        out = {"edge_from": np.zeros(gi["edges"], np.uint32), "edge_to": np.zeros(gi["edges"], np.uint32),
               "mirror": np.zeros(gi["nodes"], np.uint32), "imbalance": np.zeros(gi["nodes"], np.int32),
               "sources": np.zeros(gi["sources"], np.uint32)}
        self._check(self._l.mtg_graph_export(self._h, _ptr(out["edge_from"]), _ptr(out["edge_to"]), _ptr(out["mirror"]),
                                             _ptr(out["imbalance"]), _ptr(out["sources"])))
        return out

    # ---- step 2 ----
    def dijkstra_candidates(self, cap: int = 8, shard_rank: int = 0, shard_count: int = 1):
        self._check(self._l.mtg_dijkstra_candidates(self._h, cap, shard_rank, shard_count))

    def candidates_local(self):
        """(device pointer of records, device pointer of meta, local sources, cap)."""
        rec, meta, n, cap = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint32()
        self._check(self._l.mtg_candidates_local(self._h, C.byref(rec), C.byref(meta), C.byref(n), C.byref(cap)))
        return rec.value, meta.value, n.value, cap.value

    def candidates_export(self):
        _, _, n, cap = self.candidates_local()
        nodes, dists = np.zeros((n, cap), np.uint32), np.zeros((n, cap), np.uint32)
        meta = np.zeros(n, np.uint32)
        self._check(self._l.mtg_candidates_export(self._h, _ptr(nodes), _ptr(dists), _ptr(meta)))
        return nodes, dists, meta

    # ---- step 3 ----
    def greedy_match(self, records_ptr: int | None = None, meta_ptr: int | None = None, shard_count: int = 1,
                     export: bool = True) -> np.ndarray | int:
        """Runs the matching.  Returns the (out node, in node, distance) triples, or with ``export=False`` only their number
        (the triples stay inside the context for ``finish_walks``; copying 12 bytes per pair out is the caller's choice)."""
        n = C.c_uint64()
        self._check(self._l.mtg_greedy_match(self._h, C.c_void_p(records_ptr or 0), C.c_void_p(meta_ptr or 0), shard_count,
                                             C.byref(n)))
        if not export:
            return n.value
        tr = np.empty(3 * n.value, np.uint32)
        self._check(self._l.mtg_triples_export(self._h, _ptr(tr)))
        return tr.reshape(-1, 3)

    def finish_walks(self):
        nw, ne = C.c_uint64(), C.c_uint64()
        self._check(self._l.mtg_finish_walks(self._h, C.byref(nw), C.byref(ne)))
        self._n_walks, self._n_walk_edges = nw.value, ne.value
        return nw.value, ne.value

    def walks(self) -> list[np.ndarray]:
        edges, limits = np.zeros(self._n_walk_edges, np.uint32), np.zeros(self._n_walks, np.uint64)
        self._check(self._l.mtg_walks_export(self._h, _ptr(edges), _ptr(limits)))
        return np.split(edges, limits[:-1].astype(np.int64)) if self._n_walks else []

    def walks_capi(self, unitigs: int):
        eo, io_ = np.zeros(4 * unitigs + 4, np.int64), np.zeros(4 * unitigs + 4, np.uint64)
        lim = np.zeros(2 * unitigs + 2, np.uint64)
        self._check(self._l.mtg_walks_export_capi(self._h, _ptr(eo), _ptr(io_), _ptr(lim)))
        return eo[:self._n_walk_edges], io_[:self._n_walk_edges], lim[:self._n_walks]

    def _view(self, fn, *args) -> np.ndarray:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(fn(self._h, *args, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_char * n.value).from_address(p.value), dtype=np.uint8)

    def dup_bitvector_view(self) -> np.ndarray:
        """Bitvector bytes as a zero-copy uint8 view of the context's pinned buffer (valid until the next call)."""
        return self._view(self._l.mtg_dup_bitvector_view)

    def assemble_tigs_view(self, fmt: str = "gfa") -> np.ndarray:
        return self._view(self._l.mtg_assemble_tigs_view, 0 if fmt == "gfa" else 1)

    def dup_bitvector(self) -> bytes:
        return self.dup_bitvector_view().tobytes()

    def assemble_tigs(self, fmt: str = "gfa") -> bytes:
        return self.assemble_tigs_view(fmt).tobytes()

    # ---- multi-GPU (one process per GPU; NCCL inside the library) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0 creates the rendezvous id; ship the bytes to the other ranks any way you like."""
        buf = C.create_string_buffer(128)
        rc = _lib.load().mtg_comm_get_unique_id(buf)
        if rc != 0:
            raise MatchtigsError(rc, "NCCL is not available (mtg_comm_get_unique_id)")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        self._check(self._l.mtg_comm_init(self._h, C.c_char_p(unique_id), rank, world))
        self.rank, self.world = rank, world

    def comm_destroy(self):
        self._check(self._l.mtg_comm_destroy(self._h))

    def allgather_candidates(self):
        """Collective: all-gather of the candidate slices -> (records ptr, meta ptr) for ``greedy_match(rec, meta, world)``."""
        rec, meta = C.c_void_p(), C.c_void_p()
        self._check(self._l.mtg_allgather_candidates(self._h, C.byref(rec), C.byref(meta)))
        return rec.value, meta.value

    def build_graph_from_text_slices(self, part: np.ndarray, total_len: int, k: int, bcalm: bool):
        """Collective: every rank supplies its slice of the file (``sharding.text_slice``); the whole text is assembled over NVLink."""
        part = np.ascontiguousarray(part, np.uint8)
        self._check(self._l.mtg_build_graph_from_text_slices(self._h, _ptr(part), len(part), total_len, int(bcalm), k))

    def broadcast_walks(self, root: int = 0):
        self._check(self._l.mtg_broadcast_walks(self._h, root))

    def walk_count(self) -> int:
        n = C.c_uint64()
        self._check(self._l.mtg_walk_count(self._h, C.byref(n)))
        return n.value

    def _range_view(self, fn, *args):
        p, n, off, tot = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(fn(self._h, *args, C.byref(p), C.byref(n), C.byref(off), C.byref(tot)))
        view = np.zeros(0, np.uint8) if n.value == 0 else np.frombuffer((C.c_char * n.value).from_address(p.value), dtype=np.uint8)
        return view, off.value, tot.value

    def dup_bitvector_range_view(self, walk_lo: int, walk_hi: int):
        """(bytes of the walks [lo, hi), their offset in the whole bitvector text, length of the whole text)."""
        return self._range_view(self._l.mtg_dup_bitvector_range_view, walk_lo, walk_hi)

    def assemble_tigs_range_view(self, fmt: str, walk_lo: int, walk_hi: int):
        return self._range_view(self._l.mtg_assemble_tigs_range_view, 0 if fmt == "gfa" else 1, walk_lo, walk_hi)

    def search_stats(self) -> dict:
        st = _lib.SearchStats()
        self._check(self._l.mtg_get_search_stats(self._h, C.byref(st)))
        return st.as_dict()

    def diagnostics(self) -> dict:
        ms, bms = (C.c_double * 5)(), (C.c_double * 2)()
        self._check(self._l.mtg_get_diagnostics(self._h, ms, bms))
        # prepare: degree counting (host-prepared tail) / record kernels incl. one round trip (device-prepared tail);
        # records: building them on the host / their DMA + copy into the walk's arena
        names = ["prepare", "eulerise", "records", "euler_walk", "breaking"]
        return {"tail_ms": {n: round(float(v), 3) for n, v in zip(names, ms)},
                "build_ms": {"parse": float(bms[0]), "graph": float(bms[1])}}

    @property
    def kernel_launches(self) -> int:
        return self._l.mtg_ctx_kernel_launches(self._h)

    @property
    def stream(self) -> int:
        return self._l.mtg_ctx_stream(self._h) or 0


def host_tail(k: int, edge_from, edge_to, unitig_w, mirror, triples):
    """The host-sequential tail (dummy insertion, eulerise, Euler decomposition, breaking) on host arrays; no GPU.

    Returns (walks, dummy_w, phase_ms) with walks as a list of edge-id arrays (ids >= 2U are dummy edges)."""
    l = _lib.load()
    ef, et = np.ascontiguousarray(edge_from, np.uint32), np.ascontiguousarray(edge_to, np.uint32)
    uw, mi = np.ascontiguousarray(unitig_w, np.uint32), np.ascontiguousarray(mirror, np.uint32)
    tr = np.ascontiguousarray(triples, np.uint32).reshape(-1)
    we, wl, dw = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nw, ne, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
    ms = (C.c_double * 5)()
    err = C.create_string_buffer(256)
    rc = l.mtg_host_tail(k, len(mi), len(uw), _ptr(ef), _ptr(et), _ptr(uw), _ptr(mi), _ptr(tr), len(tr) // 3, C.byref(we),
                         C.byref(wl), C.byref(dw), C.byref(nw), C.byref(ne), C.byref(nd), ms, err, len(err))
    if rc != 0:
        raise MatchtigsError(rc, err.value.decode())
    try:
        def take(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt).copy()
        edges, limits = take(we, ne.value, np.uint32), take(wl, nw.value, np.uint64)
        dummy_w = take(dw, nd.value, np.uint32)
    finally:
        for p in (we, wl, dw):
            l.mtg_host_free(p)
    walks = np.split(edges, limits[:-1].astype(np.int64)) if nw.value else []
    return walks, dummy_w, list(ms)


class Graph:
    """An edge-centric bidirected unitig graph resident on one GPU (the reference's ``CliGraph``)."""

    def __init__(self, ctx: Context, k: int, unitigs: int):
        self.ctx, self.k, self.unitigs = ctx, k, unitigs

    def node_count(self) -> int:
        return self.ctx.graph_info()["nodes"]

    def edge_count(self) -> int:
        return self.ctx.graph_info()["edges"]


_DEVICE_PARSE_LIMIT = (1 << 35) - 64  # the device-side record parser tags 32-byte chunks with 30-bit indices


def read_bigraph_from_fasta_as_edge_centric(text: bytes, k: int, ctx: Context | None = None, device_parse: bool = True) -> Graph:
    """``--fa-in``: nodes are the distinct (k-1)-mers at unitig ends, numbered in first-seen order.

    Records are split on the device (``mtg_build_graph_from_text``); ``device_parse=False`` (or a file of 32 GiB and
    more) uses the host reader + ``mtg_build_graph_from_sequences`` instead -- same graph either way."""
    ctx = ctx or Context()
    if device_parse and len(text) < _DEVICE_PARSE_LIMIT:
        ctx.build_graph_from_text(text, k, bcalm=False)
    else:
        u = Unitigs(text, bcalm=False)
        ctx.build_graph_from_sequences(u.seq, u.offsets, k)
    return Graph(ctx, k, ctx.graph_info()["unitigs"])


def read_bigraph_from_bcalm2_as_edge_centric(text: bytes, k: int, ctx: Context | None = None, device_parse: bool = True) -> Graph:
    """``--bcalm-in``: topology from the ``L:`` links (union-find numbering of ``src/clib.rs``)."""
    ctx = ctx or Context()
    if device_parse and len(text) < _DEVICE_PARSE_LIMIT:
        ctx.build_graph_from_text(text, k, bcalm=True)
    else:
        u = Unitigs(text, bcalm=True)
        lens = np.diff(u.offsets.astype(np.int64))
        if u.count and int(lens.min()) < k:
            raise MatchtigsError(-3, "sequence shorter than k")
        weights = (lens + 1 - k).astype(np.uint64)
        ctx.build_graph_from_links(weights, u.link_a, u.strand_a, u.link_b, u.strand_b, k, u.seq, u.offsets)
    return Graph(ctx, k, ctx.graph_info()["unitigs"])


@dataclass
class GreedytigAlgorithmConfiguration:
    """Mirror of ``GreedytigAlgorithmConfiguration`` (``greedytigs/mod.rs:40-73``).

    ``threads``, the heap / node-weight-array / staged-parallelism knobs are accepted for flag
    compatibility; on the GPU path they are result-neutral (they never change outputs in the
    reference either).  ``candidate_cap`` is the only knob of the new path: list depth per source
    before a re-query phase becomes necessary (results are identical for every value)."""
    k: int
    threads: int = 1
    staged_parallelism_divisor: float | None = None
    resource_limit_factor: int = 1
    heap_type: str = "StdBinaryHeap"
    node_weight_array_type: str = "HashbrownHashMap"
    performance_data_type: str = "None"
    candidate_cap: int = 8


class GreedytigAlgorithm:
    """``TigAlgorithm`` implementation for greedy matchtigs (``greedytigs/mod.rs:75-90``)."""

    @staticmethod
    def compute_tigs(graph: Graph, configuration: GreedytigAlgorithmConfiguration) -> list[np.ndarray]:
        ctx = graph.ctx
        ctx.dijkstra_candidates(configuration.candidate_cap, 0, 1)
        ctx.greedy_match(export=False)
        ctx.finish_walks()
        return ctx.walks()


def write_walks_gfa(graph: Graph) -> bytes:
    return graph.ctx.assemble_tigs("gfa")


def write_walks_fasta(graph: Graph) -> bytes:
    return graph.ctx.assemble_tigs("fasta")


def write_duplication_bitvector(graph: Graph) -> bytes:
    return graph.ctx.dup_bitvector()
