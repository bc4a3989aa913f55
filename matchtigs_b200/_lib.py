"""ctypes binding of ``libmatchtigs_b200.so`` (the C ABI of ``include/matchtigs_b200.h``).

There is deliberately no fallback: if the shared library is missing or no CUDA device is usable
the import of the library / creation of a context raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# MTG_LIB_PATH: diagnostic override used to A/B differently compiled builds of the same library
SO_PATH = Path(os.environ.get("MTG_LIB_PATH") or Path(__file__).resolve().parent / "libmatchtigs_b200.so")

MTG_OK = 0
STATUS_NAMES = {0: "MTG_OK", -1: "MTG_ERR_INVALID", -2: "MTG_ERR_CUDA", -3: "MTG_ERR_INPUT", -4: "MTG_ERR_INTERNAL",
                -5: "MTG_ERR_UNSUPPORTED"}

# every symbol include/matchtigs_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "mtg_ctx_create", "mtg_ctx_destroy", "mtg_last_error", "mtg_ctx_set_option", "mtg_ctx_stream", "mtg_ctx_kernel_launches",
    "mtg_build_graph_from_sequences", "mtg_build_graph_from_links", "mtg_build_graph_from_text", "mtg_graph_get_info", "mtg_graph_export",
    "mtg_dijkstra_candidates", "mtg_candidates_local", "mtg_candidates_export",
    "mtg_greedy_match", "mtg_triples_export",
    "mtg_finish_walks", "mtg_walks_export", "mtg_walks_export_capi", "mtg_host_tail", "mtg_host_free",
    "mtg_dup_bitvector", "mtg_assemble_tigs", "mtg_dup_bitvector_view", "mtg_assemble_tigs_view",
    "mtg_dup_bitvector_range_view", "mtg_assemble_tigs_range_view", "mtg_walk_count",
    "mtg_comm_get_unique_id", "mtg_comm_init", "mtg_comm_destroy", "mtg_allgather_candidates", "mtg_text_slice_bytes",
    "mtg_build_graph_from_text_slices", "mtg_broadcast_walks",
    "mtg_compute_greedytigs_from_sequences", "mtg_get_search_stats", "mtg_get_diagnostics",
    "mtg_unitigs_parse", "mtg_unitigs_free", "mtg_unitigs_view",
    "matchtigs_initialise", "matchtigs_initialise_graph", "matchtigs_merge_nodes", "matchtigs_build_graph",
    "matchtigs_compute_tigs",
]


class GraphInfo(C.Structure):
    _fields_ = [("unitigs", C.c_uint64), ("nodes", C.c_uint64), ("edges", C.c_uint64), ("short_edges", C.c_uint64),
                ("sources", C.c_uint64), ("targets", C.c_uint64), ("self_mirrors_unbalanced", C.c_uint64),
                ("k", C.c_uint32)]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SearchStats(C.Structure):
    _fields_ = [("sources_searched", C.c_uint64), ("settled_nodes", C.c_uint64), ("relaxed_edges", C.c_uint64),
                ("candidates", C.c_uint64), ("truncated_sources", C.c_uint64), ("overflow_sources", C.c_uint64),
                ("match_rounds", C.c_uint64), ("requery_phases", C.c_uint64), ("matched", C.c_uint64),
                ("dijkstra_ms", C.c_float), ("match_ms", C.c_float), ("dijkstra_kernel_ms", C.c_float),
                ("match_kernel_ms", C.c_float), ("labelled_nodes", C.c_uint64), ("max_labelled_nodes", C.c_uint64),
                ("max_open_nodes", C.c_uint64), ("preextended_sources", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: (float(getattr(self, n)) if t is C.c_float else int(getattr(self, n))) for n, t in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Loads the product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(matchtigs_b200 has no CPU fallback)")
    l = C.CDLL(str(SO_PATH))
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    l.mtg_ctx_create.argtypes = [C.POINTER(vp), i32]
    l.mtg_ctx_destroy.argtypes = [vp]
    l.mtg_ctx_destroy.restype = None
    l.mtg_ctx_set_option.argtypes = [vp, C.c_char_p, i32]
    l.mtg_last_error.argtypes = [vp]
    l.mtg_last_error.restype = C.c_char_p
    l.mtg_ctx_stream.argtypes = [vp]
    l.mtg_ctx_stream.restype = vp
    l.mtg_ctx_kernel_launches.argtypes = [vp]
    l.mtg_ctx_kernel_launches.restype = u64
    l.mtg_build_graph_from_sequences.argtypes = [vp, vp, vp, u64, u32, i32]
    l.mtg_build_graph_from_links.argtypes = [vp, u64, vp, u64, vp, vp, vp, vp, u32, vp, vp]
    l.mtg_build_graph_from_text.argtypes = [vp, vp, u64, i32, u32, i32]
    l.mtg_graph_get_info.argtypes = [vp, C.POINTER(GraphInfo)]
    l.mtg_graph_export.argtypes = [vp, vp, vp, vp, vp, vp]
    l.mtg_dijkstra_candidates.argtypes = [vp, u32, u32, u32]
    l.mtg_candidates_local.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(u32)]
    l.mtg_candidates_export.argtypes = [vp, vp, vp, vp]
    l.mtg_greedy_match.argtypes = [vp, vp, vp, u32, C.POINTER(u64)]
    l.mtg_triples_export.argtypes = [vp, vp]
    l.mtg_finish_walks.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    l.mtg_walks_export.argtypes = [vp, vp, vp]
    l.mtg_walks_export_capi.argtypes = [vp, vp, vp, vp]
    l.mtg_host_tail.argtypes = [u32, u64, u64, vp, vp, vp, vp, vp, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
    l.mtg_host_free.argtypes = [vp]
    l.mtg_host_free.restype = None
    l.mtg_dup_bitvector.argtypes = [vp, vp, u64, C.POINTER(u64)]
    l.mtg_assemble_tigs.argtypes = [vp, i32, vp, u64, C.POINTER(u64)]
    l.mtg_dup_bitvector_view.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    l.mtg_assemble_tigs_view.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
    l.mtg_dup_bitvector_range_view.argtypes = [vp, u64, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    l.mtg_assemble_tigs_range_view.argtypes = [vp, i32, u64, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    l.mtg_walk_count.argtypes = [vp, C.POINTER(u64)]
    l.mtg_comm_get_unique_id.argtypes = [vp]
    l.mtg_comm_init.argtypes = [vp, vp, i32, i32]
    l.mtg_comm_destroy.argtypes = [vp]
    l.mtg_allgather_candidates.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    l.mtg_text_slice_bytes.argtypes = [u64, u32]
    l.mtg_text_slice_bytes.restype = u64
    l.mtg_build_graph_from_text_slices.argtypes = [vp, vp, u64, u64, i32, u32]
    l.mtg_broadcast_walks.argtypes = [vp, i32]
    l.mtg_compute_greedytigs_from_sequences.argtypes = [vp, vp, vp, u64, u32, u32]
    l.mtg_get_search_stats.argtypes = [vp, C.POINTER(SearchStats)]
    l.mtg_get_diagnostics.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    l.mtg_unitigs_parse.argtypes = [C.c_char_p, C.c_size_t, i32, C.POINTER(vp), C.c_char_p, C.c_size_t]
    l.mtg_unitigs_free.argtypes = [vp]
    l.mtg_unitigs_free.restype = None
    l.mtg_unitigs_view.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(vp),
                                   C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    l.matchtigs_initialise.restype = None
    l.matchtigs_initialise_graph.argtypes = [C.c_size_t]
    l.matchtigs_initialise_graph.restype = vp
    l.matchtigs_merge_nodes.argtypes = [vp, C.c_size_t, C.c_bool, C.c_size_t, C.c_bool]
    l.matchtigs_merge_nodes.restype = None
    l.matchtigs_build_graph.argtypes = [vp, vp]
    l.matchtigs_build_graph.restype = None
    l.matchtigs_compute_tigs.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_char_p, C.c_char_p, vp, vp, vp]
    l.matchtigs_compute_tigs.restype = C.c_size_t
    _lib = l
    return l
