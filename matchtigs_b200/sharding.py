"""Source sharding for the multi-GPU path (host-side arithmetic shared by bench.py and the tests).

The CSR graph is replicated on every rank; Dijkstra sources are dealt round-robin
(source i belongs to rank i % R, local index i // R) because per-source cost is heavy-tailed.
Every rank owns a slice of ``padded = ceil(S / R)`` lists, so one equal-sized all-gather produces
the layout ``[rank][local][cap]`` that ``mtg_greedy_match`` consumes (``init_lists`` in match.cu:
slot(i) = (i % R) * padded + i // R)."""
from __future__ import annotations


def padded_slice(n_sources: int, world: int) -> int:
    return max((n_sources + world - 1) // world, 1)


def local_count(n_sources: int, rank: int, world: int) -> int:
    return (n_sources - rank + world - 1) // world if n_sources > rank else 0


def global_index(local: int, rank: int, world: int) -> int:
    return local * world + rank


def gathered_slot(i: int, n_sources: int, world: int) -> int:
    return (i % world) * padded_slice(n_sources, world) + i // world


def text_slice(total_len: int, rank: int, world: int) -> tuple[int, int]:
    """Byte range [lo, hi) of the input file that `rank` uploads (mtg_text_slice_bytes: ceil(len / world), 16-byte aligned)."""
    sl = (total_len + world - 1) // world
    sl = (sl + 15) // 16 * 16
    lo = min(rank * sl, total_len)
    return lo, min(lo + sl, total_len)


def walk_share(n_walks: int, rank: int, world: int) -> tuple[int, int]:
    """Walks [lo, hi) whose output bytes `rank` assembles and downloads (sharded emission)."""
    return n_walks * rank // world, n_walks * (rank + 1) // world
