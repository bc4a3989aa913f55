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


def all_gather_candidates(ctx, dist, torch, world: int):
    """NCCL all-gather of this rank's candidate slice; returns (records[R*padded,cap], meta[R*padded]) device tensors."""
    from .api import device_tensor
    prec, pmeta, _n_local, cap = ctx.candidates_local()
    padded = padded_slice(ctx.graph_info()["sources"], world)
    rec_all = torch.empty((world * padded, cap), dtype=torch.int64, device="cuda")
    meta_all = torch.empty((world * padded,), dtype=torch.int32, device="cuda")
    dist.all_gather_into_tensor(rec_all, device_tensor(prec, (padded, cap), "<i8"))
    dist.all_gather_into_tensor(meta_all, device_tensor(pmeta, (padded,), "<i4"))
    torch.cuda.synchronize()
    return rec_all, meta_all
