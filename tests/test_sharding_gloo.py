"""world_size-2/3 CPU test (gloo) of the multi-GPU host logic: round-robin source sharding, equal-sized slices,
all-gather layout [rank][local][cap] and the slot formula the matching kernel uses to find source i."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _worker(rank, world, port, text, k, cap, ret):
    sys.path.insert(0, str(ROOT))
    import oracle
    from matchtigs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.Oracle().load_bcalm(text, k).run()
    S = o.num("sources")
    nodes, dists, lens = o.candidates(cap)
    padded = sharding.padded_slice(S, world)
    nloc = sharding.local_count(S, rank, world)
    rec = torch.zeros((padded, cap), dtype=torch.int64)
    meta = torch.zeros((padded,), dtype=torch.int32)
    for l in range(nloc):
        i = sharding.global_index(l, rank, world)
        c = min(int(lens[i]), cap)
        rec[l, :c] = torch.from_numpy((nodes[i, :c].astype(np.int64) | (dists[i, :c].astype(np.int64) << 32)))
        meta[l] = c
    rec_all = torch.empty((world * padded, cap), dtype=torch.int64)
    meta_all = torch.empty((world * padded,), dtype=torch.int32)
    dist.all_gather_into_tensor(rec_all, rec)
    dist.all_gather_into_tensor(meta_all, meta)
    ok = True
    flat_rec, flat_meta = rec_all.reshape(-1, cap), meta_all.reshape(-1)
    for i in range(S):
        slot = sharding.gathered_slot(i, S, world)
        c = min(int(lens[i]), cap)
        ok &= int(flat_meta[slot]) == c
        got = flat_rec[slot, :c].numpy()
        ok &= np.array_equal(got & 0xFFFFFFFF, nodes[i, :c].astype(np.int64)) and np.array_equal(got >> 32, dists[i, :c].astype(np.int64))
    ok &= sum(sharding.local_count(S, r, world) for r in range(world)) == S
    # sliced upload (mtg_build_graph_from_text_slices): equal-sized, 16-byte aligned slices, all-gathered back into the file
    lo, hi = sharding.text_slice(len(text), rank, world)
    sl = sharding.text_slice(len(text), 0, world)[1]
    ok &= sl % 16 == 0 and sl * world >= len(text) and hi - lo <= sl
    part = torch.zeros(sl, dtype=torch.uint8)
    part[:hi - lo] = torch.frombuffer(bytearray(text[lo:hi]), dtype=torch.uint8)
    whole = torch.empty(sl * world, dtype=torch.uint8)
    dist.all_gather_into_tensor(whole, part)
    ok &= whole[:len(text)].numpy().tobytes() == text
    # sharded emission: the walk shares partition [0, T)
    T = o.num("walks")
    shares = [sharding.walk_share(T, r, world) for r in range(world)]
    ok &= shares[0][0] == 0 and shares[-1][1] == T and all(shares[r][1] == shares[r + 1][0] for r in range(world - 1))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_layout(world):
    import tools
    anc = tools.genome(6_000, 21, families=2, copies=3, min_len=40, max_len=150, divergence=0.02)
    text, _, _ = tools.unitigs(tools.pangenome(anc, 6, 3, snp_site_rate=0.04), 15)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, text, 15, 8, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1
