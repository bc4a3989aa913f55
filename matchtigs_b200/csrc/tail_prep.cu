// tail_prep.cu -- device-side preparation of the host Euler walk.
//
// The walk itself (host_tail.cpp) is inherently sequential, but everything it consumes is not:
//   * the nodes that are still unbalanced after the matching (the input of
//     make_graph_eulerian_with_breaking_edges, src/implementation/mod.rs:408-427) are a compaction of the
//     final multiplicity array -- three flag scans;
//   * the adjacency the walk iterates (petgraph order: newest edge first) is a stable radix sort of
//     (from-node, edge id) over all original and dummy edges, written straight into the per-edge walk records
//     (WalkRec, mtg_internal.cuh: where the edge leads, its mirror's slot, and the handles of the nodes up to
//     WALK_DEPTH steps ahead) the host walk uses, DMA'd to the host;
//   * the Eulerian check (greedytigs/mod.rs:708-715) is a reduction over the final degrees.
// Only the pairing loop of eulerise (sequential by definition) runs on the host in between.
#include <algorithm>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;

__global__ void __launch_bounds__(TB) leftover_flags(const i32* __restrict__ mult, const u32* __restrict__ mirror, const u32* __restrict__ out_deg,
                                                     bool self_by_degree, u64 N, u32* __restrict__ f_self, u32* __restrict__ f_out,
                                                     u32* __restrict__ f_in) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const i32 m = mult[v];
    const bool self = mirror[v] == (u32)v;
    // odd out-degree (find_non_eulerian_binodes_with_differences pushes (v, 0)).  The multiplicity of a self-mirror is its
    // degree parity, kept up to date by the matching -- unless assumption P2 is flipped and self-mirrors never took part
    // in it: then the parity of the original degree is the answer.
    const bool odd = self_by_degree ? (out_deg[v] & 1u) != 0 : m != 0;
    f_self[v] = (self && odd) ? 1u : 0u;
    f_out[v] = (!self && m < 0) ? 1u : 0u;
    f_in[v] = (!self && m > 0) ? 1u : 0u;
}

__global__ void __launch_bounds__(TB)
    leftover_compact(const i32* __restrict__ mult, const u32* __restrict__ mirror, u64 N, const u32* __restrict__ f_self,
                     const u32* __restrict__ f_out, const u32* __restrict__ f_in, const u32* __restrict__ p_self,
                     const u32* __restrict__ p_out, const u32* __restrict__ p_in, u32* __restrict__ self_nodes,
                     u32* __restrict__ out_nodes, i32* __restrict__ out_diff, u32* __restrict__ out_partner,
                     u32* __restrict__ in_nodes, i32* __restrict__ in_diff, u32* __restrict__ in_partner) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    if (f_self[v]) self_nodes[p_self[v]] = (u32)v;
    if (f_out[v]) {
        const u32 o = p_out[v];
        out_nodes[o] = (u32)v;
        out_diff[o] = mult[v];
        out_partner[o] = p_in[mirror[v]];  // the mirror of an out-node is an in-node with the opposite difference
    }
    if (f_in[v]) {
        const u32 o = p_in[v];
        in_nodes[o] = (u32)v;
        in_diff[o] = mult[v];
        in_partner[o] = p_out[mirror[v]];
    }
}

// all dummy pairs in insertion order: matching triples (greedytigs/mod.rs:678-689), then breaking pairs
__global__ void __launch_bounds__(TB)
    gather_pairs(const u32* __restrict__ triples, u64 n_triples, const u32* __restrict__ breaking, u64 n_break,
                 const u32* __restrict__ mirror, u32* __restrict__ pair_out, u32* __restrict__ pair_in, u32* __restrict__ deg) {
    u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_triples + n_break) return;
    u32 o, i;
    if (j < n_triples) {
        o = triples[3 * j];
        i = triples[3 * j + 1];
    } else {
        o = breaking[2 * (j - n_triples)];
        i = breaking[2 * (j - n_triples) + 1];
    }
    pair_out[j] = o;
    pair_in[j] = i;
    atomicAdd(&deg[o], 1u);          // edge E0+2j:   out -> in
    atomicAdd(&deg[mirror[i]], 1u);  // edge E0+2j+1: mirror(in) -> mirror(out)
}

__device__ __forceinline__ u32 edge_from_of(u32 e, u64 E0, const u32* edge_from, const u32* pair_out, const u32* pair_in,
                                            const u32* mirror) {
    if (e < E0) return edge_from[e];
    const u32 j = (u32)((e - E0) >> 1);
    return ((e - E0) & 1) ? mirror[pair_in[j]] : pair_out[j];
}
__device__ __forceinline__ u32 edge_to_of(u32 e, u64 E0, const u32* edge_to, const u32* pair_out, const u32* pair_in,
                                          const u32* mirror) {
    if (e < E0) return edge_to[e];
    const u32 j = (u32)((e - E0) >> 1);
    return ((e - E0) & 1) ? mirror[pair_out[j]] : pair_in[j];
}

// keys in DESCENDING edge id order: a stable sort by from-node then leaves every row newest edge first
__global__ void __launch_bounds__(TB)
    edge_sort_keys(u64 E, u64 E0, const u32* __restrict__ edge_from, const u32* __restrict__ pair_out, const u32* __restrict__ pair_in,
                   const u32* __restrict__ mirror, bool oldest_first, u32* __restrict__ key, u32* __restrict__ val) {
    u64 q = (u64)blockIdx.x * TB + threadIdx.x;
    if (q >= E) return;
    const u32 e = oldest_first ? (u32)q : (u32)(E - 1 - q);  // assumption P3: newest edge first
    key[q] = edge_from_of(e, E0, edge_from, pair_out, pair_in, mirror);
    val[q] = e;
}

// Slots are handed out binode by binode: a node and its mirror sit next to each other, because taking an edge into node c
// marks its mirror edge used, and that one leaves mirror(c) -- so both marks and the look at c's used bits hit the same
// cache line of the bitset, whatever the node numbering (first-seen ids pair mirrors up, union-find ranks do not).
// cap[v] = slots of v and of its mirror, counted at the smaller id of the two (0 at the larger).
__global__ void __launch_bounds__(TB) slot_caps(const u32* __restrict__ deg, const u32* __restrict__ mirror, u64 N, u32* __restrict__ cap) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const u32 m = mirror[v];
    cap[v] = m < v ? 0u : walk_cap(deg[v]) + (m > v ? walk_cap(deg[m]) : 0u);
}

// handles, headers of big nodes and the initial used-slot bitset (padding slots and header pairs are never handed out)
__global__ void __launch_bounds__(TB)
    node_handles(const u32* __restrict__ deg, const u32* __restrict__ mirror, const u32* __restrict__ base, u64 N, u32* __restrict__ handle,
                 WalkRec* __restrict__ recs, u32* __restrict__ used0) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const u32 m = mirror[v];
    const u32 d = deg[v], b = m < v ? base[m] + walk_cap(deg[m]) : base[v], h = walk_handle(b, d), cap = walk_cap(d);
    handle[v] = h;
    const u32 first = walk_first_slot(h);
    if (d > 4) {
        recs[b].to = d;
        atomicOr(&used0[b >> 5], 3u << (b & 31));  // b is even: both header slots share a word
    }
    for (u32 sl = first + d; sl < b + cap; sl++) atomicOr(&used0[sl >> 5], 1u << (sl & 31));
}

// sorted position i holds edge val[i] of node key[i]; its rank inside the node is its iteration position
__global__ void __launch_bounds__(TB)
    place_slots(const u32* __restrict__ key, const u32* __restrict__ val, u64 E, u64 E0, const u32* __restrict__ edge_to,
                const u32* __restrict__ pair_out, const u32* __restrict__ pair_in, const u32* __restrict__ mirror,
                const u32* __restrict__ row_ptr, const u32* __restrict__ handle, WalkRec* __restrict__ recs, u32* __restrict__ slot_edge,
                u32* __restrict__ slot_of_edge, u32* __restrict__ slot_to) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= E) return;
    const u32 v = key[i], e = val[i];
    const u32 sl = walk_first_slot(handle[v]) + ((u32)i - row_ptr[v]);
    const u32 t = edge_to_of(e, E0, edge_to, pair_out, pair_in, mirror);
    slot_edge[sl] = e;
    slot_of_edge[e] = sl;
    slot_to[sl] = t;
    recs[sl].to = handle[t];
}

__global__ void __launch_bounds__(TB)
    slot_mirrors(const u32* __restrict__ slot_edge, const u32* __restrict__ slot_of_edge, u64 n_slots, u64 E0, u64 n_matching,
                 const u32* __restrict__ triples, u32 k, WalkRec* __restrict__ recs) {
    u64 sl = (u64)blockIdx.x * TB + threadIdx.x;
    if (sl >= n_slots) return;
    const u32 e = slot_edge[sl];
    if (e == NONE32) return;
    u32 m = slot_of_edge[e ^ 1u];
    if (e >= E0) {
        const u64 j = (e - E0) >> 1;  // matching dummies weigh their distance, breaking dummies k
        const u32 w = j < n_matching ? triples[3 * j + 2] : k;
        m |= SLOT_DUMMY | (w >= k ? SLOT_BREAK : 0u);
    }
    recs[sl].mslot = m;
}

__global__ void __launch_bounds__(TB)
    slot_hints(const u32* __restrict__ slot_edge, const u32* __restrict__ slot_to, const u32* __restrict__ deg, u64 n_slots, u32 level,
               WalkRec* recs) {
    u64 sl = (u64)blockIdx.x * TB + threadIdx.x;
    if (sl >= n_slots || slot_edge[sl] == NONE32) return;
    walk_fill_hints(recs, (u32)sl, deg[slot_to[sl]], level);
}

__global__ void __launch_bounds__(TB)
    start_handles(const u32* __restrict__ edge_from, const u32* __restrict__ handle, u64 E0, u32* __restrict__ out) {
    u64 e = (u64)blockIdx.x * TB + threadIdx.x;
    if (e < E0) out[e] = handle[edge_from[e]];
}

// E. (greedytigs/mod.rs:708-715, bigraph decomposes_into_eulerian_bicycles): out-degree == in-degree for every node
// (in-degree(v) == out-degree(mirror(v)) by the mirror property), even out-degree at self-mirrors.
__global__ void __launch_bounds__(TB) eulerian_check(const u32* __restrict__ deg, const u32* __restrict__ mirror, u64 N, u32* __restrict__ bad) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const u32 m = mirror[v];
    const bool ok = m == (u32)v ? !(deg[v] & 1u) : deg[v] == deg[m];
    if (!ok) atomicAdd(bad, 1u);
}

int bits_for(u64 n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) b++;
    return b;
}

}  // namespace

void tail_leftover(mtg_ctx* ctx, TailLeftover& lo) {
    cudaStream_t s = ctx->stream;
    const u64 N = ctx->N;
    lo = TailLeftover();
    if (N == 0) return;
    DBuf<u32> f_self, f_out, f_in, p_self, p_out, p_in, totals;
    for (DBuf<u32>* b : {&f_self, &f_out, &f_in, &p_self, &p_out, &p_in}) b->resize(N, s);
    totals.resize(4, s);
    MTG_LAUNCH(ctx, leftover_flags, grid_for(N, TB), TB, 0, ctx->final_mult.p, ctx->mirror.p, ctx->out_deg.p, ctx->opt.p2_self_mirror_zero != 0, N,
               f_self.p, f_out.p, f_in.p);
    exclusive_sum_u32(ctx, f_self.p, p_self.p, N, totals.p + 0);
    exclusive_sum_u32(ctx, f_out.p, p_out.p, N, totals.p + 1);
    exclusive_sum_u32(ctx, f_in.p, p_in.p, N, totals.p + 2);
    u32 h_tot[3];
    MTG_CUDA(cudaMemcpyAsync(h_tot, totals.p, sizeof(h_tot), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    const u32 ns = h_tot[0], no = h_tot[1], ni = h_tot[2];
    MTG_REQUIRE(no == ni, MTG_ERR_INTERNAL, "unbalanced out- and in-node counts differ");
    DBuf<u32> d_self, d_on, d_op, d_in, d_ip;
    DBuf<i32> d_od, d_id;
    d_self.resize(ns, s);
    d_on.resize(no, s);
    d_op.resize(no, s);
    d_od.resize(no, s);
    d_in.resize(ni, s);
    d_ip.resize(ni, s);
    d_id.resize(ni, s);
    MTG_LAUNCH(ctx, leftover_compact, grid_for(N, TB), TB, 0, ctx->final_mult.p, ctx->mirror.p, N, f_self.p, f_out.p, f_in.p, p_self.p,
               p_out.p, p_in.p, d_self.p, d_on.p, d_od.p, d_op.p, d_in.p, d_id.p, d_ip.p);
    lo.self_nodes.resize(ns);
    lo.out_nodes.resize(no);
    lo.out_partner.resize(no);
    lo.out_diff.resize(no);
    lo.in_nodes.resize(ni);
    lo.in_partner.resize(ni);
    lo.in_diff.resize(ni);
    d_self.download(lo.self_nodes.data(), s);
    d_on.download(lo.out_nodes.data(), s);
    d_op.download(lo.out_partner.data(), s);
    d_od.download(lo.out_diff.data(), s);
    d_in.download(lo.in_nodes.data(), s);
    d_ip.download(lo.in_partner.data(), s);
    d_id.download(lo.in_diff.data(), s);
    MTG_CUDA(cudaStreamSynchronize(s));
    for (DBuf<u32>* b : {&f_self, &f_out, &f_in, &p_self, &p_out, &p_in, &totals, &d_self, &d_on, &d_op, &d_in, &d_ip}) b->release(s);
    d_od.release(s);
    d_id.release(s);
}

void tail_build_records(mtg_ctx* ctx, const u32* breaking_pairs, u64 n_break, TailRecords* out) {
    cudaStream_t s = ctx->stream;
    const u64 N = ctx->N, E0 = ctx->E, P = ctx->n_triples + n_break, E = E0 + 2 * P;
    MTG_REQUIRE(E < SLOT_MASK, MTG_ERR_UNSUPPORTED, "more than 2^30 edges");
    *out = TailRecords();
    out->n_pairs = P;
    if (N == 0) return;
    DBuf<u32> d_break, pair_out, pair_in, deg, key_a, key_b, val_a, val_b, row_ptr, cap, base, total, handle, slot_edge, slot_of_edge, slot_to,
        used0, from_handle;
    DBuf<WalkRec> recs;
    d_break.upload(breaking_pairs, 2 * n_break, s);
    pair_out.resize(P, s);
    pair_in.resize(P, s);
    deg.resize(N, s);
    MTG_CUDA(cudaMemcpyAsync(deg.p, ctx->out_deg.p, N * sizeof(u32), cudaMemcpyDeviceToDevice, s));
    if (P) MTG_LAUNCH(ctx, gather_pairs, grid_for(P, TB), TB, 0, ctx->triples.p, ctx->n_triples, d_break.p, n_break, ctx->mirror.p,
                      pair_out.p, pair_in.p, deg.p);
    key_a.resize(E, s);
    key_b.resize(E, s);
    val_a.resize(E, s);
    val_b.resize(E, s);
    if (E) MTG_LAUNCH(ctx, edge_sort_keys, grid_for(E, TB), TB, 0, E, E0, ctx->edge_from.p, pair_out.p, pair_in.p, ctx->mirror.p, ctx->opt.p3_oldest_first != 0, key_a.p, val_a.p);
    const int which = radix_sort_pairs_u32(ctx, key_a.p, key_b.p, val_a.p, val_b.p, E, bits_for(N));
    row_ptr.resize(N + 1, s);
    cap.resize(N, s);
    base.resize(N, s);
    total.resize(3, s);
    total.zero(s);
    exclusive_sum_u32(ctx, deg.p, row_ptr.p, N, total.p + 0);
    MTG_LAUNCH(ctx, slot_caps, grid_for(N, TB), TB, 0, deg.p, ctx->mirror.p, N, cap.p);
    exclusive_sum_u32(ctx, cap.p, base.p, N, total.p + 1);
    MTG_LAUNCH(ctx, eulerian_check, grid_for(N, TB), TB, 0, deg.p, ctx->mirror.p, N, total.p + 2);
    u32 h_total[3];
    MTG_CUDA(cudaMemcpyAsync(h_total, total.p, sizeof(h_total), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    MTG_REQUIRE(h_total[0] == E, MTG_ERR_INTERNAL, "degree sum does not match the edge count");
    MTG_REQUIRE(h_total[2] == 0, MTG_ERR_INTERNAL, "Failed to make the graph Eulerian.");  // greedytigs/mod.rs:708-715
    const u64 n_slots = h_total[1];
    MTG_REQUIRE(n_slots < SLOT_MASK, MTG_ERR_UNSUPPORTED, "more than 2^30 edge slots");
    const u64 used_words32 = 2 * (n_slots / 64 + 2);
    recs.resize(n_slots, s);
    handle.resize(N, s);
    slot_edge.resize(n_slots, s);
    slot_edge.fill_ff(s);
    slot_of_edge.resize(E, s);
    slot_to.resize(n_slots, s);
    used0.resize(used_words32, s);
    used0.zero(s);
    from_handle.resize(E0, s);
    MTG_LAUNCH(ctx, node_handles, grid_for(N, TB), TB, 0, deg.p, ctx->mirror.p, base.p, N, handle.p, recs.p, used0.p);
    if (E) {
        MTG_LAUNCH(ctx, place_slots, grid_for(E, TB), TB, 0, which ? key_b.p : key_a.p, which ? val_b.p : val_a.p, E, E0, ctx->edge_to.p, pair_out.p,
                   pair_in.p, ctx->mirror.p, row_ptr.p, handle.p, recs.p, slot_edge.p, slot_of_edge.p, slot_to.p);
        MTG_LAUNCH(ctx, slot_mirrors, grid_for(n_slots, TB), TB, 0, slot_edge.p, slot_of_edge.p, n_slots, E0, ctx->n_triples, ctx->triples.p, ctx->k,
                   recs.p);
        for (u32 level = 2; level <= WALK_DEPTH; level++)
            MTG_LAUNCH(ctx, slot_hints, grid_for(n_slots, TB), TB, 0, slot_edge.p, slot_to.p, deg.p, n_slots, level, recs.p);
        if (E0) MTG_LAUNCH(ctx, start_handles, grid_for(E0, TB), TB, 0, ctx->edge_from.p, handle.p, E0, from_handle.p);
    }
    // DMA into page-locked staging (full link speed); the records first, they are what the walk waits for
    out->n_slots = n_slots;
    out->recs = ctx->tail_stage[0].as<WalkRec>(n_slots + 1);
    out->used0 = ctx->tail_stage[1].as<u64>(used_words32 / 2 + 1);
    out->slot_of_edge = ctx->tail_stage[2].as<u32>(E0 + 1);
    out->handle = ctx->tail_stage[3].as<u32>(E0 + 1);  // handle of the from-node of every original edge
    out->slot_edge = ctx->tail_stage[4].as<u32>(n_slots + 1);
    // in pieces, each followed by an event: the host copies piece c into the walk's arena while piece c+1 is in flight
    out->chunk_slots = (n_slots + TAIL_DMA_CHUNKS - 1) / TAIL_DMA_CHUNKS;
    for (int c = 0; c < TAIL_DMA_CHUNKS; c++) {
        const u64 lo = std::min<u64>((u64)c * out->chunk_slots, n_slots), hi = std::min<u64>(lo + out->chunk_slots, n_slots);
        if (hi > lo) MTG_CUDA(cudaMemcpyAsync(out->recs + lo, recs.p + lo, (hi - lo) * sizeof(WalkRec), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaEventRecord(ctx->tail_events[c], s));
    }
    MTG_CUDA(cudaMemcpyAsync(out->used0, used0.p, used_words32 * sizeof(u32), cudaMemcpyDeviceToHost, s));
    if (E0) {
        MTG_CUDA(cudaMemcpyAsync(out->slot_of_edge, slot_of_edge.p, E0 * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(out->handle, from_handle.p, E0 * sizeof(u32), cudaMemcpyDeviceToHost, s));
    }
    MTG_CUDA(cudaMemcpyAsync(out->slot_edge, slot_edge.p, n_slots * sizeof(u32), cudaMemcpyDeviceToHost, s));
    // no synchronisation here: the caller overlaps its own work with the copies.  The device buffers are released in
    // stream order behind them.
}

}  // namespace mtg
