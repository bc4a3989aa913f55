// tail_prep.cu -- device-side preparation of the host Euler walk.
//
// The walk itself (host_tail.cpp) is inherently sequential, but everything it consumes is not:
//   * the nodes that are still unbalanced after the matching (the input of
//     make_graph_eulerian_with_breaking_edges, src/implementation/mod.rs:408-427) are a compaction of the
//     final multiplicity array -- three flag scans;
//   * the adjacency the walk iterates (petgraph order: newest edge first) is a stable radix sort of
//     (from-node, edge id) over all original and dummy edges, written straight into the one-line node records
//     (with their prefetch hints) the walk uses, DMA'd to the host.
// Only the pairing loop of eulerise (sequential by definition) runs on the host in between.
#include <algorithm>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;

__global__ void __launch_bounds__(TB) leftover_flags(const i32* __restrict__ mult, const u32* __restrict__ mirror, u64 N,
                                                     u32* __restrict__ f_self, u32* __restrict__ f_out, u32* __restrict__ f_in) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const i32 m = mult[v];
    const bool self = mirror[v] == (u32)v;
    f_self[v] = (self && m != 0) ? 1u : 0u;  // odd out-degree (find_non_eulerian_binodes_with_differences pushes (v, 0))
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
                   const u32* __restrict__ mirror, u32* __restrict__ key, u32* __restrict__ val) {
    u64 q = (u64)blockIdx.x * TB + threadIdx.x;
    if (q >= E) return;
    const u32 e = (u32)(E - 1 - q);
    key[q] = edge_from_of(e, E0, edge_from, pair_out, pair_in, mirror);
    val[q] = e;
}

__global__ void __launch_bounds__(TB) ext_counts(const u32* __restrict__ deg, u64 N, u32* __restrict__ cnt) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v < N) cnt[v] = deg[v] > ROW_INLINE ? deg[v] : 0u;
}

__global__ void __launch_bounds__(TB)
    fill_rows(const u32* __restrict__ key, const u32* __restrict__ val, u64 E, u64 E0, const u32* __restrict__ edge_to,
              const u32* __restrict__ pair_out, const u32* __restrict__ pair_in, const u32* __restrict__ mirror,
              const u32* __restrict__ deg, const u32* __restrict__ row_ptr, const u32* __restrict__ ext_off, NodeRow* __restrict__ rows,
              AdjEntry* __restrict__ ext) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= E) return;
    const u32 v = key[i], e = val[i];
    const u32 rank = (u32)i - row_ptr[v];
    const AdjEntry a{e, edge_to_of(e, E0, edge_to, pair_out, pair_in, mirror)};
    if (deg[v] <= ROW_INLINE) rows[v].inl[rank] = a;
    else ext[ext_off[v] + rank] = a;
}

__global__ void __launch_bounds__(TB) row_headers(const u32* __restrict__ deg, const u32* __restrict__ ext_off, u64 N, NodeRow* __restrict__ rows) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v >= N) return;
    const u32 d = deg[v];
    if (d <= ROW_INLINE) {
        rows[v].cur = 0;
        rows[v].end = d;
    } else {
        rows[v].cur = ext_off[v];
        rows[v].end = (ext_off[v] + d) | ROW_EXT;
    }
}

// rows and cursors are final: prefetch hints for the host walk (two random gathers per entry, cheap here, a cache miss
// each on the host)
__global__ void __launch_bounds__(TB) row_hints(NodeRow* rows, const AdjEntry* __restrict__ ext, u64 N) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    if (v < N) fill_row_hints(rows, ext, (u32)v);
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
    MTG_LAUNCH(ctx, leftover_flags, grid_for(N, TB), TB, 0, ctx->final_mult.p, ctx->mirror.p, N, f_self.p, f_out.p, f_in.p);
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

void tail_build_rows(mtg_ctx* ctx, const u32* breaking_pairs, u64 n_break, NodeRow* h_rows, PinnedBuf& ext_stage, u64* n_ext_out,
                     u64* n_pairs_out) {
    cudaStream_t s = ctx->stream;
    const u64 N = ctx->N, E0 = ctx->E, P = ctx->n_triples + n_break, E = E0 + 2 * P;
    MTG_REQUIRE(E < 0xFFFFFFF0ull, MTG_ERR_UNSUPPORTED, "more than 2^32 edges");
    *n_pairs_out = P;
    *n_ext_out = 0;
    if (N == 0) return;
    DBuf<u32> d_break, pair_out, pair_in, deg, key_a, key_b, val_a, val_b, row_ptr, ext_cnt, ext_off, total;
    DBuf<NodeRow> rows;
    DBuf<AdjEntry> ext;
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
    if (E) MTG_LAUNCH(ctx, edge_sort_keys, grid_for(E, TB), TB, 0, E, E0, ctx->edge_from.p, pair_out.p, pair_in.p, ctx->mirror.p, key_a.p, val_a.p);
    const int which = radix_sort_pairs_u32(ctx, key_a.p, key_b.p, val_a.p, val_b.p, E, bits_for(N));
    row_ptr.resize(N + 1, s);
    ext_cnt.resize(N, s);
    ext_off.resize(N, s);
    total.resize(2, s);
    exclusive_sum_u32(ctx, deg.p, row_ptr.p, N, total.p + 0);
    MTG_LAUNCH(ctx, ext_counts, grid_for(N, TB), TB, 0, deg.p, N, ext_cnt.p);
    exclusive_sum_u32(ctx, ext_cnt.p, ext_off.p, N, total.p + 1);
    u32 h_total[2];
    MTG_CUDA(cudaMemcpyAsync(h_total, total.p, sizeof(h_total), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    MTG_REQUIRE(h_total[0] == E, MTG_ERR_INTERNAL, "degree sum does not match the edge count");
    const u64 n_ext = h_total[1];
    MTG_REQUIRE(n_ext < ROW_EXT, MTG_ERR_UNSUPPORTED, "too many edges at high-degree nodes");
    rows.resize(N, s);
    ext.resize(std::max<u64>(n_ext, 1), s);
    if (E) MTG_LAUNCH(ctx, fill_rows, grid_for(E, TB), TB, 0, which ? key_b.p : key_a.p, which ? val_b.p : val_a.p, E, E0, ctx->edge_to.p,
                      pair_out.p, pair_in.p, ctx->mirror.p, deg.p, row_ptr.p, ext_off.p, rows.p, ext.p);
    MTG_LAUNCH(ctx, row_headers, grid_for(N, TB), TB, 0, deg.p, ext_off.p, N, rows.p);
    MTG_LAUNCH(ctx, row_hints, grid_for(N, TB), TB, 0, rows.p, ext.p, N);
    MTG_CUDA(cudaMemcpyAsync(h_rows, rows.p, N * sizeof(NodeRow), cudaMemcpyDeviceToHost, s));
    AdjEntry* h_ext = ext_stage.as<AdjEntry>(std::max<u64>(n_ext, 1));
    if (n_ext) MTG_CUDA(cudaMemcpyAsync(h_ext, ext.p, n_ext * sizeof(AdjEntry), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    *n_ext_out = n_ext;
    for (DBuf<u32>* b : {&d_break, &pair_out, &pair_in, &deg, &key_a, &key_b, &val_a, &val_b, &row_ptr, &ext_cnt, &ext_off, &total}) b->release(s);
    rows.release(s);
    ext.release(s);
}

}  // namespace mtg
