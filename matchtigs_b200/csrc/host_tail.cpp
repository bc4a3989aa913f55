// host_tail.cpp -- the sequential tail of compute_greedytigs that fixes the output order and
// therefore stays on the host (SURVEY.md section 2, rows marked H*):
//
//   C  dummy-edge insertion in result order            greedytigs/mod.rs:678-689
//   D  make_graph_eulerian_with_breaking_edges         src/implementation/mod.rs:392-649
//   E  Eulerian check                                  greedytigs/mod.rs:708-715
//   F  minimum bidirected Eulerian cycle decomposition greedytigs/mod.rs:722 (bigraph)
//   G  rotate to the heaviest dummy + break            greedytigs/mod.rs:726-789
//
// Same results as the reference, different machinery: flat arrays instead of petgraph +
// BTreeMaps, two monotone cursors instead of ordered-map lookups in D, and in F a circular
// doubly linked list (rotate == move the head) with a ring of not-yet-exhausted positions, which
// makes the decomposition linear instead of quadratic in the cycle length.
#include <algorithm>
#include <cstring>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

struct Tail {
    HostGraph& g;
    u32 k;
    Tail(HostGraph& g_, u32 k_) : g(g_), k(k_) {}

    u32 add_edge(u32 a, u32 b, u32 w, bool dummy) {
        u32 e = (u32)g.from.size();
        g.from.push_back(a);
        g.to.push_back(b);
        g.weight.push_back(w);
        g.dummy.push_back(dummy ? 1 : 0);
        g.next_out.push_back(g.head_out[a]);  // newest edge first (petgraph head insertion)
        g.head_out[a] = e;
        g.out_deg[a]++;
        g.in_deg[b]++;
        return e;
    }
    void add_dummy_pair(u32 out_node, u32 in_node, u32 w) {
        add_edge(out_node, in_node, w, true);
        add_edge(g.mirror[in_node], g.mirror[out_node], w, true);
    }
};

// D. Pairs the remaining imbalance with breaking edges of weight k.
void eulerise(Tail& t) {
    HostGraph& g = t.g;
    const u32 n = g.n_nodes;
    std::vector<i32> diff(n, 0);
    std::vector<u32> outs, ins, selfs;  // ascending node id
    for (u32 v = 0; v < n; v++) {
        if (g.mirror[v] == v) {
            if (g.out_deg[v] & 1) selfs.push_back(v);
        } else {
            i32 d = (i32)g.out_deg[v] - (i32)g.in_deg[v];
            diff[v] = d;
            if (d < 0) outs.push_back(v);
            else if (d > 0) ins.push_back(v);
        }
    }
    size_t ip = 0;            // first in-node still unbalanced
    size_t op = outs.size();  // one past the last out-node still unbalanced
    auto skip_ins = [&] {
        while (ip < ins.size() && diff[ins[ip]] <= 0) ip++;
    };
    auto skip_outs = [&] {
        while (op > 0 && diff[outs[op - 1]] >= 0) op--;
    };
    // self-mirrors pairwise in ascending order; an odd one out takes the smallest in-node (:481-524)
    for (size_t i = 0; i < selfs.size(); i += 2) {
        if (i + 1 < selfs.size()) {
            t.add_dummy_pair(selfs[i], selfs[i + 1], t.k);
        } else {
            skip_ins();
            MTG_REQUIRE(ip < ins.size(), MTG_ERR_INTERNAL,
                        "Have an uneven number of self-mirrors, but no other nodes with missing in edges.");
            u32 in_node = ins[ip];
            t.add_dummy_pair(selfs[i], in_node, t.k);
            diff[in_node] -= 1;
            diff[g.mirror[in_node]] += 1;
        }
    }
    // largest out-node with smallest in-node (:526-645)
    for (;;) {
        skip_outs();
        if (op == 0) break;
        const u32 out_node = outs[op - 1];
        skip_ins();
        MTG_REQUIRE(ip < ins.size(), MTG_ERR_INTERNAL, "No further in_nodes left");
        u32 in_node = ins[ip];
        // choose_in_node_from_iterator (:252-285): do not join a node to its own mirror unless it misses >= 2 edges
        if ((in_node == g.mirror[out_node] && diff[out_node] > -2) || in_node == out_node) {
            size_t q = ip + 1;
            while (q < ins.size() && diff[ins[q]] <= 0) q++;
            MTG_REQUIRE(q < ins.size(), MTG_ERR_INTERNAL, "No further in_nodes left");
            in_node = ins[q];
        }
        const u32 mirror_out_node = g.mirror[in_node], mirror_in_node = g.mirror[out_node];
        t.add_dummy_pair(out_node, in_node, t.k);
        diff[out_node] += 1;
        diff[in_node] -= 1;
        if (diff[mirror_out_node] < 0) diff[mirror_out_node] += 1;  // only while it is still listed (:609-627)
        if (diff[mirror_in_node] > 0) diff[mirror_in_node] -= 1;    // (:628-644)
    }
    skip_ins();
    MTG_REQUIRE(ip == ins.size(), MTG_ERR_INTERNAL, "eulerise: in-nodes left over");
}

// E. decomposes_into_eulerian_bicycles
bool is_eulerian(const HostGraph& g) {
    for (u32 v = 0; v < g.n_nodes; v++) {
        if (g.mirror[v] == v) {
            if (g.out_deg[v] & 1) return false;
        } else if (g.out_deg[v] != g.in_deg[v]) {
            return false;
        }
    }
    return true;
}

// F + G fused: every finished cycle is rotated and cut into walks right away.
struct WalkSink {
    const HostGraph& g;
    u32 k;
    std::vector<u32>& edges;
    std::vector<u64>& limits;
    u64 breaking = 0;
    void emit(const u32* b, const u32* e) {
        edges.insert(edges.end(), b, e);
        limits.push_back(edges.size());
    }
    // greedytigs/mod.rs:736-788
    void cycle(std::vector<u32>& cyc) {
        u32 longest_w = 0;
        size_t longest_i = 0;
        for (size_t i = 0; i < cyc.size(); i++) {
            u32 e = cyc[i];
            if (g.dummy[e] && g.weight[e] > longest_w) {  // strict: first heaviest wins
                longest_w = g.weight[e];
                longest_i = i;
            }
        }
        if (longest_w > 0) std::rotate(cyc.begin(), cyc.begin() + longest_i, cyc.end());
        size_t offset = 0;
        const u32* p = cyc.data();
        for (size_t i = 0; i < cyc.size(); i++) {
            u32 e = cyc[i];
            if (g.dummy[e] && (g.weight[e] >= k || i == 0)) {
                if (offset < i) emit(p + offset, p + i);
                offset = i + 1;
                breaking++;
            }
        }
        if (offset < cyc.size()) {
            if (!g.dummy[cyc.back()]) emit(p + offset, p + cyc.size());
            else if (offset < cyc.size() - 1) emit(p + offset, p + cyc.size() - 1);
        }
    }
};

void euler_walks(const HostGraph& g, WalkSink& sink, u64* n_cycles) {
    const u32 E = (u32)g.from.size();
    std::vector<u8> used(E, 0);
    std::vector<u32> cursor(g.head_out);
    std::vector<u32> nxt(E), prv(E), cnxt(E), cprv(E);
    std::vector<u32> cyc;
    auto first_unused = [&](u32 v) -> u32 {
        u32 e = cursor[v];
        while (e != NONE32 && used[e]) e = g.next_out[e];
        cursor[v] = e;
        return e;
    };
    u64 cycles = 0;
    for (u32 e0 = 0; e0 < E; e0++) {
        if (used[e0]) continue;
        u32 head = NONE32, chead = NONE32;
        size_t len = 0;
        // appending to the cycle vector == inserting before `head` in the ring; `chead` is the first
        // position at or after head whose from-node may still own an unused out-edge.
        auto append = [&](u32 e) {
            if (head == NONE32) {
                head = e;
                nxt[e] = prv[e] = e;
            } else {
                u32 tail = prv[head];
                nxt[tail] = e;
                prv[e] = tail;
                nxt[e] = head;
                prv[head] = e;
            }
            if (chead == NONE32) {
                chead = e;
                cnxt[e] = cprv[e] = e;
            } else {
                u32 ct = cprv[chead];
                cnxt[ct] = e;
                cprv[e] = ct;
                cnxt[e] = chead;
                cprv[chead] = e;
            }
            len++;
        };
        u32 start_edge = e0;
        while (start_edge != NONE32) {
            used[start_edge] = used[start_edge ^ 1u] = 1;
            append(start_edge);
            u32 cur = g.to[start_edge];
            for (u32 e; (e = first_unused(cur)) != NONE32;) {
                used[e] = used[e ^ 1u] = 1;
                append(e);
                cur = g.to[e];
            }
            start_edge = NONE32;
            while (chead != NONE32) {
                u32 found = first_unused(g.from[chead]);
                if (found != NONE32) {
                    start_edge = found;
                    head = chead;  // rotate_left(position of chead)
                    break;
                }
                if (cnxt[chead] == chead) {
                    chead = NONE32;
                } else {  // exhausted for good
                    u32 a = cprv[chead], b = cnxt[chead];
                    cnxt[a] = b;
                    cprv[b] = a;
                    chead = b;
                }
            }
        }
        cyc.resize(len);
        u32 e = head;
        for (size_t i = 0; i < len; i++) {
            cyc[i] = e;
            e = nxt[e];
        }
        sink.cycle(cyc);
        cycles++;
    }
    *n_cycles = cycles;
}

}  // namespace

void finish_walks(mtg_ctx* ctx) {
    MTG_REQUIRE(ctx->have_graph && ctx->have_triples, MTG_ERR_INVALID, "mtg_greedy_match has not run");
    cudaStream_t s = ctx->stream;
    const u64 U = ctx->U, N = ctx->N, E = ctx->E;
    HostGraph& g = ctx->hg;
    g = HostGraph();
    g.n_nodes = (u32)N;
    g.n_orig_edges = (u32)E;
    std::vector<u32> from(E), to(E), uw(U);
    g.mirror.resize(N);
    if (E) {
        MTG_CUDA(cudaMemcpyAsync(from.data(), ctx->edge_from.p, E * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(to.data(), ctx->edge_to.p, E * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(uw.data(), ctx->unitig_w.p, U * sizeof(u32), cudaMemcpyDeviceToHost, s));
    }
    if (N) MTG_CUDA(cudaMemcpyAsync(g.mirror.data(), ctx->mirror.p, N * sizeof(u32), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    const u64 n_trip = ctx->n_triples;
    const size_t reserve = E + 2 * n_trip + 64;
    g.from.reserve(reserve);
    g.to.reserve(reserve);
    g.weight.reserve(reserve);
    g.dummy.reserve(reserve);
    g.next_out.reserve(reserve);
    g.head_out.assign(N, NONE32);
    g.out_deg.assign(N, 0);
    g.in_deg.assign(N, 0);
    Tail t(g, ctx->k);
    for (u64 e = 0; e < E; e++) t.add_edge(from[e], to[e], uw[e >> 1], false);
    // C. dummy edges in result order (greedytigs/mod.rs:678-689)
    const u32* tr = ctx->h_triples.data();
    for (u64 j = 0; j < n_trip; j++) t.add_dummy_pair(tr[3 * j], tr[3 * j + 1], tr[3 * j + 2]);
    eulerise(t);
    MTG_REQUIRE(is_eulerian(g), MTG_ERR_INTERNAL, "Failed to make the graph Eulerian.");
    ctx->walk_edges.clear();
    ctx->walk_limits.clear();
    ctx->walk_edges.reserve(g.from.size() / 2);
    WalkSink sink{g, ctx->k, ctx->walk_edges, ctx->walk_limits};
    u64 n_cycles = 0;
    euler_walks(g, sink, &n_cycles);
    for (size_t w = 0; w < ctx->walk_limits.size(); w++) {
        u64 b = w ? ctx->walk_limits[w - 1] : 0;
        MTG_REQUIRE(!g.dummy[ctx->walk_edges[b]], MTG_ERR_INTERNAL, "walk starts with a dummy edge");
    }
    // device copies for the output kernels
    ctx->d_walk_edges.upload(ctx->walk_edges.data(), ctx->walk_edges.size(), s);
    ctx->d_walk_limits.upload(ctx->walk_limits.data(), ctx->walk_limits.size(), s);
    std::vector<u32> dummy_w(g.weight.begin() + E, g.weight.end());
    ctx->d_dummy_w.upload(dummy_w.data(), dummy_w.size(), s);
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->have_walks = true;
}

}  // namespace mtg
