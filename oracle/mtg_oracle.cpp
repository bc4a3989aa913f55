// =====================================================================================
// mtg_oracle.cpp -- CPU restatement of the matchtigs 2.1.9 greedy-matchtig pipeline.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// path (matchtigs_b200/csrc) never links, loads or calls anything in oracle/.
//
// PARITY UNPINNED: the reference ships no golden vectors, no known-answer tests and
// no fixtures for this path (one #[test], no assertions: src/implementation/mod.rs:762-785),
// and the Rust toolchain plus the crates that hold Dijkstra / bigraph / the readers are
// absent from this environment (Cargo.lock: traitgraph-algo 8.1.2, bigraph 5.0.1,
// traitgraph 8.1.2, petgraph 0.7.1, genome-graph 11.0.0, compact-genome 12.0.1,
// disjoint-sets 0.4.2).  In-repo logic is restated from the cited file:line; the
// dependency semantics are restated from their published algorithms (SURVEY.md
// Appendix A) and every such assumption is named P1..P7 where it is used.  Each one is a switch
// (mto_set_option; same names as mtg_ctx_set_option of the product): a golden output of the real
// reference that contradicts an assumption is then a flag flip on both sides, not a rewrite.
//   p1_tie_desc, p1_exclusive_bound   Dijkstra settle order / bound        (traitgraph-algo)
//   p2_self_mirror_zero               imbalance of self-mirror nodes       (bigraph)
//   p3_oldest_first                   adjacency iteration order            (petgraph)
//   p4_*, p5_*                        Euler policy, FASTA numbering: registered, no alternative implemented (nonzero is refused)
//   p6_bcalm_kmer_numbering           bcalm2 reader numbers like the FASTA reader (genome-graph)
//   p7_first_root_wins                union-find tie rule                  (disjoint-sets)
//
// Everything is single-file C++17, no dependencies.  All arithmetic is integer.
// =====================================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;
constexpr u32 NONE = 0xFFFFFFFFu;

struct Options {
    int p1_tie_desc = 0, p1_exclusive_bound = 0, p2_self_mirror_zero = 0, p3_oldest_first = 0, p6_bcalm_kmer_numbering = 0,
        p7_first_root_wins = 0;
};

struct OracleError {
    std::string msg;
};
[[noreturn]] void fail(const std::string& m) { throw OracleError{m}; }

// -------------------------------------------------------------------------------------
// Graph container.  Restates petgraph 0.7.1 `Graph` as used through traitgraph's PetGraph
// (type alias src/bin.rs:349-355, src/clib.rs:37): nodes and edges get sequential indices,
// and every node keeps an intrusive singly linked list of outgoing / incoming edges with
// HEAD INSERTION, so neighbour iteration yields the most recently added edge first
// (assumption P3).  NodeBigraphWrapper adds the mirror-node table.
// -------------------------------------------------------------------------------------
struct Graph {
    std::vector<u32> mirror;              // NodeBigraphWrapper::mirror_node
    std::vector<u32> head_out, head_in;   // petgraph Node.next[0], next[1]
    std::vector<u32> from, to;            // per edge
    std::vector<u32> next_out, next_in;   // petgraph Edge.next[0], next[1]
    std::vector<u64> weight;              // CliEdgeData.weight (src/bin.rs:230)
    std::vector<u32> dummy_id;            // CliEdgeData.dummy_edge_id (src/bin.rs:232); 0 = original
    std::vector<u8> forward;              // CliEdgeData.forward (src/bin.rs:227)
    std::vector<u32> unitig;              // sequence handle (unitig index) for original edges
    std::vector<u32> out_deg, in_deg;
    std::vector<u32> tail_out, tail_in;   // last list element per node (only used with oldest_first)
    bool oldest_first = false;            // assumption P3 flipped: append at the tail instead of inserting at the head

    u32 node_count() const { return (u32)mirror.size(); }
    u32 edge_count() const { return (u32)from.size(); }
    u32 add_node() {
        mirror.push_back(NONE);
        head_out.push_back(NONE);
        head_in.push_back(NONE);
        tail_out.push_back(NONE);
        tail_in.push_back(NONE);
        out_deg.push_back(0);
        in_deg.push_back(0);
        return (u32)mirror.size() - 1;
    }
    void set_mirror_nodes(u32 a, u32 b) {
        mirror[a] = b;
        mirror[b] = a;
    }
    bool is_self_mirror(u32 v) const { return mirror[v] == v; }
    u32 add_edge(u32 a, u32 b, u32 handle, bool fwd, u64 w, u32 dummy) {
        u32 e = (u32)from.size();
        from.push_back(a);
        to.push_back(b);
        if (!oldest_first) {
            next_out.push_back(head_out[a]);
            head_out[a] = e;
            next_in.push_back(head_in[b]);
            head_in[b] = e;
        } else {
            next_out.push_back(NONE);
            next_in.push_back(NONE);
            if (tail_out[a] == NONE) head_out[a] = e;
            else next_out[tail_out[a]] = e;
            tail_out[a] = e;
            if (tail_in[b] == NONE) head_in[b] = e;
            else next_in[tail_in[b]] = e;
            tail_in[b] = e;
        }
        weight.push_back(w);
        dummy_id.push_back(dummy);
        forward.push_back(fwd ? 1 : 0);
        unitig.push_back(handle);
        out_deg[a]++;
        in_deg[b]++;
        return e;
    }
    // bigraph `mirror_edge_edge_centric(e: a->b)`: the out-edge of mirror(b) to mirror(a)
    // whose data equals data(e).mirror().  Every edge in this pipeline is added
    // together with its mirror as the pair (2j, 2j+1) -- originals src/clib.rs:239-248,
    // dummies greedytigs/mod.rs:683-688 and implementation/mod.rs:492-493,509-510,576-577 --
    // and (handle, dummy id) is unique per pair, so the equality search always resolves to e^1.
    u32 mirror_edge(u32 e) const { return e ^ 1u; }
    bool is_dummy(u32 e) const { return dummy_id[e] != 0; }
};

// bigraph::algo::eulerian::compute_eulerian_superfluous_out_biedges (assumption P2, SURVEY A.2;
// call site greedytigs/mod.rs:230).
inline i64 superfluous_out(const Graph& g, u32 v, bool self_mirror_zero = false) {
    if (g.is_self_mirror(v)) return self_mirror_zero ? 0 : (i64)(g.out_deg[v] % 2);
    return (i64)g.out_deg[v] - (i64)g.in_deg[v];
}

// -------------------------------------------------------------------------------------
// Sequences: the oracle keeps unitigs as ASCII (the reference keeps them 2-bit in
// compact-genome's DefaultSequenceStore<DnaAlphabet>; semantics are ACGT <-> TGCA).
// -------------------------------------------------------------------------------------
inline char comp(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
    }
    fail(std::string("non-ACGT character '") + c + "'");
}
std::string revcomp(const std::string& s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) r[s.size() - 1 - i] = comp(s[i]);
    return r;
}

struct Link {
    u32 a;
    u8 sa;
    u32 b;
    u8 sb;
};

struct ParsedFasta {
    std::vector<std::string> seqs;
    std::vector<Link> links;  // only filled in bcalm mode
};

// FASTA / bcalm2 record parsing (readers are external: genome-graph 11.0.0 io::{fasta,bcalm2},
// call sites src/bin.rs:896-899, 907-910; format SURVEY Appendix B).  Multi-line records are
// accepted; only upper-case ACGT is legal (DnaAlphabet).  In bcalm mode the record id must equal
// its position and `L:<s>:<j>:<t>` fields become links.
ParsedFasta parse_fasta(const char* text, size_t len, bool bcalm) {
    ParsedFasta out;
    size_t i = 0;
    while (i < len) {
        while (i < len && (text[i] == '\n' || text[i] == '\r')) i++;
        if (i >= len) break;
        if (text[i] != '>') fail("FASTA: expected '>'");
        size_t hs = i + 1;
        while (i < len && text[i] != '\n') i++;
        std::string header(text + hs, text + i);
        if (!header.empty() && header.back() == '\r') header.pop_back();
        u32 rec = (u32)out.seqs.size();
        if (bcalm) {
            size_t p = 0;
            while (p < header.size() && header[p] != ' ' && header[p] != '\t') p++;
            if ((u64)std::strtoull(header.substr(0, p).c_str(), nullptr, 10) != rec) fail("bcalm: record id != position");
            while (p < header.size()) {
                while (p < header.size() && (header[p] == ' ' || header[p] == '\t')) p++;
                size_t q = p;
                while (q < header.size() && header[q] != ' ' && header[q] != '\t') q++;
                if (q - p >= 7 && header[p] == 'L' && header[p + 1] == ':') {
                    // L:<+/->:<id>:<+/->
                    char s = header[p + 2];
                    size_t c2 = header.find(':', p + 4);
                    if (c2 == std::string::npos || c2 + 1 >= q) fail("bcalm: malformed L field");
                    u32 j = (u32)std::strtoull(header.substr(p + 4, c2 - (p + 4)).c_str(), nullptr, 10);
                    char t = header[c2 + 1];
                    if ((s != '+' && s != '-') || (t != '+' && t != '-')) fail("bcalm: malformed L sign");
                    out.links.push_back(Link{rec, (u8)(s == '+'), j, (u8)(t == '+')});
                }
                p = q;
            }
        }
        std::string seq;
        while (i < len && text[i] != '>') {
            char c = text[i++];
            if (c == '\n' || c == '\r') continue;
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') fail(std::string("FASTA: illegal character '") + c + "'");
            seq.push_back(c);
        }
        out.seqs.push_back(std::move(seq));
    }
    return out;
}

// disjoint-sets 0.4.2 UnionFind (assumption P7): union by rank; on equal rank the
// FIRST argument's root is attached below the SECOND argument's root.  Path compression does
// not change roots.
struct UnionFind {
    std::vector<u32> parent;
    std::vector<u8> rank;
    bool first_root_wins = false;  // assumption P7 flipped
    explicit UnionFind(size_t n) : parent(n), rank(n, 0) {
        for (size_t i = 0; i < n; i++) parent[i] = (u32)i;
    }
    u32 find(u32 x) {
        while (parent[x] != x) {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        return x;
    }
    void unite(u32 a, u32 b) {
        a = find(a);
        b = find(b);
        if (a == b) return;
        if (rank[a] > rank[b]) parent[b] = a;
        else if (rank[b] > rank[a]) parent[a] = b;
        else if (first_root_wins) {
            parent[b] = a;
            rank[a]++;
        } else {
            parent[a] = b;
            rank[b]++;
        }
    }
};

struct Walk {
    std::vector<u32> edges;
};

struct Stats {
    u64 dijkstra_calls = 0, settled = 0, relaxed = 0, heap_pops = 0, candidates = 0;
    // the reference's DijkstraPerformanceCounter (greedytigs/mod.rs:647-673): iterations = heap pops, unnecessary heap
    // elements = stale pops, heap / distance-array sizes as maxima per search (max and sum over the searches)
    u64 stale_pops = 0, max_max_heap = 0, sum_max_heap = 0, max_max_dist_array = 0, sum_max_dist_array = 0;
    u64 sources = 0, in_nodes = 0, self_mirror_unbalanced = 0;
    u64 breaking_edges = 0, cycles = 0;
    double t_parse = 0, t_build = 0, t_scan = 0, t_dijkstra = 0, t_insert = 0, t_eulerise = 0, t_euler = 0, t_break = 0, t_write = 0, t_bitvector = 0;
};

struct Oracle {
    u32 k = 0;
    bool have_seqs = false;
    std::vector<std::string> seqs;
    Graph g;
    u32 n_original_edges = 0;
    // phase outputs
    std::vector<u32> out_nodes;        // sources, ascending (greedytigs/mod.rs:223,235,242)
    std::vector<u8> in_node_map0;      // initial target map (greedytigs/mod.rs:225,233,239)
    std::vector<i64> mult0;            // initial node_multiplicities (greedytigs/mod.rs:226)
    std::vector<u32> triples;          // (out,in,dist)* in `results` order (greedytigs/mod.rs:461,646)
    std::vector<std::vector<u32>> cycles;
    std::vector<Walk> walks;
    std::string gfa, fasta, bitvec;
    std::vector<i64> c_edge_out;
    std::vector<u64> c_insert_out, c_limits;
    Stats st;
    Options opt;
    std::string err;
    int euler_fast = 0;
    // which outputs run_greedy produces (bench arms produce exactly what the GPU arm produces): 1 GFA, 2 FASTA, 4 bitvector, 8 C API
    int outputs = 15;
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// -------------------------------------------------------------------------------------
// Graph construction, FASTA flavour (assumption P5; call site src/bin.rs:896-899).
// Records in file order; record i => edges 2i (forward: prefix node -> suffix node) and 2i+1
// (rc(suffix) node -> rc(prefix) node).  Nodes are created on first sight through a
// (k-1)-mer -> node map: a missing (k-1)-mer creates its node and, unless palindromic, directly
// afterwards the node of its reverse complement.
// -------------------------------------------------------------------------------------
void build_from_kmers(Oracle& o) {
    const u32 k = o.k;
    Graph& g = o.g;
    std::unordered_map<std::string, u32> id_map;
    auto get_or_create = [&](const std::string& kmer) -> u32 {
        auto it = id_map.find(kmer);
        if (it != id_map.end()) return it->second;
        u32 n = g.add_node();
        std::string rc = revcomp(kmer);
        if (rc == kmer) {
            g.set_mirror_nodes(n, n);
            id_map.emplace(kmer, n);
        } else {
            u32 m = g.add_node();
            g.set_mirror_nodes(n, m);
            id_map.emplace(kmer, n);
            id_map.emplace(rc, m);
        }
        return n;
    };
    for (u32 i = 0; i < o.seqs.size(); i++) {
        const std::string& s = o.seqs[i];
        if (s.size() < k) fail("sequence shorter than k");
        std::string pre = s.substr(0, k - 1), suf = s.substr(s.size() - (k - 1));
        u32 pre_plus = get_or_create(pre);
        u32 pre_minus = g.mirror[pre_plus];
        u32 suf_plus = get_or_create(suf);
        u32 suf_minus = g.mirror[suf_plus];
        // compute_edge_weights: weight = len + 1 - k (src/bin.rs:359-379)
        u64 w = (u64)s.size() + 1 - k;
        g.add_edge(pre_plus, suf_plus, i, true, w, 0);
        g.add_edge(suf_minus, pre_minus, i, false, w, 0);
    }
    o.n_original_edges = g.edge_count();
}

// Endpoint slots of the C API (src/clib.rs:104-122).
inline u32 fwd_in(u32 u) { return u * 4; }
inline u32 fwd_out(u32 u) { return u * 4 + 2; }
inline u32 bwd_in(u32 u) { return u * 4 + 3; }
inline u32 bwd_out(u32 u) { return u * 4 + 1; }

// Graph construction from links: matchtigs_merge_nodes (src/clib.rs:135-170) for every link in
// order, then matchtigs_build_graph (src/clib.rs:180-259).  Used for the C API and, under
// assumption P6, for --bcalm-in.
void build_from_links(Oracle& o, u32 U, const std::vector<Link>& links, const std::vector<u64>& weights) {
    Graph& g = o.g;
    UnionFind uf((size_t)U * 4);
    uf.first_root_wins = o.opt.p7_first_root_wins != 0;
    for (const Link& l : links) {
        if (l.a >= U || l.b >= U) fail("link references unknown unitig");
        u32 out_a = l.sa ? fwd_out(l.a) : bwd_out(l.a);
        u32 in_b = l.sb ? fwd_in(l.b) : bwd_in(l.b);
        u32 mirror_in_a = l.sa ? bwd_in(l.a) : fwd_in(l.a);
        u32 mirror_out_b = l.sb ? bwd_out(l.b) : fwd_out(l.b);
        uf.unite(out_a, in_b);
        uf.unite(mirror_in_a, mirror_out_b);
    }
    std::vector<u32> reps((size_t)U * 4);
    for (u32 i = 0; i < U * 4; i++) reps[i] = uf.find(i);
    std::sort(reps.begin(), reps.end());
    reps.erase(std::unique(reps.begin(), reps.end()), reps.end());
    for (size_t i = 0; i < reps.size(); i++) g.add_node();
    auto node_of = [&](u32 slot) -> u32 {
        u32 r = uf.find(slot);
        return (u32)(std::lower_bound(reps.begin(), reps.end(), r) - reps.begin());
    };
    for (u32 u = 0; u < U; u++) {
        u32 n1 = node_of(fwd_in(u)), n2 = node_of(fwd_out(u));
        u32 mirror_n2 = node_of(bwd_in(u)), mirror_n1 = node_of(bwd_out(u));
        g.set_mirror_nodes(n1, mirror_n1);
        g.set_mirror_nodes(n2, mirror_n2);
        g.add_edge(n1, n2, u, true, weights[u], 0);
        g.add_edge(mirror_n2, mirror_n1, u, false, weights[u], 0);
    }
    // verify_node_pairing (src/clib.rs:251): mirror is an involution on all nodes.
    for (u32 v = 0; v < g.node_count(); v++)
        if (g.mirror[v] == NONE || g.mirror[g.mirror[v]] != v) fail("node pairing violated (inconsistent links)");
    o.n_original_edges = g.edge_count();
}

// -------------------------------------------------------------------------------------
// Dijkstra::shortest_path_lens (traitgraph-algo 8.1.2; assumption P1; call site
// greedytigs/mod.rs:324-335).  Lazy-deletion min-heap on (weight, node index); stops when the
// popped weight exceeds max_weight (inclusive bound) or target_amount targets were collected.
// The epoch array mirrors EpochNodeWeightArray (greedytigs/mod.rs:134-155): results do not
// depend on the weight store.
// -------------------------------------------------------------------------------------
struct Dijkstra {
    std::vector<u32> epoch_of;
    std::vector<u32> dist;
    u32 epoch = 0;
    u32 tie_flip = 0;  // assumption P1 flipped (p1_tie_desc): ties inside a distance go to the larger node id
    std::priority_queue<std::pair<u64, u32>, std::vector<std::pair<u64, u32>>, std::greater<>> heap;  // (weight, node ^ tie_flip)
    explicit Dijkstra(u32 n) : epoch_of(n, 0), dist(n, 0) {}

    template <class IsTarget>
    void shortest_path_lens(const Graph& g, u32 source, IsTarget&& is_target, size_t target_amount, u64 max_weight,
                            bool forbid_source_target, std::vector<std::pair<u32, u64>>& distances, Stats* st) {
        epoch++;
        while (!heap.empty()) heap.pop();
        distances.clear();
        auto get = [&](u32 v) -> u64 { return epoch_of[v] == epoch ? dist[v] : ~0ull; };
        auto set = [&](u32 v, u64 w) {
            epoch_of[v] = epoch;
            dist[v] = (u32)w;
        };
        set(source, 0);
        heap.push({0, source ^ tie_flip});
        u64 settled = 0, relaxed = 0, pops = 0, stale = 0, labels = 1, max_heap = 1;
        while (!heap.empty()) {
            auto [w, vk] = heap.top();
            const u32 v = vk ^ tie_flip;
            heap.pop();
            pops++;
            if (get(v) < w) {  // stale
                stale++;
                continue;
            }
            if (w > max_weight) break;
            settled++;
            if (is_target(v) && !(forbid_source_target && v == source)) {
                distances.push_back({v, w});
                if (distances.size() == target_amount) break;
            }
            for (u32 e = g.head_out[v]; e != NONE; e = g.next_out[e]) {
                relaxed++;
                u64 nw = w + g.weight[e];
                u32 u = g.to[e];
                if (nw < get(u)) {
                    labels += epoch_of[u] != epoch;
                    set(u, nw);
                    heap.push({nw, u ^ tie_flip});
                    max_heap = std::max<u64>(max_heap, heap.size());
                }
            }
        }
        if (st) {
            st->dijkstra_calls++;
            st->settled += settled;
            st->relaxed += relaxed;
            st->heap_pops += pops;
            st->stale_pops += stale;
            st->max_max_heap = std::max(st->max_max_heap, max_heap);
            st->sum_max_heap += max_heap;
            st->max_max_dist_array = std::max(st->max_max_dist_array, labels);
            st->sum_max_dist_array += labels;
        }
    }
};

// Byte spin locks standing in for Vec<Mutex<isize>> (greedytigs/mod.rs:268-272).
struct SpinLocks {
    std::unique_ptr<std::atomic<u8>[]> l;
    explicit SpinLocks(size_t n) : l(new std::atomic<u8>[n]) {
        for (size_t i = 0; i < n; i++) l[i].store(0, std::memory_order_relaxed);
    }
    void lock(size_t i) {
        u8 e = 0;
        while (!l[i].compare_exchange_weak(e, 1, std::memory_order_acquire)) e = 0;
    }
    void unlock(size_t i) { l[i].store(0, std::memory_order_release); }
};

// -------------------------------------------------------------------------------------
// compute_greedytigs phases A+B (greedytigs/mod.rs:222-646).  With threads == 1 this is the
// normative, deterministic semantics (assumption P8).  threads > 1 restates the reference's
// worker scheme (shared offset, adaptive chunks :557-627, sorted multi-lock :366-397) and is
// used for CPU-baseline timing only: like the reference it is not deterministic.
// -------------------------------------------------------------------------------------
void greedy_paths(Oracle& o, u32 threads) {
    Graph& g = o.g;
    const u32 n = g.node_count();
    const u64 k = o.k;
    double t0 = now_s();
    // A. imbalance scan (greedytigs/mod.rs:222-245)
    o.out_nodes.clear();
    o.in_node_map0.assign(n, 0);
    o.mult0.assign(n, 0);
    for (u32 v = 0; v < n; v++) {
        i64 diff = superfluous_out(g, v, o.opt.p2_self_mirror_zero != 0);
        if (g.is_self_mirror(v) && diff != 0) {
            o.st.in_nodes++;
            o.in_node_map0[v] = 1;
            o.mult0[v] = diff;
            o.out_nodes.push_back(v);
            o.st.self_mirror_unbalanced++;
        } else if (diff > 0) {
            o.st.in_nodes++;
            o.in_node_map0[v] = 1;
            o.mult0[v] = diff;
        } else if (diff < 0) {
            o.out_nodes.push_back(v);
            o.mult0[v] = diff;
        }
    }
    o.st.sources = o.out_nodes.size();
    double t1 = now_s();
    o.st.t_scan += t1 - t0;

    // B. Dijkstras + on-the-fly matching (greedytigs/mod.rs:276-646)
    std::unique_ptr<std::atomic<u8>[]> in_node_map(new std::atomic<u8>[n ? n : 1]);
    for (u32 v = 0; v < n; v++) in_node_map[v].store(o.in_node_map0[v], std::memory_order_relaxed);
    std::vector<i64> mult = o.mult0;
    SpinLocks locks(n ? n : 1);
    std::mutex results_mutex, offset_mutex;
    std::vector<u32>& results = o.triples;
    results.clear();
    size_t offset = 0;
    const std::vector<u32>& out_nodes = o.out_nodes;
    std::vector<Stats> tstats(threads);

    const u64 max_weight = k - 1 - (o.opt.p1_exclusive_bound ? 1 : 0);
    auto compute_dijkstras = [&](Dijkstra& dj, std::vector<std::pair<u32, u64>>& distances, std::vector<u32>& shortest_paths,
                                 size_t lo, size_t hi, Stats& st) {
        auto is_target = [&](u32 v) { return in_node_map[v].load(std::memory_order_relaxed) != 0; };
        for (size_t i = lo; i < hi; i++) {
            u32 out_node = out_nodes[i];
            bool out_self = g.is_self_mirror(out_node);
            u32 out_mirror = g.mirror[out_node];
            locks.lock(out_mirror);
            i64 m = mult[out_mirror];  // :306-311
            locks.unlock(out_mirror);
            if (m == 0) continue;  // :318-320
            while (m > 0) {        // :322
                size_t target_amount = (size_t)(m + 1);
                dj.shortest_path_lens(g, out_node, is_target, target_amount, max_weight, true, distances, &st);
                if (distances.empty()) break;  // :338-346
                bool abort_after_this = distances.size() < target_amount;  // :348
                for (auto& [in_node, distance] : distances) {              // :350
                    bool is_self_mirror_edge = false;
                    if (in_node == out_mirror) {  // :352-358
                        if (m < 2) continue;
                        is_self_mirror_edge = true;
                    }
                    u32 in_mirror = g.mirror[in_node];
                    bool in_self = g.is_self_mirror(in_node);
                    // lock set, acquired in ascending index order (:366-391)
                    u32 idx[4];
                    int ni = 0;
                    idx[ni++] = out_node;
                    if (!out_self) idx[ni++] = out_mirror;
                    if (!is_self_mirror_edge) {
                        idx[ni++] = in_node;
                        if (!in_self) idx[ni++] = in_mirror;
                    }
                    u32 sorted[4];
                    std::copy(idx, idx + ni, sorted);
                    std::sort(sorted, sorted + ni);
                    int ns = (int)(std::unique(sorted, sorted + ni) - sorted);
                    for (int q = 0; q < ns; q++) locks.lock(sorted[q]);
                    auto unlock_all = [&] {
                        for (int q = ns - 1; q >= 0; q--) locks.unlock(sorted[q]);
                    };
                    const int in_off = out_self ? 1 : 2;                 // :398
                    const i64 reduction = is_self_mirror_edge ? 2 : 1;  // :399
                    m = out_self ? mult[idx[0]] : -mult[idx[0]];        // :401-410
                    if (m == 0) {                                        // :412-414
                        unlock_all();
                        break;
                    }
                    if (!is_self_mirror_edge) {  // :416-459
                        if (mult[idx[in_off]] == 0) {
                            in_node_map[in_node].store(0, std::memory_order_relaxed);
                            unlock_all();
                            continue;
                        }
                    }
                    shortest_paths.push_back(out_node);  // :461
                    shortest_paths.push_back(in_node);
                    shortest_paths.push_back((u32)distance);
                    if (out_self) {  // :463-473
                        mult[idx[0]] -= 1;
                    } else {
                        mult[idx[0]] += reduction;
                        mult[idx[1]] -= reduction;
                    }
                    m = -mult[idx[0]];           // :474
                    if (!is_self_mirror_edge) {  // :476-491
                        mult[idx[in_off]] -= 1;
                        if (!in_self) mult[idx[in_off + 1]] += 1;
                    }
                    if (m == 0) in_node_map[out_mirror].store(0, std::memory_order_relaxed);  // :493-495
                    if (!is_self_mirror_edge && mult[idx[in_off]] == 0)                         // :497-501
                        in_node_map[in_node].store(0, std::memory_order_relaxed);
                    unlock_all();
                }
                if (abort_after_this) break;  // :504-511
            }
        }
    };

    auto worker = [&](u32 tid) {
        Dijkstra dj(n);
        dj.tie_flip = o.opt.p1_tie_desc ? 0xFFFFFFFFu : 0u;
        std::vector<std::pair<u32, u64>> distances;
        std::vector<u32> shortest_paths;
        size_t chunk_size = 1024;  // :570
        for (;;) {
            size_t cur, lim;
            {
                std::lock_guard<std::mutex> lk(offset_mutex);
                cur = offset;
                size_t remaining = out_nodes.size() - offset;
                if (remaining == 0) break;
                // :583-587
                chunk_size = std::min(std::max(std::max(std::min(chunk_size, remaining / threads), (size_t)10), chunk_size / 10), remaining);
                lim = cur + chunk_size;
                offset = lim;
            }
            double s = now_s();
            compute_dijkstras(dj, distances, shortest_paths, cur, lim, tstats[tid]);
            double d = now_s() - s;
            if (d <= 0) d = 1e-9;
            double cs = (double)chunk_size * (5.0 / d);  // TARGET_DIJKSTRA_BLOCK_TIME = 5 s (implementation/mod.rs:35)
            chunk_size = cs > 1e15 ? (size_t)1e15 : (size_t)cs;
            chunk_size = std::max(chunk_size, (size_t)10);
        }
        std::lock_guard<std::mutex> lk(results_mutex);  // :618
        results.insert(results.end(), shortest_paths.begin(), shortest_paths.end());
    };
    if (threads <= 1) {
        worker(0);
    } else {
        std::vector<std::thread> ts;
        for (u32 t = 0; t < threads; t++) ts.emplace_back(worker, t);
        for (auto& t : ts) t.join();
    }
    for (auto& s : tstats) {
        o.st.dijkstra_calls += s.dijkstra_calls;
        o.st.settled += s.settled;
        o.st.relaxed += s.relaxed;
        o.st.heap_pops += s.heap_pops;
        o.st.stale_pops += s.stale_pops;
        o.st.max_max_heap = std::max(o.st.max_max_heap, s.max_max_heap);
        o.st.sum_max_heap += s.sum_max_heap;
        o.st.max_max_dist_array = std::max(o.st.max_max_dist_array, s.max_max_dist_array);
        o.st.sum_max_dist_array += s.sum_max_dist_array;
    }
    o.st.t_dijkstra += now_s() - t1;
}

// C. dummy-edge insertion in `results` order (greedytigs/mod.rs:678-689).
u32 insert_dummies(Oracle& o) {
    double t0 = now_s();
    Graph& g = o.g;
    u32 dummy_edge_id = 0;
    for (size_t j = 0; j + 3 <= o.triples.size(); j += 3) {
        u32 out_node = o.triples[j], in_node = o.triples[j + 1];
        u64 dist = o.triples[j + 2];
        dummy_edge_id++;
        g.add_edge(out_node, in_node, NONE, true, dist, dummy_edge_id);
        g.add_edge(g.mirror[in_node], g.mirror[out_node], NONE, false, dist, dummy_edge_id);
    }
    o.st.t_insert += now_s() - t0;
    return dummy_edge_id;
}

// bigraph find_non_eulerian_binodes_with_differences (use implementation/mod.rs:408-427).
std::vector<std::pair<u32, i64>> non_eulerian(const Graph& g) {
    std::vector<std::pair<u32, i64>> r;
    for (u32 v = 0; v < g.node_count(); v++) {
        if (g.is_self_mirror(v)) {
            if (g.out_deg[v] % 2 != 0) r.push_back({v, 0});
        } else {
            i64 d = (i64)g.out_deg[v] - (i64)g.in_deg[v];
            if (d != 0) r.push_back({v, d});
        }
    }
    return r;
}

// D. make_graph_eulerian_with_breaking_edges (implementation/mod.rs:392-649) with
// choose_in_node_from_iterator (:252-285).
void eulerise(Oracle& o, u32& dummy_edge_id) {
    double t0 = now_s();
    Graph& g = o.g;
    const u64 k = o.k;
    auto nd = non_eulerian(g);
    std::map<u32, i64, std::greater<u32>> out_diff;  // BTreeMap<Reverse<node>, diff> :409,416
    std::map<u32, i64> in_diff;                      // BTreeMap<node, diff> :410
    std::vector<u32> self_mirrors;
    for (auto& [v, d] : nd) {
        if (d < 0) out_diff[v] = d;
        else if (d > 0) in_diff[v] = d;
        else self_mirrors.push_back(v);
    }
    auto add_pair = [&](u32 out_node, u32 in_node) {
        u32 mirror_out_node = g.mirror[in_node], mirror_in_node = g.mirror[out_node];
        dummy_edge_id++;
        g.add_edge(out_node, in_node, NONE, true, k, dummy_edge_id);
        g.add_edge(mirror_out_node, mirror_in_node, NONE, false, k, dummy_edge_id);
    };
    // self-mirrors in chunks of two (:481-524)
    for (size_t i = 0; i < self_mirrors.size(); i += 2) {
        if (i + 1 < self_mirrors.size()) {
            add_pair(self_mirrors[i], self_mirrors[i + 1]);
        } else {
            if (in_diff.empty()) fail("Have an uneven number of self-mirrors, but no other nodes with missing in edges.");
            auto it = in_diff.begin();
            u32 in_node = it->first;
            add_pair(self_mirrors[i], in_node);
            it->second -= 1;
            if (it->second == 0) {
                in_diff.erase(it);
                if (!out_diff.erase(g.mirror[in_node])) fail("Mirror of in_node not found");
            } else {
                auto mo = out_diff.find(g.mirror[in_node]);
                if (mo == out_diff.end()) fail("Mirror of in_node not found");
                mo->second += 1;
            }
        }
    }
    // main loop (:526-645)
    while (!out_diff.empty()) {
        auto oit = out_diff.begin();  // largest node id
        u32 out_node = oit->first;
        i64 od = oit->second;
        // choose_in_node_from_iterator (:252-285)
        auto iit = in_diff.begin();
        if (iit == in_diff.end()) fail("No further in_nodes left");
        if ((iit->first == g.mirror[out_node] && od > -2) || iit->first == out_node) {
            ++iit;
            if (iit == in_diff.end()) fail("No further in_nodes left");
        }
        u32 in_node = iit->first;
        u32 mirror_out_node = g.mirror[in_node], mirror_in_node = g.mirror[out_node];
        add_pair(out_node, in_node);
        oit->second += 1;
        iit->second -= 1;
        bool remove_out = oit->second == 0, remove_in = iit->second == 0;
        if (remove_out) out_diff.erase(oit);
        if (remove_in) in_diff.erase(iit);
        auto mo = out_diff.find(mirror_out_node);  // :609-627
        if (mo != out_diff.end()) {
            mo->second += 1;
            if (mo->second == 0) out_diff.erase(mo);
        }
        auto mi = in_diff.find(mirror_in_node);  // :628-644
        if (mi != in_diff.end()) {
            mi->second -= 1;
            if (mi->second == 0) in_diff.erase(mi);
        }
    }
    if (!in_diff.empty()) fail("eulerise: in-nodes left over");
    o.st.t_eulerise += now_s() - t0;
}

// F. bigraph compute_minimum_bidirected_eulerian_cycle_decomposition (assumption P4, SURVEY A.4;
// call site greedytigs/mod.rs:722).  `faithful`: std::vector + rotate exactly as recalled.
void euler_decomposition_faithful(Oracle& o) {
    const Graph& g = o.g;
    std::vector<u8> used(g.edge_count(), 0);
    o.cycles.clear();
    for (u32 e0 = 0; e0 < g.edge_count(); e0++) {
        if (used[e0]) continue;
        std::vector<u32> cycle;
        u32 start_edge = e0;
        while (start_edge != NONE) {
            used[start_edge] = 1;
            used[g.mirror_edge(start_edge)] = 1;
            cycle.push_back(start_edge);
            u32 cur = g.to[start_edge];
            bool has_neighbor = true;
            while (has_neighbor) {
                has_neighbor = false;
                for (u32 e = g.head_out[cur]; e != NONE; e = g.next_out[e]) {
                    if (!used[e]) {
                        cycle.push_back(e);
                        used[e] = 1;
                        used[g.mirror_edge(e)] = 1;
                        has_neighbor = true;
                        cur = g.to[e];
                        break;
                    }
                }
            }
            start_edge = NONE;
            for (size_t ci = 0; ci < cycle.size(); ci++) {
                u32 fn = g.from[cycle[ci]];
                u32 found = NONE;
                for (u32 e = g.head_out[fn]; e != NONE; e = g.next_out[e])
                    if (!used[e]) {
                        found = e;
                        break;
                    }
                if (found != NONE) {
                    start_edge = found;
                    std::rotate(cycle.begin(), cycle.begin() + ci, cycle.end());
                    break;
                }
            }
        }
        o.cycles.push_back(std::move(cycle));
    }
}

// Same function, linear time: the cycle is a circular linked list (rotate == move the head),
// per-node cursors skip used out-edges, and a second ring of not-yet-exhausted positions makes
// each re-root scan amortised O(1).  Must produce identical cycles (tests check it); used so
// that the CPU baseline is not handicapped by the quadratic rotate.
void euler_decomposition_fast(Oracle& o) {
    const Graph& g = o.g;
    const u32 E = g.edge_count();
    std::vector<u8> used(E, 0);
    std::vector<u32> cursor(g.head_out);  // first possibly-unused out-edge per node
    std::vector<u32> nxt(E, NONE), prv(E, NONE);    // ring over edges of the current cycle (index = edge id)
    std::vector<u32> cnxt(E, NONE), cprv(E, NONE);  // candidate ring (subset, same circular order)
    auto first_unused = [&](u32 v) -> u32 {
        u32 e = cursor[v];
        while (e != NONE && used[e]) e = g.next_out[e];
        cursor[v] = e;
        return e;
    };
    o.cycles.clear();
    for (u32 e0 = 0; e0 < E; e0++) {
        if (used[e0]) continue;
        u32 head = NONE;   // first element of the cycle vector
        u32 chead = NONE;  // candidate-ring element that is the first candidate at/after head
        size_t len = 0;
        auto push_back = [&](u32 e) {  // append at the end of the vector == insert before head in both rings
            if (head == NONE) {
                head = e;
                nxt[e] = prv[e] = e;
                chead = e;
                cnxt[e] = cprv[e] = e;
            } else {
                u32 tail = prv[head];
                nxt[tail] = e;
                prv[e] = tail;
                nxt[e] = head;
                prv[head] = e;
                if (chead == NONE) {
                    chead = e;
                    cnxt[e] = cprv[e] = e;
                } else {
                    // candidates are kept in cycle order starting from chead (== first candidate at or after head),
                    // so the new last element goes right before chead.
                    u32 ct = cprv[chead];
                    cnxt[ct] = e;
                    cprv[e] = ct;
                    cnxt[e] = chead;
                    cprv[chead] = e;
                }
            }
            len++;
        };
        u32 start_edge = e0;
        while (start_edge != NONE) {
            used[start_edge] = 1;
            used[g.mirror_edge(start_edge)] = 1;
            push_back(start_edge);
            u32 cur = g.to[start_edge];
            for (;;) {
                u32 e = first_unused(cur);
                if (e == NONE) break;
                used[e] = 1;
                used[g.mirror_edge(e)] = 1;
                push_back(e);
                cur = g.to[e];
            }
            // re-root: first position (from head) whose from-node still has an unused out-edge
            start_edge = NONE;
            while (chead != NONE) {
                u32 found = first_unused(g.from[chead]);
                if (found != NONE) {
                    start_edge = found;
                    head = chead;  // rotate_left(ci)
                    break;
                }
                // exhausted for good: drop from the candidate ring
                if (cnxt[chead] == chead) {
                    chead = NONE;
                } else {
                    u32 a = cprv[chead], b = cnxt[chead];
                    cnxt[a] = b;
                    cprv[b] = a;
                    chead = b;
                }
            }
        }
        std::vector<u32> cycle;
        cycle.reserve(len);
        u32 e = head;
        for (size_t i = 0; i < len; i++) {
            cycle.push_back(e);
            e = nxt[e];
        }
        o.cycles.push_back(std::move(cycle));
    }
}

// G. rotate each cycle to its heaviest dummy and break it (greedytigs/mod.rs:726-789).
void break_cycles(Oracle& o) {
    double t0 = now_s();
    const Graph& g = o.g;
    const u64 k = o.k;
    o.walks.clear();
    for (auto& cyc : o.cycles) {
        u64 longest_w = 0;
        size_t longest_i = 0;
        for (size_t i = 0; i < cyc.size(); i++) {
            u32 e = cyc[i];
            if (g.is_dummy(e) && g.weight[e] > longest_w) {  // strict > (:741)
                longest_w = g.weight[e];
                longest_i = i;
            }
        }
        if (longest_w > 0) std::rotate(cyc.begin(), cyc.begin() + longest_i, cyc.end());
        size_t offset = 0;
        for (size_t i = 0; i < cyc.size(); i++) {
            u32 e = cyc[i];
            if ((g.weight[e] >= k && g.is_dummy(e)) || (g.is_dummy(e) && i == 0)) {  // :767-769
                if (offset < i) o.walks.push_back(Walk{std::vector<u32>(cyc.begin() + offset, cyc.begin() + i)});
                offset = i + 1;
                o.st.breaking_edges++;
            }
        }
        if (offset < cyc.size()) {  // :779-788
            if (!g.is_dummy(cyc.back())) o.walks.push_back(Walk{std::vector<u32>(cyc.begin() + offset, cyc.end())});
            else if (offset < cyc.size() - 1) o.walks.push_back(Walk{std::vector<u32>(cyc.begin() + offset, cyc.end() - 1)});
        }
    }
    o.st.cycles = o.cycles.size();
    o.st.t_break += now_s() - t0;
}

// write_walks_gfa (src/bin.rs:667-818) / write_walks_fasta (:466-606): identical sequence
// assembly, different framing.
void assemble(const Oracle& o, std::string& out, bool gfa) {
    const Graph& g = o.g;
    const size_t k = o.k;
    if (gfa) out += "H\tKL:Z:" + std::to_string(k) + "\n";  // :688-693
    for (size_t i = 0; i < o.walks.size(); i++) {
        const auto& w = o.walks[i].edges;
        if (gfa) out += "S\t" + std::to_string(i + 1) + "\t";  // :704
        else out += ">" + std::to_string(i + 1) + "\n";        // :492
        u32 first = w[0];
        if (g.is_dummy(first)) fail("walk starts with a dummy edge");
        const std::string& fs = o.seqs[g.unitig[first]];
        out += g.forward[first] ? fs : revcomp(fs);  // :709-713
        u32 prev = first;
        for (size_t j = 1; j < w.size(); j++) {
            u32 cur = w[j];
            if (g.is_dummy(cur)) {  // :731-743
                prev = cur;
                continue;
            }
            size_t off = !g.is_dummy(prev) ? k - 1 : k - 1 - (size_t)g.weight[prev];  // :745-749
            const std::string& s = o.seqs[g.unitig[cur]];
            if (g.forward[cur]) {
                out.append(s, off, std::string::npos);  // :751-778
            } else {
                out += revcomp(s.substr(0, s.size() - off));  // :780-808
            }
            prev = cur;
        }
        out += "\n";
    }
}

// write_duplication_bitvector (src/implementation/mod.rs:671-702).
void bitvector(const Oracle& o, std::string& out) {
    const Graph& g = o.g;
    for (auto& wk : o.walks) {
        if (wk.edges.empty()) fail("Found empty walk when writing duplication bitvector");
        for (u32 e : wk.edges) out.append((size_t)g.weight[e], g.is_dummy(e) ? '0' : '1');
        out += "\n";
    }
}

// C-API output encoding (src/clib.rs:393-407).
void encode_capi(Oracle& o) {
    const Graph& g = o.g;
    o.c_edge_out.clear();
    o.c_insert_out.clear();
    o.c_limits.clear();
    u64 limit = 0;
    for (auto& wk : o.walks) {
        for (u32 e : wk.edges) {
            // dummy edges carry the default sequence handle, i.e. unitig id 0 (src/clib.rs:63-69 with handle default)
            i64 uid = g.is_dummy(e) ? 0 : (i64)g.unitig[e];
            o.c_edge_out.push_back(uid * (g.forward[e] ? 1 : -1));
            o.c_insert_out.push_back(g.is_dummy(e) ? g.weight[e] : 0);
        }
        limit += wk.edges.size();
        o.c_limits.push_back(limit);
    }
}

void run_greedy(Oracle& o, u32 threads) {
    greedy_paths(o, threads);
    u32 dummy_edge_id = insert_dummies(o);
    eulerise(o, dummy_edge_id);
    if (!non_eulerian(o.g).empty()) fail("Failed to make the graph Eulerian.");  // greedytigs/mod.rs:708-715
    double t0 = now_s();
    if (o.euler_fast) euler_decomposition_fast(o);
    else euler_decomposition_faithful(o);
    o.st.t_euler += now_s() - t0;
    break_cycles(o);
    double t1 = now_s();
    o.gfa.clear();
    o.fasta.clear();
    o.bitvec.clear();
    if (o.have_seqs && (o.outputs & 1)) assemble(o, o.gfa, true);
    if (o.have_seqs && (o.outputs & 2)) assemble(o, o.fasta, false);
    double t2 = now_s();
    if (o.outputs & 4) bitvector(o, o.bitvec);
    double t3 = now_s();
    if (o.outputs & 8) encode_capi(o);
    o.st.t_write += (t2 - t1) + (now_s() - t3);
    o.st.t_bitvector += t3 - t2;
}

// L(src): every initially-open in-node within distance k-1 of `src`, in settle order
// (dist, node id) -- SURVEY 3.3.  Used by the tests to check the CUDA candidate lists entry by
// entry; it is the same Dijkstra with target_amount = infinity against the initial target map.
void candidates_of(Oracle& o, Dijkstra& dj, u32 src, std::vector<std::pair<u32, u64>>& out) {
    auto is_target = [&](u32 v) { return o.in_node_map0[v] != 0; };
    dj.tie_flip = o.opt.p1_tie_desc ? 0xFFFFFFFFu : 0u;
    dj.shortest_path_lens(o.g, src, is_target, (size_t)-1, (u64)o.k - 1 - (o.opt.p1_exclusive_bound ? 1 : 0), true, out, nullptr);
}

}  // namespace

// =====================================================================================
// C ABI for the test harness (ctypes).  Return 0 on success, -1 on error (mto_error()).
// =====================================================================================
extern "C" {

void* mto_create() { return new Oracle(); }
void mto_destroy(void* h) { delete (Oracle*)h; }
const char* mto_error(void* h) { return ((Oracle*)h)->err.c_str(); }

#define MTO_TRY(...)                     \
    Oracle& o = *(Oracle*)h;             \
    try {                                \
        __VA_ARGS__;                        \
        return 0;                        \
    } catch (const OracleError& e) {     \
        o.err = e.msg;                   \
        return -1;                       \
    } catch (const std::exception& e) {  \
        o.err = e.what();                \
        return -1;                       \
    }

static void reset_keep_options(Oracle& o) {
    const int ef = o.euler_fast, outs = o.outputs;
    const Options opt = o.opt;
    o = Oracle();
    o.euler_fast = ef;
    o.outputs = outs;
    o.opt = opt;
    o.g.oldest_first = opt.p3_oldest_first != 0;
}

// mode 0: --fa-in semantics (k-mer hashing); mode 1: --bcalm-in semantics (links, P6).
int mto_load_text(void* h, const char* text, size_t len, int k, int mode) {
    MTO_TRY({
        double t0 = now_s();
        reset_keep_options(o);
        o.k = (u32)k;
        if (k < 2) fail("k must be >= 2");
        ParsedFasta pf = parse_fasta(text, len, mode == 1);
        o.st.t_parse = now_s() - t0;
        t0 = now_s();
        o.seqs = std::move(pf.seqs);
        o.have_seqs = true;
        if (mode == 0 || o.opt.p6_bcalm_kmer_numbering) {  // P6 flipped: the bcalm2 reader numbers like the FASTA reader
            build_from_kmers(o);
        } else {
            std::vector<u64> w(o.seqs.size());
            for (size_t i = 0; i < o.seqs.size(); i++) {
                if (o.seqs[i].size() < (size_t)k) fail("sequence shorter than k");
                w[i] = o.seqs[i].size() + 1 - k;
            }
            build_from_links(o, (u32)o.seqs.size(), pf.links, w);
        }
        o.st.t_build = now_s() - t0;
    })
}

// C-API flavour: links as 4 parallel arrays, weights per unitig (src/clib.rs:135-259).
int mto_load_links(void* h, size_t U, const size_t* weights, size_t n_links, const size_t* a, const uint8_t* sa,
                   const size_t* b, const uint8_t* sb, int k) {
    MTO_TRY({
        double t0 = now_s();
        reset_keep_options(o);
        o.k = (u32)k;
        std::vector<Link> links(n_links);
        for (size_t i = 0; i < n_links; i++) links[i] = Link{(u32)a[i], sa[i], (u32)b[i], sb[i]};
        std::vector<u64> w(weights, weights + U);
        build_from_links(o, (u32)U, links, w);
        o.st.t_build = now_s() - t0;
    })
}

int mto_set_option(void* h, const char* name, int value) {
    Oracle& o = *(Oracle*)h;
    if (!std::strcmp(name, "euler_fast")) {
        o.euler_fast = value;
        return 0;
    }
    if (!std::strcmp(name, "outputs")) {
        o.outputs = value;
        return 0;
    }
    struct {
        const char* name;
        int* slot;
    } table[] = {{"p1_tie_desc", &o.opt.p1_tie_desc},
                 {"p1_exclusive_bound", &o.opt.p1_exclusive_bound},
                 {"p2_self_mirror_zero", &o.opt.p2_self_mirror_zero},
                 {"p3_oldest_first", &o.opt.p3_oldest_first},
                 {"p6_bcalm_kmer_numbering", &o.opt.p6_bcalm_kmer_numbering},
                 {"p7_first_root_wins", &o.opt.p7_first_root_wins}};
    for (auto& t : table)
        if (!std::strcmp(name, t.name)) {
            *t.slot = value;
            return 0;
        }
    // registered assumptions without an implemented alternative: only the assumed value is accepted
    if (!std::strcmp(name, "p4_euler_policy") || !std::strcmp(name, "p5_fasta_numbering")) return value == 0 ? 0 : -1;
    return -1;
}

int mto_run_greedy(void* h, int threads) { MTO_TRY(run_greedy(o, (u32)std::max(1, threads))) }

// Named array access: copies nothing, returns pointer + element count.
int mto_get(void* h, const char* name, const void** ptr, size_t* count) {
    Oracle& o = *(Oracle*)h;
    std::string n(name);
#define RET(vec)                     \
    {                                \
        *ptr = (vec).data();         \
        *count = (vec).size();       \
        return 0;                    \
    }
    if (n == "mirror") RET(o.g.mirror)
    if (n == "edge_from") RET(o.g.from)
    if (n == "edge_to") RET(o.g.to)
    if (n == "edge_weight") RET(o.g.weight)
    if (n == "edge_dummy_id") RET(o.g.dummy_id)
    if (n == "edge_forward") RET(o.g.forward)
    if (n == "edge_unitig") RET(o.g.unitig)
    if (n == "out_nodes") RET(o.out_nodes)
    if (n == "in_node_map0") RET(o.in_node_map0)
    if (n == "mult0") RET(o.mult0)
    if (n == "triples") RET(o.triples)
    if (n == "gfa") RET(o.gfa)
    if (n == "fasta") RET(o.fasta)
    if (n == "bitvector") RET(o.bitvec)
    if (n == "c_edge_out") RET(o.c_edge_out)
    if (n == "c_insert_out") RET(o.c_insert_out)
    if (n == "c_limits") RET(o.c_limits)
#undef RET
    return -1;
}

size_t mto_num(void* h, const char* name) {
    Oracle& o = *(Oracle*)h;
    std::string n(name);
    if (n == "nodes") return o.g.node_count();
    if (n == "edges") return o.g.edge_count();
    if (n == "original_edges") return o.n_original_edges;
    if (n == "unitigs") return o.n_original_edges / 2;
    if (n == "walks") return o.walks.size();
    if (n == "cycles") return o.cycles.size();
    if (n == "sources") return o.out_nodes.size();
    if (n == "dijkstra_calls") return o.st.dijkstra_calls;
    if (n == "settled") return o.st.settled;
    if (n == "relaxed") return o.st.relaxed;
    if (n == "heap_pops" || n == "iterations") return o.st.heap_pops;
    if (n == "unnecessary_heap_elements") return o.st.stale_pops;
    if (n == "max_max_heap_size") return o.st.max_max_heap;
    if (n == "sum_max_heap_size") return o.st.sum_max_heap;
    if (n == "max_max_distance_array_size") return o.st.max_max_dist_array;
    if (n == "sum_max_distance_array_size") return o.st.sum_max_dist_array;
    if (n == "breaking_edges") return o.st.breaking_edges;
    return (size_t)-1;
}

double mto_time(void* h, const char* name) {
    Oracle& o = *(Oracle*)h;
    std::string n(name);
    if (n == "parse") return o.st.t_parse;
    if (n == "build") return o.st.t_build;
    if (n == "bitvector") return o.st.t_bitvector;
    if (n == "scan") return o.st.t_scan;
    if (n == "dijkstra") return o.st.t_dijkstra;
    if (n == "insert") return o.st.t_insert;
    if (n == "eulerise") return o.st.t_eulerise;
    if (n == "euler") return o.st.t_euler;
    if (n == "break") return o.st.t_break;
    if (n == "write") return o.st.t_write;
    return -1.0;
}

// Walk i as edge ids.
int mto_walk(void* h, size_t i, const uint32_t** ptr, size_t* count) {
    Oracle& o = *(Oracle*)h;
    if (i >= o.walks.size()) return -1;
    *ptr = o.walks[i].edges.data();
    *count = o.walks[i].edges.size();
    return 0;
}
int mto_cycle(void* h, size_t i, const uint32_t** ptr, size_t* count) {
    Oracle& o = *(Oracle*)h;
    if (i >= o.cycles.size()) return -1;
    *ptr = o.cycles[i].data();
    *count = o.cycles[i].size();
    return 0;
}

// Candidate lists L(src) for sources [lo, hi): writes up to `cap` (node, dist) pairs per source
// into out_nodes/out_dists (row-major, cap per source), full length into out_len.
// Requires a prior mto_run_greedy (phase A fills the initial target map).
int mto_candidates(void* h, size_t lo, size_t hi, size_t cap, uint32_t* out_nodes, uint32_t* out_dists, uint32_t* out_len) {
    MTO_TRY({
        Dijkstra dj(o.g.node_count());
        std::vector<std::pair<u32, u64>> d;
        // the search must not see dummy edges: they are appended after the originals, so mask them by
        // temporarily restoring the pre-insertion adjacency heads.
        Graph& g = o.g;
        std::vector<u32> saved = g.head_out;
        for (u32 v = 0; v < g.node_count(); v++) {
            u32 e = g.head_out[v];
            while (e != NONE && e >= o.n_original_edges) e = g.next_out[e];
            g.head_out[v] = e;
        }
        for (size_t s = lo; s < hi; s++) {
            candidates_of(o, dj, o.out_nodes[s], d);
            out_len[s - lo] = (u32)d.size();
            for (size_t j = 0; j < d.size() && j < cap; j++) {
                out_nodes[(s - lo) * cap + j] = d[j].first;
                out_dists[(s - lo) * cap + j] = (u32)d[j].second;
            }
        }
        g.head_out = saved;
    })
}

}  // extern "C"
