// host_tail.cpp -- the sequential tail of compute_greedytigs that fixes the output order and
// therefore stays on the host (SURVEY.md section 2, rows marked H*):
//
//   C  dummy-edge insertion in result order            greedytigs/mod.rs:678-689
//   D  make_graph_eulerian_with_breaking_edges         src/implementation/mod.rs:392-649
//   E  Eulerian check                                  greedytigs/mod.rs:708-715
//   F  minimum bidirected Eulerian cycle decomposition greedytigs/mod.rs:722 (bigraph)
//   G  rotate to the heaviest dummy + break            greedytigs/mod.rs:726-789
//
// Same results as the reference, different machinery, laid out for the cache instead of for
// generality: D works on degree arrays with two monotone cursors instead of BTreeMaps and runs
// before any adjacency exists; the adjacency is then built once as a CSR whose rows are ordered
// newest edge first (petgraph's iteration order); F keeps the cycle as a circular
// doubly linked list (rotate == move the head), a FIFO of not-yet-exhausted positions instead of
// rescanning the cycle, and per-edge walk records with a used-slot bitset (WalkRec, mtg_internal.cuh), which makes it
// linear in the number of edges with one record line touched per step -- prefetched WALK_DEPTH steps ahead.
#include <sys/mman.h>

#include <algorithm>
#include <omp.h>
#include <sched.h>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "mtg_internal.cuh"

namespace mtg {

// Anonymous mapping advised to use transparent huge pages: the Euler walk is a chain of dependent random
// accesses over tens of megabytes, and with 4 KiB pages most of them also miss the TLB.
void* HugeBuf::ensure(size_t bytes) {
    if (bytes <= cap) return p;
    release();
    const size_t two_mb = size_t(2) << 20;
    size_t want = (bytes + bytes / 8 + two_mb - 1) / two_mb * two_mb;
    void* q = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    MTG_REQUIRE(q != MAP_FAILED, MTG_ERR_INTERNAL, "out of host memory (mmap)");
    madvise(q, want, MADV_HUGEPAGE);
    memset(q, 0, want);  // first touch by the thread that will run the sequential walk: keeps the pages on its NUMA node
    p = q;
    cap = want;
    return p;
}
void* HugeBuf::ensure_pinned(size_t bytes) {
    if (bytes <= cap && pinned) return p;
    ensure(bytes);
    if (!pinned) {
        MTG_CUDA(cudaHostRegister(p, cap, cudaHostRegisterDefault));
        pinned = true;
    }
    return p;
}
void HugeBuf::release() {
    if (p && pinned) cudaHostUnregister(p);
    if (p) munmap(p, cap);
    p = nullptr;
    cap = 0;
    pinned = false;
}

}  // namespace mtg

namespace mtg {

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Threads for the data-parallel parts of the tail: the cores this process may run on (at most 16), NOT OpenMP's default --
// launchers such as torchrun export OMP_NUM_THREADS=1 for every rank, which would serialise the record copy and the piece
// emission of the one rank that runs the tail.  MTG_HOST_THREADS overrides.
int host_threads() {
    static const int n = [] {
        if (const char* e = getenv("MTG_HOST_THREADS")) return std::max(1, atoi(e));
        cpu_set_t set;
        int cpus = 1;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
        return std::max(1, std::min(cpus, 16));
    }();
    return n;
}

struct TailInput {
    u32 k;
    u64 n_nodes, n_orig;  // n_orig = 2U original edges
    const u32 *from, *to, *unitig_w, *mirror;
    const u32* triples;
    u64 n_triples;
    bool oldest_first;  // assumption P3 flipped: out-edges are iterated oldest first
};

struct TailOutput {
    std::vector<u32> walk_edges;
    PinnedVec<u32>* edges_pinned = nullptr;  // if set, the walk edges are written there instead (page-locked: uploaded right after)
    std::vector<u64> walk_limits;
    std::vector<u32> dummy_w;  // weight of dummy edge e at [(e - n_orig)]
    u64 cycles = 0, breaking = 0;
    double ms_degrees = 0, ms_eulerise = 0, ms_csr = 0, ms_walk = 0, ms_break = 0;
};

struct Pair {
    u32 out_node, in_node, w;
};

// Minimal vector-like view over a cached HugeBuf (trivially copyable element types only).
template <class T>
struct HVec {
    T* p = nullptr;
    size_t n = 0;
    void bind(HugeBuf& b, size_t capacity, size_t count, bool zero) {
        p = static_cast<T*>(b.ensure(std::max<size_t>(capacity, 1) * sizeof(T)));
        n = count;
        if (zero && count) memset(p, 0, count * sizeof(T));
    }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    void push_back(const T& v) { p[n++] = v; }
    size_t size() const { return n; }
    void clear() { n = 0; }
    T* data() { return p; }
    T* begin() { return p; }
    T* end() { return p + n; }
};

// D. Pairs the remaining imbalance with breaking edges of weight k.  Needs degrees and mirrors only.
void eulerise(const TailInput& in, const u32* out_deg, const u32* in_deg, HVec<i32>& diff, std::vector<Pair>& pairs) {
    const u32 n = (u32)in.n_nodes;
    const u32* mirror = in.mirror;
    std::vector<u32> outs, ins, selfs;  // ascending node id
    for (u32 v = 0; v < n; v++) {
        if (mirror[v] == v) {
            if (out_deg[v] & 1) selfs.push_back(v);
        } else {
            i32 d = (i32)out_deg[v] - (i32)in_deg[v];
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
            pairs.push_back({selfs[i], selfs[i + 1], in.k});
        } else {
            skip_ins();
            MTG_REQUIRE(ip < ins.size(), MTG_ERR_INTERNAL,
                        "Have an uneven number of self-mirrors, but no other nodes with missing in edges.");
            u32 in_node = ins[ip];
            pairs.push_back({selfs[i], in_node, in.k});
            diff[in_node] -= 1;
            diff[mirror[in_node]] += 1;
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
        if ((in_node == mirror[out_node] && diff[out_node] > -2) || in_node == out_node) {
            size_t q = ip + 1;
            while (q < ins.size() && diff[ins[q]] <= 0) q++;
            MTG_REQUIRE(q < ins.size(), MTG_ERR_INTERNAL, "No further in_nodes left");
            in_node = ins[q];
        }
        const u32 mirror_out_node = mirror[in_node], mirror_in_node = mirror[out_node];
        pairs.push_back({out_node, in_node, in.k});
        diff[out_node] += 1;
        diff[in_node] -= 1;
        if (diff[mirror_out_node] < 0) diff[mirror_out_node] += 1;  // only while it is still listed (:609-627)
        if (diff[mirror_in_node] > 0) diff[mirror_in_node] -= 1;    // (:628-644)
    }
    skip_ins();
    MTG_REQUIRE(ip == ins.size(), MTG_ERR_INTERNAL, "eulerise: in-nodes left over");
}

// What the walk needs (see WalkRec in mtg_internal.cuh): the records, the initial used-slot bitset (padding slots and
// headers of big nodes are marked from the start; the walk marks the rest), node handles and the two slot tables.
struct WalkInput {
    u32 k;
    u64 n_nodes, E0, E, n_slots;
    const WalkRec* recs;
    u64* used;               // bitset over slots, 8 spare bytes behind the last slot
    const u32* handle;       // [N] (host-prepared path; else null and start_handle is set)
    const u32* slot_edge;    // [n_slots] edge id (NONE32 for padding / headers)
    const u32* slot_of_edge; // [E0] slot of every original edge (walk starts)
    const u32* from;         // [E0] from-node of every original edge (host-prepared path)
    const u32* start_handle; // [E0] handle of the from-node of every original edge (device-prepared path)
    const u32* dummy_w;      // weight of dummy edge e at [e - E0]
    bool matching_light;     // every matching dummy weighs less than k (always true behind the GPU matching)
    bool hints;              // the records carry their lookahead levels (not worth building while everything is cache-resident)
};

void walk_and_break(const WalkInput& w, TailOutput& out, TailScratch& scratch);

// Host-side record builder (small graphs, the host-only entry, tests): same records as tail_prep.cu builds on the device.
// `edge_end(e, &from, &to)` yields the end nodes of any original or dummy edge.
template <class EdgeEnds>
void build_walk_records(u32 n, u64 E0, u64 E, u32 k, const u32* out_deg, const u32* mirror, const u32* dummy_w, bool oldest_first,
                        EdgeEnds&& edge_ends, TailScratch& scratch, WalkInput& w) {
    u32* handle = static_cast<u32*>(scratch.handle.ensure(std::max<size_t>(n, 1) * sizeof(u32)));
    u64 n_slots = 0;
    for (u32 v = 0; v < n; v++) {  // binode by binode: a node and its mirror own neighbouring slots (see tail_prep.cu)
        const u32 m = mirror[v];
        if (m < v) continue;
        handle[v] = walk_handle((u32)n_slots, out_deg[v]);
        n_slots += walk_cap(out_deg[v]);
        if (m > v) {
            handle[m] = walk_handle((u32)n_slots, out_deg[m]);
            n_slots += walk_cap(out_deg[m]);
        }
    }
    MTG_REQUIRE(n_slots < SLOT_MASK, MTG_ERR_UNSUPPORTED, "more than 2^30 edge slots");
    WalkRec* recs = static_cast<WalkRec*>(scratch.recs.ensure(std::max<u64>(n_slots, 1) * sizeof(WalkRec)));
    u32* slot_edge = static_cast<u32*>(scratch.slot_edge.ensure(std::max<u64>(n_slots, 1) * sizeof(u32)));
    u32* slot_of_edge = static_cast<u32*>(scratch.slot_of_edge.ensure(std::max<u64>(E, 1) * sizeof(u32)));
    const size_t used_words = n_slots / 64 + 2;
    u64* used = static_cast<u64*>(scratch.used.ensure(used_words * sizeof(u64)));
    memset(used, 0, used_words * sizeof(u64));
    memset(slot_edge, 0xFF, n_slots * sizeof(u32));
    auto mark = [&](u64 s) { used[s >> 6] |= 1ull << (s & 63); };
    // placing the edges in descending id order leaves every node's slots in iteration order (newest edge first);
    // `fill` = next free entry slot per node, kept in the (not yet needed) record array of the node's first slot
    std::vector<u32> fill(n);
    for (u32 v = 0; v < n; v++) {
        const u32 d = out_deg[v], base = handle[v] & H_BASE, cap = walk_cap(d);
        fill[v] = walk_first_slot(handle[v]);
        if (d > 4) {
            recs[base].to = d;  // header of a big node
            mark(base), mark(base + 1);
        }
        for (u32 j = fill[v] - base + d; j < cap; j++) mark(base + j);  // padding
    }
    for (u64 q = 0; q < E; q++) {
        const u64 e = oldest_first ? q : E - 1 - q;
        u32 f, t;
        edge_ends((u32)e, &f, &t);
        const u32 sl = fill[f]++;
        slot_edge[sl] = (u32)e;
        slot_of_edge[e] = sl;
        recs[sl].to = handle[t];
    }
    const bool par = n_slots > (1u << 18);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (i64 sl = 0; sl < (i64)n_slots; sl++) {
        const u32 e = slot_edge[sl];
        if (e == NONE32) continue;
        u32 m = slot_of_edge[e ^ 1u];
        if (e >= E0) m |= SLOT_DUMMY | (dummy_w[e - E0] >= k ? SLOT_BREAK : 0u);
        recs[sl].mslot = m;
    }
    // hint levels, one after the other; degree of the target = what its handle and header say
    auto deg_of_handle = [&](u32 h) -> u32 {
        if (h & H_BIG) return recs[h & H_BASE].to;
        const u32 base = h & H_BASE, cap = (h & H_FOUR) ? 4u : 2u;
        u32 d = 0;
        for (u32 j = 0; j < cap; j++) d += slot_edge[base + j] != NONE32;
        return d;
    };
    // The lookahead levels only pay off once the records outgrow the caches; below that (every graph the library prepares
    // on the host) building them would cost more than the walk itself.
    w.hints = n_slots * sizeof(WalkRec) > (64u << 20) || getenv("MTG_TAIL_FORCEHINT");
    for (u32 level = 2; w.hints && level <= WALK_DEPTH; level++) {
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
        for (i64 sl = 0; sl < (i64)n_slots; sl++)
            if (slot_edge[sl] != NONE32) walk_fill_hints(recs, (u32)sl, deg_of_handle(recs[sl].to), level);
    }
    w.n_slots = n_slots;
    w.recs = recs;
    w.used = used;
    w.handle = handle;
    w.slot_edge = slot_edge;
    w.slot_of_edge = slot_of_edge;
}

void run_tail(const TailInput& in, TailOutput& out, TailScratch& scratch) {
    const u32 n = (u32)in.n_nodes;
    const u64 E0 = in.n_orig;
    double t0 = now_ms();
    // ---- degrees after the matching dummies (C) ----
    HVec<u32> out_deg, in_deg;
    HVec<i32> diff;
    out_deg.bind(scratch.out_deg, n, n, true);
    in_deg.bind(scratch.in_deg, n, n, true);
    diff.bind(scratch.diff, n, n, true);
    // Counting is order-independent, so it runs on all host cores for big inputs (relaxed atomic increments); only the
    // pairing loop and the walk are inherently sequential.  Small graphs stay on one thread, where a plain increment
    // is ~10x cheaper than a locked one.
    const bool par = E0 > (1u << 18);
    auto bump = [par](u32& x) {
        if (par) return __atomic_fetch_add(&x, 1u, __ATOMIC_RELAXED);
        return x++;
    };
    u32* od = out_deg.data();
    u32* id = in_deg.data();
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (i64 e = 0; e < (i64)E0; e++) {
        bump(od[in.from[e]]);
        bump(id[in.to[e]]);
    }
    std::vector<Pair> pairs(in.n_triples);
    pairs.reserve(in.n_triples + 1024);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (i64 j = 0; j < (i64)in.n_triples; j++) {
        const u32 o = in.triples[3 * j], i = in.triples[3 * j + 1];
        pairs[j] = {o, i, in.triples[3 * j + 2]};
        bump(od[o]);
        bump(id[i]);
        bump(od[in.mirror[i]]);
        bump(id[in.mirror[o]]);
    }
    double t1 = now_ms();
    // ---- D ----
    eulerise(in, out_deg.data(), in_deg.data(), diff, pairs);
    for (size_t j = in.n_triples; j < pairs.size(); j++) {
        out_deg[pairs[j].out_node]++;
        in_deg[pairs[j].in_node]++;
        out_deg[in.mirror[pairs[j].in_node]]++;
        in_deg[in.mirror[pairs[j].out_node]]++;
    }
    // ---- E ----
    for (u32 v = 0; v < n; v++) {
        const bool ok = in.mirror[v] == v ? !(out_deg[v] & 1) : out_deg[v] == in_deg[v];
        MTG_REQUIRE(ok, MTG_ERR_INTERNAL, "Failed to make the graph Eulerian.");
    }
    double t2 = now_ms();
    // ---- walk records.  Dummy pair j = edges E0+2j (out->in) and E0+2j+1 (mirror(in)->mirror(out)) ----
    const u64 E = E0 + 2 * pairs.size();
    MTG_REQUIRE(E < SLOT_MASK, MTG_ERR_UNSUPPORTED, "more than 2^30 edges");
    out.dummy_w.resize(2 * pairs.size());
    u32 max_matching_w = 0;
    for (size_t j = 0; j < pairs.size(); j++) {
        out.dummy_w[2 * j] = out.dummy_w[2 * j + 1] = pairs[j].w;
        if (j < in.n_triples) max_matching_w = std::max(max_matching_w, pairs[j].w);
    }
    WalkInput w{in.k, in.n_nodes, E0, E, 0, nullptr, nullptr, nullptr, nullptr, nullptr, in.from, nullptr, out.dummy_w.data(), max_matching_w < in.k,
                false};
    build_walk_records(n, E0, E, in.k, od, in.mirror, out.dummy_w.data(), in.oldest_first,
                       [&](u32 e, u32* f, u32* t) {
                           if (e < E0) {
                               *f = in.from[e], *t = in.to[e];
                           } else {
                               const Pair& p = pairs[(e - E0) >> 1];
                               if ((e - E0) & 1) *f = in.mirror[p.in_node], *t = in.mirror[p.out_node];
                               else *f = p.out_node, *t = p.in_node;
                           }
                       },
                       scratch, w);
    double t3 = now_ms();
    out.ms_degrees = t1 - t0;
    out.ms_eulerise = t2 - t1;
    out.ms_csr = t3 - t2;
    walk_and_break(w, out, scratch);
}

// ---- F: one run of the walk ----
// Bit helpers over the used-slot bitset.  Loads and marks use the same aligned 64-bit words, so a load right behind a
// mark of the same word is served by store forwarding (a byte store under a wider load is not).
static inline u32 slot_bits_of(const u64* used, u32 base) {
    const u32 sh = base & 63;
    u64 x = used[base >> 6] >> sh;
    if (__builtin_expect(sh > 60, 0)) x |= used[(base >> 6) + 1] << (64 - sh);  // four slots starting at bit 62
    return (u32)x;
}
static inline u32 small_mask(u32 h) { return 3u + 12u * (h & H_FOUR); }  // two or four slots (H_FOUR == 1)
static_assert(H_FOUR == 1u, "small_mask relies on the flag being bit 0");

// What a run reads and appends to.  Kept apart from walk_and_break so that the hot loop is a small function whose state
// the compiler keeps in registers (inside the big function most of it lived on the stack).
struct RunIO {
    const WalkRec* recs;
    u64* used;
    u32 *q_slot, *q_from;  // the cycle under construction (same length)
    size_t q_n;
    u32* cand;             // positions whose from-node had slots left when the walk passed
    size_t cand_n;
    int n_sources;         // records used as hint sources (A/B switch)
    bool fast_path;        // A/B switch
    // diagnostics (DIAG instantiation only)
    u64 dg_steps, dg_reset, dg_have[8], dg_nohint, dg_big, dg_end;
    u64 probe_hist[5][40];  // 16-tick buckets; [4] = the probe itself on a line that is certainly in L1
    u64 probe_tick;
    bool probe;
    u32 spin;
    u64 spin_sink;
};

// first unused slot of the node with handle h (NONE32: exhausted); *more = the node owns further unused slots behind it
static inline u32 first_unused_of(const WalkRec* recs, const u64* used, u32 h, bool* more) {
    const u32 base = h & H_BASE;
    if (__builtin_expect(!(h & H_BIG), 1)) {
        const u32 free_bits = ~slot_bits_of(used, base) & small_mask(h);
        *more = (free_bits & (free_bits - 1)) != 0;
        return free_bits ? base + (u32)__builtin_ctz(free_bits) : NONE32;
    }
    const u32 d = recs[base].to;
    u32 first = NONE32;
    *more = false;
    for (u32 s = base + 2; s < base + 2 + d; s++)
        if (!((used[s >> 6] >> (s & 63)) & 1u)) {
            if (first == NONE32) first = s;
            else {
                *more = true;
                break;
            }
        }
    return first;
}

// Follows first-unused out-edges from slot `s` (leaving the node with handle from_h) until the walk is stuck.
// HINTS: the records carry their lookahead levels.  PF: prefetch hint (0 = t0, 1 = nta, 2 = t2).  NTS: the queues are
// appended with non-temporal stores.  DIAG: statistics and latency probes.
// (No SLP vectorisation here: the compiler would fuse the fixed-length shift of the chain arrays into 16-byte loads that
// overlap the 4-byte stores of the step before -- a store-forwarding failure on the loop-carried path of every step,
// each one waiting for the store buffer to drain.)
template <bool HINTS, int PF, bool NTS, bool DIAG>
__attribute__((noinline, optimize("no-tree-slp-vectorize", "no-tree-vectorize"))) static void walk_run(RunIO& io, u32 s, u32 from_h) {
    const WalkRec* const recs = io.recs;
    u64* const used = io.used;
    u32* const q_slot = io.q_slot;
    u32* const q_from = io.q_from;
    u32* const cand = io.cand;
    size_t q_n = io.q_n, cand_n = io.cand_n;
    const int n_sources = io.n_sources;
    const bool fast_path = io.fast_path;
    auto mark = [&](u32 x) { used[x >> 6] |= 1ull << (x & 63); };
    auto prefetch_rec = [&](const void* p) {
        if (PF == 1) __builtin_prefetch(p, 0, 0);
        else if (PF == 2) __builtin_prefetch(p, 0, 1);
        else __builtin_prefetch(p, 0, 3);
    };
    auto append = [&](u32* v, size_t at, u32 x) {
#if defined(__x86_64__)
        if (NTS) {
            __builtin_ia32_movnti(reinterpret_cast<int*>(v + at), (int)x);
            return;
        }
#endif
        v[at] = x;
    };
    // the slot the node with handle h would hand out next, if the walk can tell without its records (small nodes)
    auto peek = [&](u32 h, u32* j) -> u32 {
        if (__builtin_expect((h & H_BIG) != 0, 0)) return NONE32;
        const u32 base = h & H_BASE;
        const u32 free_bits = ~slot_bits_of(used, base) & small_mask(h);
        if (__builtin_expect(!free_bits, 0)) return NONE32;
        *j = (u32)__builtin_ctz(free_bits);
        return base + *j;
    };
    auto probe_load = [&](int which, const volatile void* addr, int bytes) {
#if defined(__x86_64__)
        unsigned aux;
        _mm_lfence();
        const u64 a = __rdtscp(&aux);
        _mm_lfence();
        if (bytes == 8) (void)*(const volatile u64*)addr;
        else (void)*(const volatile u32*)addr;
        _mm_lfence();
        const u64 b = __rdtscp(&aux);
        io.probe_hist[which][std::min<u64>((b - a) >> 4, 39)]++;
#endif
    };
    // What the walk expects 1 .. `have` steps from now -- P[L] = slot, Q[L] = its index inside its node -- carried from
    // step to step: as long as the next slot is the one expected, everything moves one step closer and only the far end
    // of the chain has to be extended (normally by one level: one look at the used bits, one prefetch).  The chain ends
    // at a node that will be left through its third or fourth slot (the records follow two slots per node behind their
    // own target) and grows again from the record of the next step, or once that node is the next one.
    u32 P[WALK_DEPTH + 4] = {}, Q[WALK_DEPTH + 4] = {};
    u32 have = 1;  // deepest level known
    // Extends the chain from the hints of record `src`, whose own slot sits `b` steps ahead (b = 0: the record in hand,
    // b = 1: the record of the next step).  Level L' of src's hints is level b + L' from here, indexed by the slot choices
    // Q[b + 1] (at src.to; 0 .. 3) and Q[b + 2 ..] (0 .. 1) on the way.
    auto extend = [&](const WalkRec& src, u32 b) {
        const u32 cc = src.to;
        if (cc & H_BIG) return;
        if (have < b + 1) {  // (b >= 1 only) the slot behind src straight from src.to
            u32 j = 0;
            const u32 sl = peek(cc, &j);
            if (sl == NONE32) return;
            prefetch_rec(&recs[sl]);
            P[b + 1] = sl, Q[b + 1] = j, have = b + 1;
        }
        const bool four = (cc & H_FOUR) != 0;
        const u32 last = b + (four ? WALK_DEPTH - 1 : WALK_DEPTH);  // deepest level this record knows
        u32 idx = Q[b + 1];
        bool open_end = true;
        for (u32 M = b + 2; M <= have; M++) {
            open_end = Q[M] < 2;
            idx = 2 * idx + Q[M];
        }
        for (u32 L = have + 1; open_end && L <= last; L++) {
            u32 j = 2;
            const u32 sl = peek(src.h[(four ? walk_level_four(L - b) : walk_level_two(L - b)) + idx], &j);
            if (sl == NONE32) break;
            prefetch_rec(&recs[sl]);
            if (WALK_DEPTH >= 5) prefetch_rec(reinterpret_cast<const char*>(&recs[sl]) + 64);
            P[L] = sl, Q[L] = j, have = L;
            open_end = j < 2;
            idx = 2 * idx + j;
        }
    };
    while (s != NONE32) {
        const WalkRec& r = recs[s];
        const bool probing = DIAG && io.probe && (++io.probe_tick & 63) == 0;
        if (probing) {
            probe_load(4, &io.probe_tick, 8);
            probe_load(0, &r.to, 4);
        }
        const u32 ms = r.mslot;
        if (probing) probe_load(3, &used[(ms & SLOT_MASK) >> 6], 8);
        mark(s);
        mark(ms & SLOT_MASK);
        append(q_slot, q_n, s | (ms & ~SLOT_MASK));
        append(q_from, q_n, from_h);
        q_n++;
        const u32 c = r.to;
        if (probing) probe_load(1, &used[(c & H_BASE) >> 6], 8);
        bool more;
        const u32 nxt = first_unused_of(recs, used, c, &more);  // the next step is certain
        // (branch-free: whether a node still has a second unused slot is a coin flip the predictor cannot learn)
        cand[cand_n] = (u32)q_n;
        cand_n += more;
        if (HINTS && nxt != NONE32 && !(c & H_BIG)) {
            // Slot choice at `to` (Q[1]).  When the next slot is the expected one the choice comes from the chain, not from
            // `nxt`: the extension (hint -> used bits of the far node -> prefetch) then does not wait for the used bits
            // of `to`, it runs ahead under the predicted branch.
            if (__builtin_expect(have >= 2 && nxt == P[2], 1)) {  // as expected: everything moves one step closer
#pragma GCC unroll 8
                for (u32 M = 1; M < WALK_DEPTH + 2; M++) P[M] = P[M + 1], Q[M] = Q[M + 1];  // fixed length: no loop branch
                have--;
            } else {
                P[1] = nxt, Q[1] = nxt - (c & H_BASE);
                prefetch_rec(&recs[nxt]);
                have = 1;
                if (DIAG) io.dg_reset++;
            }
            u32 open_bits = 0;
#pragma GCC unroll 8
            for (u32 M = 2; M < WALK_DEPTH; M++) open_bits |= Q[M];
            if (__builtin_expect(fast_path && have == WALK_DEPTH - 1 && !(c & H_FOUR) && open_bits < 2, 1)) {
                // The steady state, written out: the chain lacks exactly its deepest level, the record in hand has the
                // two-slot layout and the path stays on first/second slots.  One hint, one look at the used bits, one
                // prefetch -- no loops whose trip counts the branch predictor would have to guess.
                u32 idx = 0;
#pragma GCC unroll 8
                for (u32 M = 1; M < WALK_DEPTH; M++) idx = 2 * idx + Q[M];
                u32 j = 0;
                const u32 hh = r.h[walk_level_two(WALK_DEPTH) + idx];
                if (probing) probe_load(2, &used[(hh & H_BASE) >> 6], 8);
                const u32 sl = peek(hh, &j);
                if (sl != NONE32) {
                    prefetch_rec(&recs[sl]);
                    if (WALK_DEPTH >= 5) prefetch_rec(reinterpret_cast<const char*>(&recs[sl]) + 64);  // 128-byte records: both lines
                    P[WALK_DEPTH] = sl, Q[WALK_DEPTH] = j, have = WALK_DEPTH;
                }
            } else {
                extend(r, 0);
                // The record of the next step was asked for several steps ago and has normally arrived: its hints go on where
                // this record's stop (behind a node with three or four slots, which the record in hand only follows when
                // that node is its own target; one level short when its own target has four slots).
                if (n_sources >= 2 && (have < WALK_DEPTH || !fast_path)) extend(recs[nxt], 1);
                if (n_sources >= 3 && have >= 2) extend(recs[P[2]], 2);  // experiment: may not have arrived yet
            }
            if (DIAG) io.dg_have[have & 7]++;
        } else {
            if (nxt != NONE32) prefetch_rec(&recs[nxt]);
            have = 1;
            if (DIAG) {
                if (nxt == NONE32) io.dg_end++;
                else if (c & H_BIG) io.dg_big++;
                else io.dg_nohint++;
            }
        }
        if (DIAG) {
            io.dg_steps++;
            // experiment (MTG_WALK_SPIN=n): n dependent multiplies per step, off the walk's own dependency chain -- does extra
            // core work lengthen a step (core-bound) or disappear in the wait for memory?
            u64 x = io.probe_tick | 1;
            for (u32 i = 0; i < io.spin; i++) x = x * 0x9E3779B97F4A7C15ull + i;
            io.spin_sink += x;
        }
        from_h = c;
        s = nxt;
    }
    io.q_n = q_n;
    io.cand_n = cand_n;
}

// The run loop used by default (MTG_WALK_CHAIN=carried selects walk_run above): the same carried
// chain, kept in a handful of scalars so that the steady state lives in registers -- the slots expected 2 .. WALK_DEPTH
// steps from now (p[]), the slot choices on the way as the low bits of one word (they ARE the index of the deepest
// hint), and one flag saying that the chain is complete and regular (two-slot layout, first/second slots all the way).
// A step of the steady state reads one hint, looks at the used bits of one far node and issues one prefetch.  Whenever
// the chain does not hold (a run start, a target with three or four slots, a third or fourth slot on the way, a big
// node, a slot used up in between) the step re-derives all levels from the record in hand, one look at the used bits
// per level, and prefetches what it finds; the steady state resumes as soon as that yields a full regular chain.
// (walk_run shifts two arrays through the stack every step: about 110 instructions per step against about 60 here; the
// perfect-lookahead replay shows what they cost: every instruction of a step is paid once per step, T = L / d + w.)
template <int PF, bool NTS, bool DIAG>
__attribute__((noinline, optimize("no-tree-slp-vectorize", "no-tree-vectorize"))) static void walk_run_lean(RunIO& io, u32 s, u32 from_h) {
    const WalkRec* const recs = io.recs;
    u64* const used = io.used;
    u32* const q_slot = io.q_slot;
    u32* const q_from = io.q_from;
    u32* const cand = io.cand;
    size_t q_n = io.q_n, cand_n = io.cand_n;
    auto mark = [&](u32 x) { used[x >> 6] |= 1ull << (x & 63); };
    // (DIAG) how many steps before its use a record was asked for: the last 32 prefetches, looked up at every step
    const WalkRec* pf_rec[32] = {};
    u64 pf_step[32] = {};
    u32 pf_n = 0;
    auto prefetch_rec = [&](const void* p) {
        if (PF == 1) __builtin_prefetch(p, 0, 0);
        else if (PF == 2) __builtin_prefetch(p, 0, 1);
        else __builtin_prefetch(p, 0, 3);
        if (WALK_DEPTH >= 5) {
            if (PF == 1) __builtin_prefetch(static_cast<const char*>(p) + 64, 0, 0);
            else __builtin_prefetch(static_cast<const char*>(p) + 64, 0, 3);
        }
        if (DIAG) {
            bool known = false;
            for (u32 x = 0; x < 32; x++) known |= pf_rec[x] == p;
            if (!known) pf_rec[pf_n & 31] = static_cast<const WalkRec*>(p), pf_step[pf_n & 31] = io.dg_steps, pf_n++;
        }
    };
    auto append = [&](u32* v, size_t at, u32 x) {
#if defined(__x86_64__)
        if (NTS) {
            __builtin_ia32_movnti(reinterpret_cast<int*>(v + at), (int)x);
            return;
        }
#endif
        v[at] = x;
    };
    auto peek = [&](u32 h, u32* j) -> u32 {
        if (__builtin_expect((h & H_BIG) != 0, 0)) return NONE32;
        const u32 base = h & H_BASE;
        const u32 free_bits = ~slot_bits_of(used, base) & small_mask(h);
        if (__builtin_expect(!free_bits, 0)) return NONE32;
        *j = (u32)__builtin_ctz(free_bits);
        return base + *j;
    };
    constexpr u32 IDX_MASK = (1u << (WALK_DEPTH - 1)) - 1u;
    // p[0 .. deep): the slots expected for the next `deep` steps; the low `deep` bits of `path`: the slot choices that lead
    // to them (0 / 1 each while the chain is regular) == the index of the hint one level further
    u32 p[WALK_DEPTH + 1] = {};
    u32 path = 0, deep = 0;
    bool regular = false;
    while (s != NONE32) {
        const WalkRec& r = recs[s];
        if (DIAG) {
            u64 lead = 0;
            for (u32 x = 0; x < 32; x++)
                if (pf_rec[x] == &r) lead = io.dg_steps - pf_step[x];
            io.probe_hist[0][std::min<u64>(lead, 39)]++;
        }
        const u32 ms = r.mslot;
        mark(s);
        mark(ms & SLOT_MASK);
        append(q_slot, q_n, s | (ms & ~SLOT_MASK));
        append(q_from, q_n, from_h);
        q_n++;
        const u32 c = r.to;
        bool more;
        const u32 nxt = first_unused_of(recs, used, c, &more);  // the next step is certain
        cand[cand_n] = (u32)q_n;
        cand_n += more;
        if (__builtin_expect(regular && nxt == p[0] && !(c & H_BIG), 1)) {
            if (__builtin_expect(deep == WALK_DEPTH - 1 && !(c & H_FOUR), 1)) {
                // steady state: the far end grows by one level, everything moves one step closer
                u32 j = 0;
                const u32 sl = peek(r.h[walk_level_two(WALK_DEPTH) + (path & IDX_MASK)], &j);
#pragma GCC unroll 8
                for (u32 M = 0; M + 2 < WALK_DEPTH; M++) p[M] = p[M + 1];
                if (__builtin_expect(sl != NONE32, 1)) {
                    prefetch_rec(&recs[sl]);
                    p[WALK_DEPTH - 2] = sl;
                    path = 2 * path + j;
                    regular = j < 2;  // a third or fourth slot ends what the records behind it follow
                } else {
                    regular = false;
                }
            } else {
                // behind a target with three or four slots (whose layout ends one level earlier) the chain is one level short
                // and catches up from the next record: as many levels as this record knows beyond the chain
                // (`p` itself is only ever indexed by constants, so that it lives in registers: the variable part works on a copy)
                u32 pa[WALK_DEPTH + 2];
#pragma GCC unroll 8
                for (u32 M = 0; M <= WALK_DEPTH; M++) pa[M] = p[M];
                const bool four = (c & H_FOUR) != 0;
                const u32 last = four ? WALK_DEPTH - 1 : WALK_DEPTH;
                while (regular && deep < last) {
                    u32 j = 0;
                    const u32 L = deep + 1;
                    const u32 sl = peek(r.h[(four ? walk_level_four(L) : walk_level_two(L)) + (path & ((1u << deep) - 1u))], &j);
                    if (sl == NONE32) {
                        regular = false;
                        break;
                    }
                    prefetch_rec(&recs[sl]);
                    pa[deep++] = sl;
                    path = 2 * path + j;
                    regular = j < 2;
                }
#pragma GCC unroll 8
                for (u32 M = 0; M < WALK_DEPTH; M++) p[M] = pa[M + 1];
                deep--;
                if (deep == 0) regular = false;
            }
        } else if (nxt != NONE32) {
            // all levels again from the record in hand
            if (DIAG) io.dg_reset++, io.dg_big += (c & H_FOUR) != 0;
            regular = false;
            prefetch_rec(&recs[nxt]);
            if (!(c & H_BIG)) {
                const bool four = (c & H_FOUR) != 0;
                const u32 last = four ? WALK_DEPTH - 1 : WALK_DEPTH;
                u32 idx = nxt - (c & H_BASE);  // slot choice at `to`
                bool open = true;  // (idx may be 2 or 3 here: the four-slot layout is indexed by it, and the steps behind no longer need it)
                u32 found = 0;  // levels 2 .. found + 1: the slots of the `found` steps behind the next one
                u32 pa[WALK_DEPTH + 1] = {};
                for (u32 L = 2; open && L <= last; L++) {
                    u32 j = 0;
                    const u32 sl = peek(r.h[(four ? walk_level_four(L) : walk_level_two(L)) + idx], &j);
                    if (sl == NONE32) {
                        open = false;
                        break;
                    }
                    prefetch_rec(&recs[sl]);
                    pa[found++] = sl;
                    open = j < 2;
                    idx = 2 * idx + j;
                }
#pragma GCC unroll 8
                for (u32 M = 0; M < WALK_DEPTH; M++) p[M] = pa[M];
                deep = found;
                path = idx;
                regular = open && found > 0;
            }
        }
        if (DIAG) io.dg_steps++, io.dg_have[regular ? WALK_DEPTH : 1]++;
        from_h = c;
        s = nxt;
    }
    io.q_n = q_n;
    io.cand_n = cand_n;
}

// Diagnostic (MTG_WALK_REPLAY="mode:depth,..."): the finished walk is run again from its own slot sequence, with the
// record `depth` steps ahead prefetched from that sequence -- the walk loop with a perfect lookahead of any depth, which
// separates what the loop itself costs from what the hint chain costs.  mode bits: 1 = the prefetch address depends on
// the record in hand (as a hint does), 2 = and on the used-bit word of the far slot, 4 = no used-bit work at all,
// 8 = no queue appends, 16 = the slot of a step depends on the result of the step before (the walk's own dependence),
// 32 = the prefetch is the first thing a step does.  Returns the number of steps whose first-unused slot differed from the sequence (must be 0).
template <int PF>
__attribute__((noinline, optimize("no-tree-slp-vectorize", "no-tree-vectorize"))) static u64 walk_replay(
    const WalkRec* recs, u64* used, const u32* seq, size_t b, size_t e, u32 from_h, u32 depth, int mode, u32 zero, u32* q_slot, u32* q_from,
    u32* cand) {
    size_t q_n = b, cand_n = b;
    u64 wrong = 0;
    auto mark = [&](u32 x) { used[x >> 6] |= 1ull << (x & 63); };
    u32 carried = 0;  // mode 16: the slot of a step depends on what the step before found, as in the walk itself
    for (size_t i = b; i < e; i++) {
        const u32 s = seq[i] ^ carried;
        const WalkRec& r = recs[s];
        const u32 ms = r.mslot;
        const u32 c = r.to;
        if (mode & 32) {  // the prefetch before everything else
            u32 x = seq[i + depth];
            if (mode & 1) x ^= c & zero;
            if (mode & 2) x ^= (u32)used[x >> 6] & zero;
            if (PF == 1) __builtin_prefetch(&recs[x], 0, 0);
            else __builtin_prefetch(&recs[x], 0, 3);
        }
        if (!(mode & 4)) {
            mark(s);
            mark(ms & SLOT_MASK);
        }
        if (!(mode & 8)) {
#if defined(__x86_64__)
            __builtin_ia32_movnti(reinterpret_cast<int*>(q_slot + q_n), (int)(s | (ms & ~SLOT_MASK)));
            __builtin_ia32_movnti(reinterpret_cast<int*>(q_from + q_n), (int)from_h);
#else
            q_slot[q_n] = s | (ms & ~SLOT_MASK), q_from[q_n] = from_h;
#endif
            q_n++;
        }
        if (!(mode & 4)) {
            bool more;
            const u32 nxt = first_unused_of(recs, used, c, &more);
            cand[cand_n] = (u32)q_n;
            cand_n += more;
            wrong += (i + 1 < e) & (nxt != seq[i + 1]);
            if (mode & 16) carried = nxt & zero;
        } else if (mode & 16) {
            carried = c & zero;
        }
        if (mode & 32) {
            from_h = c;
            continue;
        }
        u32 x = seq[i + depth];
        if (mode & 1) x ^= c & zero;
        if (mode & 2) x ^= (u32)used[x >> 6] & zero;
        if (PF == 1) __builtin_prefetch(&recs[x], 0, 0);
        else __builtin_prefetch(&recs[x], 0, 3);
        if (sizeof(WalkRec) > 64) {  // 128-byte records: both lines, and the loop reads the second one as the walk does
            __builtin_prefetch(reinterpret_cast<const char*>(&recs[x]) + 64, 0, PF == 1 ? 0 : 3);
            wrong += r.h[WALK_HINTS - 1] == 0xFFFFFFF0u;
        }
        from_h = c;
    }
    return wrong;
}

using WalkRunFn = void (*)(RunIO&, u32, u32);
template <bool HINTS, int PF>
static WalkRunFn pick_walk_run(bool nts, bool diag) {
    if (diag) return nts ? walk_run<HINTS, PF, true, true> : walk_run<HINTS, PF, false, true>;
    return nts ? walk_run<HINTS, PF, true, false> : walk_run<HINTS, PF, false, false>;
}
static WalkRunFn pick_walk_run(bool hints, int pf, bool nts, bool diag) {
    if (!hints) return pick_walk_run<false, 0>(nts, diag);
    if (pf == 1) return pick_walk_run<true, 1>(nts, diag);
    if (pf == 2) return pick_walk_run<true, 2>(nts, diag);
    return pick_walk_run<true, 0>(nts, diag);
}

// ---- F + G ----
void walk_and_break(const WalkInput& in, TailOutput& out, TailScratch& scratch) {
    const u64 E0 = in.E0, E = in.E;
    const WalkRec* const recs = in.recs;
    u64* const used = in.used;
    double t3 = now_ms();
    auto is_used = [&](u32 x) { return (u32)(used[x >> 6] >> (x & 63)) & 1u; };
    // The cycle under construction: element i = (q_slot[i] = slot | flags, q_from[i] = handle of the node it leaves).
    // Every extension appends a contiguous run (creation order == the order in which positions are scanned for leftover
    // out-edges).  A run created while position i was the head sits, in cycle order, immediately before element i
    // ("push_back on the rotated vector"); all runs created at i are consecutive, so i owns one block [begin, end) of later
    // elements.  Heads are visited in ascending position, so `children` stays sorted.  Cycle order == in-order expansion
    // of this tree -- a short list of slices of q_slot, because re-roots are rare (about 100 per million steps).
    HVec<u32> q_slot, q_from;
    q_slot.bind(scratch.queue, E / 2 + 16, 0, false);
    q_from.bind(scratch.cyc, E / 2 + 16, 0, false);
    struct Child {
        u32 pos, begin, end;
    };
    std::vector<Child> children;
    struct Slice {
        u32 b, e;
    };
    std::vector<Slice> rope, order;
    struct Frame {
        u32 next, end;   // pending slice start / end of the block
        size_t ci, ce;   // children of this block: children[ci, ce)
    };
    std::vector<Frame> stack;
    // Records are fetched with the non-temporal hint and the queues appended with non-temporal stores: neither is touched
    // again during the walk, and kept out of L2 they leave it to the used-slot bitset, whose words sit on the path from
    // "record in hand" to "next prefetch issued" (bench boxes, chr1: prefetchnta 37.3 against 40.6 ns per step with
    // prefetcht0; the pangenome, whose bitset is a third of the size, does not care).  MTG_WALK_PREFETCH=t0|2,
    // MTG_WALK_NTSTORE=0, MTG_WALK_FAST=0, MTG_WALK_SOURCES=n and MTG_TAIL_NOHINT=1 are A/B switches; MTG_TRACE /
    // MTG_WALK_PROBE / MTG_WALK_SPIN select the instantiation with statistics, latency probes and the spin experiment.
    const bool use_hints = in.hints && !getenv("MTG_TAIL_NOHINT");
    const char* pf_env = getenv("MTG_WALK_PREFETCH");
    const bool beyond_caches = in.n_slots * sizeof(WalkRec) > (size_t(64) << 20);  // (small record arrays stay cached: keep them there)
    const int pf_kind = pf_env ? (pf_env[0] == 'n' ? 1 : pf_env[0] == '2' ? 2 : 0) : (beyond_caches ? 1 : 0);
    const char* nts_env = getenv("MTG_WALK_NTSTORE");
    const bool nt_store = !(nts_env && nts_env[0] == '0');
    const bool diag = trace_slow_calls() || getenv("MTG_WALK_PROBE") != nullptr || getenv("MTG_WALK_SPIN") != nullptr;
    WalkRunFn run = pick_walk_run(use_hints, pf_kind, nt_store, diag);
    bool lean_loop = false;
    const char* chain_env = getenv("MTG_WALK_CHAIN");  // A/B: carried = walk_run (also used for the probes), default = walk_run_lean
    const bool probes = getenv("MTG_WALK_PROBE") != nullptr || getenv("MTG_WALK_SPIN") != nullptr;  // only walk_run has them
    if (use_hints && !probes && !(chain_env && chain_env[0] == 'c')) {
        lean_loop = true;
        if (diag) run = pf_kind == 1 ? walk_run_lean<1, true, true> : walk_run_lean<0, true, true>;
        else if (pf_kind == 1) run = nt_store ? walk_run_lean<1, true, false> : walk_run_lean<1, false, false>;
        else if (pf_kind == 2) run = nt_store ? walk_run_lean<2, true, false> : walk_run_lean<2, false, false>;
        else run = nt_store ? walk_run_lean<0, true, false> : walk_run_lean<0, false, false>;
    }

    // MTG_WALK_REPLAY: keep the initial used bits and every run's slot sequence for the replay diagnostic below
    const char* replay_env = getenv("MTG_WALK_REPLAY");
    std::vector<u64> replay_used0;
    std::vector<u32> replay_seq;
    struct ReplayRun {
        size_t b, e;
        u32 from_h;
    };
    std::vector<ReplayRun> replay_runs;
    if (replay_env) replay_used0.assign(used, used + in.n_slots / 64 + 2);
    RunIO io{};
    io.recs = recs;
    io.used = used;
    io.n_sources = getenv("MTG_WALK_SOURCES") ? atoi(getenv("MTG_WALK_SOURCES")) : 2;
    io.fast_path = !(getenv("MTG_WALK_FAST") && getenv("MTG_WALK_FAST")[0] == '0');
    io.probe = getenv("MTG_WALK_PROBE") != nullptr;
    io.spin = getenv("MTG_WALK_SPIN") ? (u32)atoi(getenv("MTG_WALK_SPIN")) : 0;
    struct DiagAtExit {
        const RunIO* io;
        bool on, lean;
        ~DiagAtExit() {
            if (!on || io->dg_steps < 100000) return;
            const u64* h = io->dg_have;
            if (lean)
                fprintf(stderr, "[mtg trace] walk (lean loop): %llu steps, chain complete and regular after %llu of them, %llu steps re-derived all "
                        "levels from the record in hand (%llu of them at a target with three or four slots)\n", (unsigned long long)io->dg_steps,
                        (unsigned long long)h[WALK_DEPTH], (unsigned long long)io->dg_reset, (unsigned long long)io->dg_big);
            else
            fprintf(stderr, "[mtg trace] walk: %llu steps, chain known to level 1/2/3/4/5 after a step: %llu/%llu/%llu/%llu/%llu, unexpected "
                    "next slot %llu, big-node steps %llu, run ends %llu, without hints %llu\n", (unsigned long long)io->dg_steps,
                    (unsigned long long)h[1], (unsigned long long)h[2], (unsigned long long)h[3], (unsigned long long)h[4], (unsigned long long)h[5],
                    (unsigned long long)io->dg_reset, (unsigned long long)io->dg_big, (unsigned long long)io->dg_end, (unsigned long long)io->dg_nohint);
            if (!io->probe) {
                fprintf(stderr, "[mtg trace] walk: steps by how many steps earlier their record was asked for (0 = never):");
                for (int b = 0; b < 12; b++) fprintf(stderr, " %d:%llu", b, (unsigned long long)io->probe_hist[0][b]);
                fprintf(stderr, "\n");
                return;
            }
            const char* names[5] = {"record load", "used bits of the target", "used bits of the far node", "used bits of the mirror slot",
                                    "(probe alone, L1 hit)"};
            for (int w = 0; w < 5; w++) {
                fprintf(stderr, "[mtg probe] %-28s TSC ticks, 16 per bucket:", names[w]);
                for (int b = 0; b < 40; b++)
                    if (io->probe_hist[w][b]) fprintf(stderr, " %d:%llu", 16 * b, (unsigned long long)io->probe_hist[w][b]);
                fprintf(stderr, "\n");
            }
        }
    } diag_at_exit{&io, diag, lean_loop};
    // Positions whose from-node may still own an unused out-edge, in cycle order from the head.  A position is only
    // recorded if its node had slots left when the walk passed (exhaustion is permanent), which skips about half of
    // the re-root probes.
    HVec<u32> cand;
    cand.bind(scratch.cand, E + 16, 0, false);  // one per step at most, plus one per run
    std::vector<u32> walk_slots;  // output pieces as slots; translated to edge ids at the end (independent gathers)
    walk_slots.reserve(E / 2);
    out.walk_limits.clear();
    double ms_break = 0;
    const bool light = in.matching_light;
    auto is_dummy = [&](u32 q) { return (q & SLOT_DUMMY) != 0; };
    auto weight_of = [&](u32 q) { return (q & SLOT_BREAK) ? in.k : in.dummy_w[in.slot_edge[q & SLOT_MASK] - E0]; };  // q is a dummy
    auto breaks = [&](u32 q) { return (q & SLOT_BREAK) != 0; };
    u64 steps_total = 0;
    const bool timing = trace_slow_calls();
    double ms_runs = 0;
    const double t_loop = now_ms();
    for (u64 e0 = 0; e0 < E0 && steps_total < E / 2; e0++) {
        if (is_used(in.slot_of_edge[e0])) continue;
        // one closed walk per component, started at the lowest unused edge id (every node owns an original edge, so the
        // lowest unused edge of a component is never a dummy)
        size_t cf = 0;
        q_slot.clear();
        q_from.clear();
        cand.clear();
        children.clear();
        u32 start_slot = in.slot_of_edge[e0], start_from = in.start_handle ? in.start_handle[e0] : in.handle[in.from[e0]];
        u32 head_idx = 0;  // position of the element the (rotated) cycle vector currently starts with
        u32 n0 = 0;        // length of the initial closed walk == the root block [0, n0)
        bool rooted = false;
        for (;;) {
            cand.push_back((u32)q_slot.size());  // a walk start is always probed again
            io.q_slot = q_slot.p, io.q_from = q_from.p, io.q_n = q_slot.n;
            io.cand = cand.p, io.cand_n = cand.n;
            const size_t run_b = io.q_n;
            const double tr0 = timing ? now_ms() : 0.0;
            run(io, start_slot, start_from);
            if (timing) ms_runs += now_ms() - tr0;
            if (replay_env) {
                replay_runs.push_back({replay_seq.size(), replay_seq.size() + (io.q_n - run_b), start_from});
                for (size_t i = run_b; i < io.q_n; i++) replay_seq.push_back(q_slot.p[i] & SLOT_MASK);
            }
            q_slot.n = q_from.n = io.q_n;
            cand.n = io.cand_n;
            if (rooted) children.back().end = (u32)q_slot.size();  // the run just appended belongs to the head's block
            else n0 = (u32)q_slot.size();
            // re-root at the first cycle position (from the head) whose from-node still has an unused out-edge
            bool found = false;
            while (cf < cand.size()) {
                const u32 qf = cand[cf];
                bool more_unused;
                const u32 a = first_unused_of(recs, used, q_from[qf], &more_unused);
                if (a != NONE32) {
                    head_idx = qf;  // rotate_left(position)
                    rooted = true;
                    if (children.empty() || children.back().pos != qf)  // first run spliced before qf
                        children.push_back({qf, (u32)q_slot.size(), (u32)q_slot.size()});
                    start_slot = a;
                    start_from = q_from[qf];
                    found = true;
                    break;
                }
                cf++;  // exhausted for good
            }
            if (!found) break;
        }
        double tb = now_ms();
        const size_t len = q_slot.size();
        steps_total += len;
        // in-order expansion of the block tree into slices of q_slot; the root block is the initial closed walk [0, n0)
        const auto child_range = [&](u32 b, u32 e, size_t* ci, size_t* ce) {
            const auto lt = [](const Child& c, u32 v) { return c.pos < v; };
            *ci = std::lower_bound(children.begin(), children.end(), b, lt) - children.begin();
            *ce = std::lower_bound(children.begin(), children.end(), e, lt) - children.begin();
        };
        rope.clear();
        stack.clear();
        {
            Frame f{0u, n0, 0, 0};
            child_range(0u, n0, &f.ci, &f.ce);
            stack.push_back(f);
        }
        size_t total = 0;
        while (!stack.empty()) {
            Frame& f = stack.back();
            if (f.ci == f.ce) {
                if (f.next < f.end) rope.push_back({f.next, f.end}), total += f.end - f.next;
                stack.pop_back();
                continue;
            }
            const Child c = children[f.ci++];
            if (f.next < c.pos) rope.push_back({f.next, c.pos}), total += c.pos - f.next;
            f.next = c.pos;  // the element itself follows its block
            Frame g{c.begin, c.end, 0, 0};
            child_range(c.begin, c.end, &g.ci, &g.ce);
            stack.push_back(g);  // invalidates f
        }
        MTG_REQUIRE(total == len, MTG_ERR_INTERNAL, "cycle expansion lost elements");
        // the reference's cycle vector starts at the head element: rotate the slice list accordingly
        size_t s0 = 0;
        while (s0 < rope.size() && rope[s0].b != head_idx) s0++;
        MTG_REQUIRE(s0 < rope.size(), MTG_ERR_INTERNAL, "head element is not at a slice start");
        order.assign(rope.begin() + s0, rope.end());
        order.insert(order.end(), rope.begin(), rope.begin() + s0);
        // G. greedytigs/mod.rs:736-788: start at the heaviest dummy (the first one on ties, strict `>`), cut at every dummy
        // of weight >= k and at a dummy in position 0.  Nothing is rotated or copied: pieces go straight from the slices
        // to the output.
        const u32* qe = q_slot.data();
        size_t rot_s = 0;
        u32 rot_o = 0, best_w = 0;
        for (size_t si = 0; si < order.size(); si++) {
            const Slice sl = order[si];
            bool done = false;
            for (u32 j = sl.b; j < sl.e; j++) {
                const u32 x = qe[j];
                if (!is_dummy(x)) continue;
                if (light && !breaks(x) && best_w >= in.k) continue;
                const u32 w = weight_of(x);
                if (w > best_w) {
                    best_w = w;
                    rot_s = si;
                    rot_o = j - sl.b;
                    if (light && w == in.k) {  // nothing is heavier than a breaking dummy
                        done = true;
                        break;
                    }
                }
            }
            if (done) break;
        }
        // Every element except the cutting dummies goes to the output, in order.  The emission order is a short list of
        // ranges of the queue; the host cores take equal shares of it: count, prefix, write.  A cut closes a piece only if
        // the piece is not empty, so the limits are the distinct output offsets at which cuts happen.
        struct Range {
            u32 b, e;
            u64 g0;  // position of b in emission order
        };
        std::vector<Range> ranges;
        {
            u64 g = 0;
            auto add = [&](u32 b, u32 e) {
                if (b < e) ranges.push_back({b, e, g}), g += e - b;
            };
            add(order[rot_s].b + rot_o, order[rot_s].e);
            for (size_t si = rot_s + 1; si < order.size(); si++) add(order[si].b, order[si].e);
            for (size_t si = 0; si < rot_s; si++) add(order[si].b, order[si].e);
            add(order[rot_s].b, order[rot_s].b + rot_o);
            MTG_REQUIRE(g == len, MTG_ERR_INTERNAL, "emission order lost elements");
        }
        const size_t base = walk_slots.size();
        walk_slots.resize(base + len);
        u32* const wbase = walk_slots.data() + base;
        const int n_parts = len > (1u << 16) ? host_threads() : 1;
        std::vector<u64> part_out(n_parts + 1, 0), part_cuts(n_parts + 1, 0);
        std::vector<std::vector<u64>> part_cut_at(n_parts);
        // calls f(x, g) for every element of emission positions [g_lo, g_hi)
        auto for_positions = [&](u64 g_lo, u64 g_hi, auto&& f) {
            if (g_lo >= g_hi) return;
            size_t ri = std::upper_bound(ranges.begin(), ranges.end(), g_lo, [](u64 v, const Range& r) { return v < r.g0; }) - ranges.begin() - 1;
            u64 g = g_lo;
            for (; ri < ranges.size() && g < g_hi; ri++) {
                const Range& r = ranges[ri];
                const u32 jb = r.b + (u32)(g - r.g0), je = (u32)std::min<u64>(r.e, r.b + (g_hi - r.g0));
                for (u32 j = jb; j < je; j++, g++) f(qe[j], g);
            }
        };
        const auto cuts = [&](u32 x, u64 g) { return is_dummy(x) && (g == 0 || breaks(x)); };
#pragma omp parallel for schedule(static, 1) num_threads(n_parts) if (n_parts > 1)
        for (int t = 0; t < n_parts; t++) {
            u64 n_out = 0, n_cut = 0;
            for_positions(len * t / n_parts, len * (t + 1) / n_parts, [&](u32 x, u64 g) {
                const bool c = cuts(x, g);
                n_cut += c;
                n_out += !c;
            });
            part_out[t + 1] = n_out;
            part_cuts[t + 1] = n_cut;
        }
        for (int t = 0; t < n_parts; t++) part_out[t + 1] += part_out[t];
#pragma omp parallel for schedule(static, 1) num_threads(n_parts) if (n_parts > 1)
        for (int t = 0; t < n_parts; t++) {
            u32* wp = wbase + part_out[t];
            std::vector<u64>& cut_at = part_cut_at[t];
            cut_at.reserve(part_cuts[t + 1]);
            for_positions(len * t / n_parts, len * (t + 1) / n_parts, [&](u32 x, u64 g) {
                if (cuts(x, g)) cut_at.push_back((u64)(wp - walk_slots.data()));
                else *wp++ = x;
            });
        }
        u64 piece = base;  // start of the piece being collected
        auto close = [&](u64 at) {
            if (at == piece) return;
            MTG_REQUIRE(!is_dummy(walk_slots[piece]), MTG_ERR_INTERNAL, "walk starts with a dummy edge");
            out.walk_limits.push_back(at);
            piece = at;
        };
        for (int t = 0; t < n_parts; t++)
            for (const u64 at : part_cut_at[t]) {
                close(at);
                out.breaking++;
            }
        u64 end = base + part_out[n_parts];
        if (end != piece && is_dummy(walk_slots[end - 1])) end--;  // a trailing (light) dummy is dropped
        close(end);
        walk_slots.resize(end);
        out.cycles++;
        ms_break += now_ms() - tb;
    }
    // E. (greedytigs/mod.rs:708-715) every edge pair was walked exactly once, or the graph was not Eulerian after all
    MTG_REQUIRE(steps_total == E / 2, MTG_ERR_INTERNAL, "Failed to make the graph Eulerian (the closed walks do not cover every edge).");
    double tt = now_ms();
    if (timing && steps_total > 100000)
        fprintf(stderr, "[mtg trace] walk: set-up %.2f ms, run loops %.2f ms (%.2f ns per step), start scan + re-roots %.2f ms, breaking %.2f ms\n",
                t_loop - t3, ms_runs, 1e6 * ms_runs / (double)steps_total, tt - t_loop - ms_runs - ms_break, ms_break);
    // slots -> edge ids: independent gathers, spread over the host cores
    if (out.edges_pinned) out.edges_pinned->resize(walk_slots.size());
    else out.walk_edges.resize(walk_slots.size());
    {
        const u32* ws = walk_slots.data();
        u32* we = out.edges_pinned ? out.edges_pinned->data() : out.walk_edges.data();
        const u32* se = in.slot_edge;
        const i64 n = (i64)walk_slots.size();
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (n > (1 << 16))
        for (i64 i = 0; i < n; i++) we[i] = se[ws[i] & SLOT_MASK];
    }
    double t4 = now_ms();
    if (trace_slow_calls() && getenv("MTG_TRACE_ALL"))
        fprintf(stderr, "[mtg trace] breaking: cycle order + cuts + piece emission %.2f ms, slot -> edge gather %.2f ms\n", ms_break, t4 - tt);
    out.ms_break = ms_break + (t4 - tt);
    out.ms_walk = t4 - t3 - out.ms_break;
    if (replay_env && !replay_seq.empty()) {
        const size_t n = replay_seq.size();
        replay_seq.resize(n + 64, replay_seq[0]);
        std::vector<u32> q2(n + 64), f2(n + 64), c2(n + 64);
        volatile u32 zero_src = 0;
        const u32 zero = zero_src;
        std::string spec = std::string("0:4,") + replay_env;  // the first pass also touches the scratch pages: discarded
        bool first = true;
        for (size_t p = 0; p < spec.size();) {
            size_t q = spec.find(',', p);
            if (q == std::string::npos) q = spec.size();
            int mode = 0, depth = 4;
            sscanf(spec.substr(p, q - p).c_str(), "%d:%d", &mode, &depth);
            depth = std::max(0, std::min(depth, 63));
            p = q + 1;
            memcpy(used, replay_used0.data(), replay_used0.size() * sizeof(u64));
            u64 wrong = 0;
            const double ta = now_ms();
            for (const ReplayRun& r : replay_runs)
                wrong += (pf_kind == 1 ? walk_replay<1> : walk_replay<0>)(recs, used, replay_seq.data(), r.b, r.e, r.from_h, (u32)depth, mode, zero,
                                                                          q2.data(), f2.data(), c2.data());
            const double tb = now_ms();
            if (!first)
                fprintf(stderr, "[mtg replay] mode %d depth %2d: %6.2f ns per step (%zu steps, %zu runs, %llu steps off the sequence)\n", mode, depth,
                        1e6 * (tb - ta) / (double)n, n, replay_runs.size(), (unsigned long long)wrong);
            first = false;
        }
    }
}

}  // namespace

}  // namespace mtg

namespace mtg {

// D on the compacted leftover lists (same rules as `eulerise` above; positions instead of node-indexed arrays).
static void eulerise_sparse(u32 k, TailLeftover& lo, std::vector<u32>& breaking) {
    (void)k;
    auto& ins = lo.in_nodes;
    auto& outs = lo.out_nodes;
    auto& idiff = lo.in_diff;
    auto& odiff = lo.out_diff;
    size_t ip = 0, op = outs.size();
    auto skip_ins = [&] {
        while (ip < ins.size() && idiff[ip] <= 0) ip++;
    };
    auto skip_outs = [&] {
        while (op > 0 && odiff[op - 1] >= 0) op--;
    };
    auto& selfs = lo.self_nodes;
    for (size_t i = 0; i < selfs.size(); i += 2) {  // src/implementation/mod.rs:481-524
        if (i + 1 < selfs.size()) {
            breaking.push_back(selfs[i]);
            breaking.push_back(selfs[i + 1]);
        } else {
            skip_ins();
            MTG_REQUIRE(ip < ins.size(), MTG_ERR_INTERNAL,
                        "Have an uneven number of self-mirrors, but no other nodes with missing in edges.");
            breaking.push_back(selfs[i]);
            breaking.push_back(ins[ip]);
            idiff[ip] -= 1;
            odiff[lo.in_partner[ip]] += 1;
        }
    }
    for (;;) {  // :526-645
        skip_outs();
        if (op == 0) break;
        const size_t o = op - 1;
        skip_ins();
        MTG_REQUIRE(ip < ins.size(), MTG_ERR_INTERNAL, "No further in_nodes left");
        size_t i = ip;
        if (i == lo.out_partner[o] && odiff[o] > -2) {  // choose_in_node_from_iterator :252-285 (in_node == mirror(out_node))
            size_t q = ip + 1;
            while (q < ins.size() && idiff[q] <= 0) q++;
            MTG_REQUIRE(q < ins.size(), MTG_ERR_INTERNAL, "No further in_nodes left");
            i = q;
        }
        breaking.push_back(outs[o]);
        breaking.push_back(ins[i]);
        odiff[o] += 1;
        idiff[i] -= 1;
        const size_t mo = lo.in_partner[i], mi = lo.out_partner[o];  // mirror(in_node) among the outs, mirror(out_node) among the ins
        if (odiff[mo] < 0) odiff[mo] += 1;
        if (idiff[mi] > 0) idiff[mi] -= 1;
    }
    skip_ins();
    MTG_REQUIRE(ip == ins.size(), MTG_ERR_INTERNAL, "eulerise: in-nodes left over");
}

// Variant kept for A/B measurements (MTG_TAIL_HOST=1): everything after the matching on the host.
// Copies of the graph arrays the host-side preparation reads, into page-locked staging.  Issued right behind the graph
// build for graphs that will take the host path, so that they have long arrived when the tail starts.
void stage_tail_inputs(mtg_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    const u64 U = ctx->U, N = ctx->N, E = ctx->E;
    u32* from = ctx->tail_stage[0].as<u32>(E + 1);
    u32* to = ctx->tail_stage[1].as<u32>(E + 1);
    u32* uw = ctx->tail_stage[2].as<u32>(U + 1);
    u32* mirror = ctx->tail_stage[3].as<u32>(N + 1);
    if (E) {
        MTG_CUDA(cudaMemcpyAsync(from, ctx->edge_from.p, E * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(to, ctx->edge_to.p, E * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(uw, ctx->unitig_w.p, U * sizeof(u32), cudaMemcpyDeviceToHost, s));
    }
    if (N) MTG_CUDA(cudaMemcpyAsync(mirror, ctx->mirror.p, N * sizeof(u32), cudaMemcpyDeviceToHost, s));
    ctx->tail_inputs_staged = true;
}

static void finish_walks_host_prep(mtg_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    const u64 U = ctx->U, N = ctx->N, E = ctx->E;
    u32* from = ctx->tail_stage[0].as<u32>(E + 1);  // page-locked staging: full-speed DMA
    u32* to = ctx->tail_stage[1].as<u32>(E + 1);
    u32* uw = ctx->tail_stage[2].as<u32>(U + 1);
    u32* mirror = ctx->tail_stage[3].as<u32>(N + 1);
    if (!ctx->tail_inputs_staged) {  // small graphs: stage_tail_inputs already ran behind the graph build
        stage_tail_inputs(ctx);
        MTG_CUDA(cudaStreamSynchronize(s));
    }
    TailInput in{ctx->k, N, E, from, to, uw, mirror, ctx->h_triples.data(), ctx->n_triples, ctx->opt.p3_oldest_first != 0};
    TailOutput out;
    out.edges_pinned = &ctx->walk_edges;
    run_tail(in, out, ctx->tail_scratch);
    ctx->walk_limits.swap(out.walk_limits);
    ctx->h_dummy_w.swap(out.dummy_w);
    ctx->tail_ms[0] = out.ms_degrees;
    ctx->tail_ms[1] = out.ms_eulerise;
    ctx->tail_ms[2] = out.ms_csr;
    ctx->tail_ms[3] = out.ms_walk;
    ctx->tail_ms[4] = out.ms_break;
    ctx->d_walk_edges.upload(ctx->walk_edges.data(), ctx->walk_edges.size(), s);
    ctx->d_walk_limits.upload(ctx->walk_limits.data(), ctx->walk_limits.size(), s);
    ctx->d_dummy_w.upload(ctx->h_dummy_w.data(), ctx->h_dummy_w.size(), s);  // pageable sources: staged before the call returns
    ctx->n_walk_edges_dev = ctx->walk_edges.size();
    ctx->n_walks_dev = ctx->walk_limits.size();
    ctx->have_walks = true;
}

void finish_walks(mtg_ctx* ctx) {
    MTG_REQUIRE(ctx->have_graph && ctx->have_triples, MTG_ERR_INVALID, "mtg_greedy_match has not run");
    // Small graphs: the device-side preparation is a dozen launches and four round trips, more than the host needs
    // to do the same work (MTG_TAIL_HOST=0/1 forces either path for A/B measurements).
    bool host_prep = ctx->N < TAIL_HOST_PREP_MAX_NODES;
    if (const char* e = getenv("MTG_TAIL_HOST")) host_prep = (*e == '1');
    if (host_prep) return finish_walks_host_prep(ctx);
    ctx->tail_inputs_staged = false;  // this path reuses the staging buffers
    cudaStream_t s = ctx->stream;
    const u64 N = ctx->N, E0 = ctx->E;
    TailScratch& scratch = ctx->tail_scratch;
    double t0 = now_ms();
    // leftover imbalance (device compaction) -> pairing loop (host) -> breaking pairs
    TailLeftover lo;
    tail_leftover(ctx, lo);
    std::vector<u32> breaking;
    eulerise_sparse(ctx->k, lo, breaking);
    const u64 n_break = breaking.size() / 2;
    double t1 = now_ms();
    // walk records on the device (includes the Eulerian check), DMA into page-locked staging
    TailRecords tr;
    tail_build_records(ctx, breaking.data(), n_break, &tr);
    double t1b = now_ms();  // kernels launched (one round trip for the totals in between), DMA queued behind them
    const u64 P = tr.n_pairs;
    (void)N;
    // The walk works on a copy inside its huge-page arena (page-locking the arena itself loses the huge pages); copied in
    // cache-sized chunks by a few threads: one big memcpy would use non-temporal stores and leave everything cold.
    const int copy_threads = host_threads();
    // streaming stores only for record arrays that cannot stay in the caches anyway: a small graph's records are copied
    // with plain stores and are warm when the walk starts (E. coli-size: 0.20 instead of 0.55 ms of walk)
    const bool stream_copy = tr.n_slots * sizeof(WalkRec) > (size_t(64) << 20) &&
                             !(getenv("MTG_TAIL_COPY") && getenv("MTG_TAIL_COPY")[0] == 'm');  // A/B: MTG_TAIL_COPY=memcpy
    auto warm_copy = [copy_threads, stream_copy](void* dst, const void* src, size_t bytes) {
        const size_t chunk = 256 << 10;  // (64 KB .. 8 MB measured alike)
        const i64 n_chunks = (i64)((bytes + chunk - 1) / chunk);
#pragma omp parallel for schedule(static) num_threads(copy_threads) if (bytes > (8u << 20))
        for (i64 c = 0; c < n_chunks; c++) {
            const size_t o = (size_t)c * chunk, len = std::min(chunk, bytes - o);
#if defined(__x86_64__)
            // Streaming stores: the arena lines are written without being read first (a third less memory traffic than a
            // plain copy of half a gigabyte, which cannot stay in the caches anyway).  Records are 64-byte aligned.
            if (stream_copy && len % 64 == 0 && (reinterpret_cast<uintptr_t>((char*)dst + o) & 15u) == 0) {
                const __m128i* a = reinterpret_cast<const __m128i*>((const char*)src + o);
                __m128i* b = reinterpret_cast<__m128i*>((char*)dst + o);
                for (size_t q = 0; q < len / 16; q += 4) {
                    const __m128i x0 = _mm_loadu_si128(a + q), x1 = _mm_loadu_si128(a + q + 1), x2 = _mm_loadu_si128(a + q + 2),
                                  x3 = _mm_loadu_si128(a + q + 3);
                    _mm_stream_si128(b + q, x0);
                    _mm_stream_si128(b + q + 1, x1);
                    _mm_stream_si128(b + q + 2, x2);
                    _mm_stream_si128(b + q + 3, x3);
                }
                _mm_sfence();
                continue;
            }
#endif
            memcpy((char*)dst + o, (const char*)src + o, len);
        }
    };
    // while the records are in flight: dummy weights (matching dummies carry their distance, breaking dummies weigh k)
    TailOutput out;
    out.edges_pinned = &ctx->walk_edges;
    out.dummy_w.resize(2 * P);
    const u32* trp = ctx->h_triples.data();
    u32 max_matching_w = 0;
    for (u64 j = 0; j < ctx->n_triples; j++) {
        out.dummy_w[2 * j] = out.dummy_w[2 * j + 1] = trp[3 * j + 2];
        max_matching_w = std::max(max_matching_w, trp[3 * j + 2]);
    }
    for (u64 j = ctx->n_triples; j < P; j++) out.dummy_w[2 * j] = out.dummy_w[2 * j + 1] = ctx->k;
    const WalkRec* recs = tr.recs;
    double ms_wait = 0, ms_copy = 0;
    if (!getenv("MTG_TAIL_NOCOPY")) {
        WalkRec* arena = static_cast<WalkRec*>(scratch.recs.ensure(std::max<u64>(tr.n_slots, 1) * sizeof(WalkRec)));
        for (int c = 0; c < TAIL_DMA_CHUNKS && tr.n_slots; c++) {  // piece c is copied while piece c+1 is still on the link
            const u64 lo = std::min<u64>((u64)c * tr.chunk_slots, tr.n_slots), hi = std::min<u64>(lo + tr.chunk_slots, tr.n_slots);
            const double ta = now_ms();
            MTG_CUDA(cudaEventSynchronize(ctx->tail_events[c]));
            const double tb = now_ms();
            if (hi > lo) warm_copy(arena + lo, tr.recs + lo, (hi - lo) * sizeof(WalkRec));
            ms_wait += tb - ta, ms_copy += now_ms() - tb;
        }
        recs = arena;
    }
    MTG_CUDA(cudaStreamSynchronize(s));  // the small tables
    const size_t used_bytes = (tr.n_slots / 64 + 2) * sizeof(u64);
    u64* used = static_cast<u64*>(scratch.used.ensure(used_bytes));
    if (tr.used0) memcpy(used, tr.used0, used_bytes);
    else memset(used, 0, used_bytes);  // empty graph
    double t2 = now_ms();
    if (trace_slow_calls() && (t2 - t1b > 30.0 || getenv("MTG_TRACE_ALL")))
        fprintf(stderr, "[mtg trace] tail records, %llu slots: waiting for the DMA %.1f ms, copying %.1f ms, rest %.1f ms\n",
                (unsigned long long)tr.n_slots, ms_wait, ms_copy, t2 - t1b - ms_wait - ms_copy);
    WalkInput w{ctx->k, N, E0, E0 + 2 * P, tr.n_slots, recs, used, nullptr, tr.slot_edge, tr.slot_of_edge, nullptr, tr.handle,
                out.dummy_w.data(), max_matching_w < ctx->k, true};
    walk_and_break(w, out, scratch);
    ctx->walk_limits.swap(out.walk_limits);
    ctx->h_dummy_w.swap(out.dummy_w);
    ctx->tail_ms[0] = t1b - t1;  // device-prepared path: record kernels up to the point where the DMA is queued
    ctx->tail_ms[1] = t1 - t0;
    ctx->tail_ms[2] = t2 - t1b;  // records arriving and being copied into the walk's arena
    ctx->tail_ms[3] = out.ms_walk;
    ctx->tail_ms[4] = out.ms_break;
    // device copies for the output kernels
    ctx->d_walk_edges.upload(ctx->walk_edges.data(), ctx->walk_edges.size(), s);
    ctx->d_walk_limits.upload(ctx->walk_limits.data(), ctx->walk_limits.size(), s);
    ctx->d_dummy_w.upload(ctx->h_dummy_w.data(), ctx->h_dummy_w.size(), s);
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->n_walk_edges_dev = ctx->walk_edges.size();
    ctx->n_walks_dev = ctx->walk_limits.size();
    ctx->have_walks = true;
}

}  // namespace mtg

// Host-only entry: the sequential tail on caller-supplied arrays (no GPU involved).  Lets the host
// logic be tested and timed on its own.
extern "C" int mtg_host_tail(uint32_t k, uint64_t nodes, uint64_t unitigs, const uint32_t* edge_from, const uint32_t* edge_to,
                             const uint32_t* unitig_w, const uint32_t* mirror, const uint32_t* triples, uint64_t n_triples,
                             uint32_t** walk_edges, uint64_t** walk_limits, uint32_t** dummy_w, uint64_t* n_walks,
                             uint64_t* n_walk_edges, uint64_t* n_dummy_edges, double* phase_ms /* [5] or NULL */, char* errbuf,
                             size_t errcap) {
    using namespace mtg;
    auto fail = [&](int code, const std::string& m) {
        if (errbuf && errcap) {
            size_t n = std::min(errcap - 1, m.size());
            memcpy(errbuf, m.data(), n);
            errbuf[n] = 0;
        }
        return code;
    };
    if (!walk_edges || !walk_limits || !dummy_w || !n_walks || !n_walk_edges || !n_dummy_edges) return fail(MTG_ERR_INVALID, "null output");
    try {
        const char* p3 = getenv("MTG_ASSUME_P3_OLDEST_FIRST");
        TailInput in{k, nodes, 2 * unitigs, edge_from, edge_to, unitig_w, mirror, triples, n_triples, p3 && *p3 == '1'};
        TailOutput out;
        static thread_local TailScratch scratch;  // arenas are reused across calls, like the context does
        run_tail(in, out, scratch);
        auto dup = [](const auto& v) {
            using T = typename std::decay<decltype(v)>::type::value_type;
            T* p = (T*)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
            if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
            return p;
        };
        *walk_edges = dup(out.walk_edges);
        *walk_limits = dup(out.walk_limits);
        *dummy_w = dup(out.dummy_w);
        *n_walks = out.walk_limits.size();
        *n_walk_edges = out.walk_edges.size();
        *n_dummy_edges = out.dummy_w.size();
        if (phase_ms) {
            phase_ms[0] = out.ms_degrees;
            phase_ms[1] = out.ms_eulerise;
            phase_ms[2] = out.ms_csr;
            phase_ms[3] = out.ms_walk;
            phase_ms[4] = out.ms_break;
        }
        return MTG_OK;
    } catch (const Error& e) {
        return fail(e.code, e.msg);
    } catch (const std::exception& e) {
        return fail(MTG_ERR_INTERNAL, e.what());
    }
}

extern "C" void mtg_host_free(void* p) { free(p); }
