// mtg_internal.cuh -- context, device buffers and launch helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "matchtigs_b200.h"

namespace mtg {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i8 = int8_t;
using i32 = int32_t;
using i64 = int64_t;

constexpr u32 NONE32 = 0xFFFFFFFFu;
constexpr int TAIL_DMA_CHUNKS = 8;  // pieces in which the walk records travel to the host
constexpr int NUM_SMS_B200 = 148;

// MTG_TRACE=1: calls that take unexpectedly long say on stderr where the time went (allocation, launch, wait).
inline bool trace_slow_calls() {
    static const bool on = [] { const char* e = getenv("MTG_TRACE"); return e && *e && *e != '0'; }();
    return on;
}
inline double wall_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

struct Error {
    int code;
    std::string msg;
};

#define MTG_CUDA(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            throw ::mtg::Error{MTG_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e)};       \
    } while (0)

#define MTG_REQUIRE(cond, code, text)                   \
    do {                                                \
        if (!(cond)) throw ::mtg::Error{(code), (text)}; \
    } while (0)

// Device blocks are handed out in size classes (eight per power of two) and a released block goes to a free list
// of its stream and class instead of back to the driver (prims.cu); the next request of that class on that stream takes it.
// A repeated job therefore allocates nothing after its first pass.  Going back to the driver's stream-ordered pool on
// every release looked equivalent but is not: a 250 MB request that fits no cached block makes the pool rebuild a
// contiguous range out of scattered physical pieces, which was measured at 0.2 - 1.2 s per occurrence inside chr1-size
// jobs (MTG_TRACE=1).  Reuse on the same stream is ordered by the stream itself.  block_cache_trim returns everything.
size_t block_class_bytes(size_t bytes);
void* block_alloc(size_t class_bytes, cudaStream_t s);
void block_free(void* p, size_t class_bytes, cudaStream_t s);
void block_cache_trim(cudaStream_t s);

// Stream-ordered device array that only ever grows (see the block cache above).
// Owning and move-only: a buffer that goes out of scope -- normally or because an MTG_REQUIRE / MTG_CUDA threw --
// returns its memory on the stream it was last sized on.  borrow() wraps memory owned by somebody else.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    size_t n = 0;
    size_t bytes = 0;  // size class of the block
    cudaStream_t st = nullptr;
    bool owned = true;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), cap(o.cap), n(o.n), bytes(o.bytes), st(o.st), owned(o.owned) { o.p = nullptr, o.cap = o.n = o.bytes = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) {
            release(st);
            p = o.p, cap = o.cap, n = o.n, bytes = o.bytes, st = o.st, owned = o.owned;
            o.p = nullptr, o.cap = o.n = o.bytes = 0;
        }
        return *this;
    }
    ~DBuf() { release(st); }
    void resize(size_t count, cudaStream_t s) {
        st = s;
        if (count > cap || !owned) {
            release(s);
            bytes = block_class_bytes((count + 16) * sizeof(T));
            p = static_cast<T*>(block_alloc(bytes, s));
            cap = bytes / sizeof(T);
        }
        n = count;
    }
    void borrow(T* q, size_t count) {
        release(st);
        p = q, n = count, cap = 0, owned = false;
    }
    void zero(cudaStream_t s) {
        if (n) MTG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void fill_ff(cudaStream_t s) {
        if (n) MTG_CUDA(cudaMemsetAsync(p, 0xFF, n * sizeof(T), s));
    }
    void release(cudaStream_t s) {
        if (p && owned) block_free(p, bytes, s);
        p = nullptr;
        cap = n = bytes = 0;
        owned = true;
    }
    void upload(const T* h, size_t count, cudaStream_t s) {
        resize(count, s);
        if (count) MTG_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* h, cudaStream_t s) const {
        if (n) MTG_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// Page-locked host staging memory (grow-only): DMA runs at full PCIe speed only from/to pinned memory.
struct PinnedBuf {
    char* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t want = bytes + bytes / 8 + 4096;
        MTG_CUDA(cudaHostAlloc((void**)&p, want, cudaHostAllocDefault));
        cap = want;
    }
    template <class T>
    T* as(size_t count) {
        ensure(count * sizeof(T));
        return reinterpret_cast<T*>(p);
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Vector-like view over page-locked memory (contents are NOT preserved when it grows): results that come back from the
// device land here at full link speed instead of being staged through the driver's bounce buffer.
template <class T>
struct PinnedVec {
    PinnedBuf buf;
    size_t n = 0;
    T* data() { return reinterpret_cast<T*>(buf.p); }
    const T* data() const { return reinterpret_cast<const T*>(buf.p); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return data()[i]; }
    const T& operator[](size_t i) const { return data()[i]; }
    void resize(size_t count) {
        buf.ensure(std::max<size_t>(count, 1) * sizeof(T));
        n = count;
    }
    void clear() { n = 0; }
    void release() {
        buf.release();
        n = 0;
    }
};

// ---- records of the host Euler walk (built on the device by tail_prep.cu, or on the host for small graphs) ----
// Every out-edge of every node owns one slot; the slots of a node are consecutive, in petgraph's iteration order
// (newest edge first).  A node is addressed by its HANDLE: first slot (even) | H_FOUR (four slots instead
// of two) | H_BIG (more than four out-edges: slot `base` is a header whose `to` holds the degree, entries start at
// base + 2).  Which slots are used up lives in a bitset over slots (padding slots are marked from the start), so "first
// unused out-edge of node h" is one unaligned 64-bit load and a count-trailing-zeros -- for ANY node whose handle is
// known, without touching its records.  That is what the records exploit: the walk is one dependent cache miss per
// step, and a record carries the handles of the nodes 2 .. WALK_DEPTH steps ahead along the first two slots of every
// node on the way, so the walk can work out exactly which record it will need WALK_DEPTH steps from now and prefetch
// that one line (instead of fanning out over 2^depth candidates).
#ifndef MTG_WALK_DEPTH
#define MTG_WALK_DEPTH 4
#endif
static_assert(MTG_WALK_DEPTH >= 3 && MTG_WALK_DEPTH <= 5, "walk records carry 3, 4 or 5 levels");
constexpr u32 WALK_DEPTH = MTG_WALK_DEPTH;
constexpr u32 WALK_HINTS = (1u << WALK_DEPTH) - 2;  // 2 + 4 + ... + 2^(depth-1) handles
constexpr u32 H_BIG = 0x80000000u, H_FOUR = 1u, H_BASE = 0x7FFFFFFEu;
constexpr u32 SLOT_MASK = 0x3FFFFFFFu, SLOT_DUMMY = 0x40000000u, SLOT_BREAK = 0x80000000u;
struct alignas(MTG_WALK_DEPTH >= 4 ? 64 : 32) WalkRec {
    u32 to;     // handle of the node this edge leads to
    u32 mslot;  // slot of the mirror edge | SLOT_DUMMY | SLOT_BREAK (dummy of weight >= k)
    // Handles of the nodes ahead: level L = where the walk can be L steps after taking this edge, indexed by the slot
    // choices j1 (at `to`), j2, ... j(L-1) on the way.  Behind `to` only the first two slots of every node are followed.
    // Two layouts, told apart by the node `to` (which the walk knows anyway):
    //   `to` owns two slots (H_FOUR clear):  level L at h[2^(L-1) - 2 ...], index = binary number j1 j2 .. j(L-1); L = 2 .. depth
    //   `to` owns four slots (H_FOUR set):   level L at h[2^L - 4 ...],     index = j1 * 2^(L-2) + binary number j2 .. j(L-1),
    //                                        j1 = 0 .. 3; L = 2 .. depth - 1 (one level less, but a third or fourth visit of
    //                                        `to` is still announced)
    // h[0], h[1] (where `to`'s slots 0 and 1 lead) mean the same in both layouts.
    u32 h[WALK_HINTS];
};
static_assert(sizeof(WalkRec) == 4u << MTG_WALK_DEPTH, "record size: 32, 64 or 128 bytes");
__host__ __device__ inline u32 walk_level_two(u32 level) { return (1u << (level - 1)) - 2u; }   // first index of level L, two-slot layout
__host__ __device__ inline u32 walk_level_four(u32 level) { return (1u << level) - 4u; }        // ... four-slot layout
__host__ __device__ inline bool walk_is_four(u32 handle) { return (handle & (H_FOUR | H_BIG)) == H_FOUR; }
__host__ __device__ inline u32 walk_cap(u32 d) { return d <= 2 ? 2u : d <= 4 ? 4u : 2u + ((d + 1) & ~1u); }
__host__ __device__ inline u32 walk_handle(u32 base, u32 d) { return base | ((d > 2 && d <= 4) ? H_FOUR : 0u) | (d > 4 ? H_BIG : 0u); }
// first entry slot and entry count of node h given its degree (entries of a big node sit behind its header pair)
__host__ __device__ inline u32 walk_first_slot(u32 h) { return (h & H_BASE) + ((h & H_BIG) ? 2u : 0u); }
// Level `level` of the hints of record r from level - 1 of the records of `to`'s slots (levels are built one after the
// other over all records).  `to_deg` = out-degree of the node r leads to.
__host__ __device__ inline void walk_fill_hints(WalkRec* recs, u32 s, u32 to_deg, u32 level) {
    WalkRec& r = recs[s];
    const u32 c0 = walk_first_slot(r.to);
    const bool four = walk_is_four(r.to);
    if (four && level >= WALK_DEPTH) return;  // the four-slot layout ends one level earlier
    const u32 fan = four ? 4u : 2u;           // slots of `to` this record follows
    const u32 width = 1u << (level - 2);      // entries per slot of `to` at this level
    const u32 dst0 = four ? walk_level_four(level) : walk_level_two(level);
    for (u32 j = 0; j < fan; j++) {
        const bool have = j < to_deg;
        const WalkRec& c = recs[c0 + (have ? j : 0u)];
        if (level == 2) {
            r.h[dst0 + j] = have ? c.to : r.to;
        } else {  // the child's level - 1, restricted to its first two slots, sits where ITS layout puts it
            const u32 src0 = walk_is_four(c.to) ? walk_level_four(level - 1) : walk_level_two(level - 1);
            for (u32 x = 0; x < width; x++) r.h[dst0 + j * width + x] = c.h[src0 + x];
        }
    }
}

// Host arena backed by an anonymous mapping with MADV_HUGEPAGE (host_tail.cpp); grow-only, cached across calls.
struct HugeBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    void* ensure(size_t bytes);
    void* ensure_pinned(size_t bytes);  // additionally page-locked (cudaHostRegister) so the GPU can DMA into it
    void release();
    HugeBuf() = default;
    HugeBuf(const HugeBuf&) = delete;
    HugeBuf& operator=(const HugeBuf&) = delete;
    ~HugeBuf() { release(); }
};
struct TailScratch {
    HugeBuf out_deg, in_deg, diff, recs, used, queue, cand, cyc, slot_edge, slot_of_edge, handle;
};
// Nodes still unbalanced after the matching, ascending node id (device compaction, tail_prep.cu).
// partner = position of the node's mirror in the opposite list.
struct TailLeftover {
    std::vector<u32> self_nodes, out_nodes, out_partner, in_nodes, in_partner;
    std::vector<i32> out_diff, in_diff;
};

// Device-side counters of the search / matching kernels.
struct DevStats {
    unsigned long long sources_searched;
    unsigned long long settled;
    unsigned long long relaxed;
    unsigned long long candidates;
    unsigned long long truncated;
    unsigned long long overflow;
    unsigned long long labels;      // labelled nodes summed over the searches (tier 0): the "distance array size" of a search
    unsigned long long max_labels;  // largest number of labelled nodes of one search (tier 0)
    unsigned long long max_open;    // largest number of open (labelled, unsettled) nodes of one search: the "heap size"
};

}  // namespace mtg

// The parity assumptions of SURVEY.md Appendix C as switches (defaults = the assumptions as stated there).  A golden
// output of the real reference that disagrees with an assumption is then a flag flip here and in the oracle
// (mto_set_option, same names), not a rewrite.  Set with mtg_ctx_set_option or MTG_ASSUME_<NAME>=1 in the environment.
struct mtg_options {
    int p1_tie_desc = 0;          // P1: ties inside a distance level settle the LARGER node id first
    int p1_exclusive_bound = 0;   // P1: the search bound k-1 is exclusive
    int p2_self_mirror_zero = 0;  // P2: a self-mirror node never counts as unbalanced in the imbalance scan
    int p3_oldest_first = 0;      // P3: out-edges are iterated oldest first
    int p6_bcalm_kmer_numbering = 0;  // P6: --bcalm-in numbers nodes like --fa-in (k-mer join, links ignored)
    int p7_first_root_wins = 0;   // P7: union of equal ranks attaches the second root below the first
};

struct mtg_ctx {
    int device = 0;
    mtg_options opt;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t tail_events[mtg::TAIL_DMA_CHUNKS] = {};  // DMA progress of the walk records
    cudaEvent_t ev_build[3] = {nullptr, nullptr, nullptr};  // start of the last graph build, end of its parsing, end of the build
    bool build_timed = false, in_text_build = false;
    float last_kernel_ms = 0;  // tier-0 search kernel of the last run_searches call
    mtg::DevStats h_dstats{};   // host copy of the search counters taken by the last run_searches call ...
    bool h_dstats_final = false;  // ... and whether nothing ran after that copy
    std::string err;
    uint64_t launches = 0;
    int num_sms = mtg::NUM_SMS_B200;

    // ---- resident graph (step 1) ----
    uint32_t k = 0;
    uint64_t U = 0, N = 0, E = 0, Es = 0, S = 0, T = 0, self_mirror_unbalanced = 0;
    // Work space of the device-side text parser (parse.cu), kept across calls (~1 byte per text byte): per 32-byte chunk
    // its line-state key, its record / sequence-byte / link counts and their exclusive sums; plus the staged text.
    struct ParseScratch {
        mtg::DBuf<mtg::u32> key, n_rec, n_seq, n_link, rbase, sbase, lbase, totals;
        mtg::DBuf<char> text;
    } parse_ws;
    bool tail_inputs_staged = false;  // host copies of edge_from/edge_to/unitig_w/mirror are in tail_stage[0..3]
    uint64_t target_mult_total = 0;  // sum of the targets' multiplicities = upper bound for the number of matches
    bool have_graph = false, have_seqs = false;
    bool graph_needs_k = false;  // links graph built without k (matchtigs_build_graph): finish_deferred_graph completes it
    mtg::DBuf<mtg::u64> seq_words;    // 2-bit store: base i at bits [2(i%32), 2(i%32)+1] of word i/32; A0 C1 T2 G3
    mtg::DBuf<mtg::u64> seq_off;      // [U+1] base offsets
    uint64_t total_bases = 0;
    mtg::DBuf<mtg::u32> unitig_w;     // [U] k-mers per unitig
    mtg::DBuf<mtg::u32> edge_from, edge_to;  // [2U]
    mtg::DBuf<mtg::u32> mirror;       // [N]
    mtg::DBuf<mtg::u32> out_deg;      // [N]
    mtg::DBuf<mtg::i32> imbalance;    // [N] initial node_multiplicities
    mtg::DBuf<mtg::u32> target_bits;  // [(N+31)/32] initial in_node_map
    mtg::DBuf<mtg::u32> sources;      // [S] ascending
    mtg::DBuf<mtg::u32> row_s;        // [N+1] short-edge CSR
    mtg::DBuf<mtg::u32> col_s;        // [Es]
    mtg::DBuf<mtg::u8> w_s;           // [Es]

    // ---- candidates (step 2) ----
    uint32_t cap = 0, shard_rank = 0, shard_count = 1;
    uint64_t S_local = 0;
    mtg::DBuf<mtg::u64> cand;         // [S_local * cap]  node | dist << 32
    mtg::DBuf<mtg::u32> cand_meta;    // [S_local]        count | truncated << 31
    mtg::DBuf<mtg::DevStats> dstats;  // [1]
    bool have_cand = false;

    // ---- matching (step 3) ----
    mtg::DBuf<mtg::i32> final_mult;   // [N] node multiplicities after the matching == leftover imbalance
    mtg::DBuf<mtg::u32> triples;      // [3 * n_triples] device
    uint64_t n_triples = 0;
    mtg::PinnedVec<uint32_t> h_triples;  // page-locked: 12 B per matched pair come back after every matching
    bool have_triples = false;

    // ---- host tail ----
    std::vector<uint32_t> h_dummy_w;  // weight of dummy edge e at [e - 2U]
    double tail_ms[5] = {0, 0, 0, 0, 0};  // degrees, eulerise, csr, walk, break
    mtg::PinnedVec<uint32_t> walk_edges;  // page-locked: written by the tail, uploaded for the output kernels
    std::vector<uint64_t> walk_limits;
    bool have_walks = false;
    mtg::DBuf<mtg::u32> d_walk_edges;
    mtg::DBuf<mtg::u64> d_walk_limits;
    mtg::DBuf<mtg::u32> d_dummy_w;    // weights of dummy edges, index = edge id - 2U
    uint64_t n_walk_edges_dev = 0, n_walks_dev = 0;  // lengths of d_walk_edges / d_walk_limits (ranks that received the walks hold no host copy)

    // ---- multi-GPU (comm.cpp) ----
    void* comm = nullptr;             // ncclComm_t
    uint32_t comm_rank = 0, comm_world = 1;
    mtg::DBuf<mtg::u64> gathered_rec;   // candidate slices of all ranks: [rank][local source][cap]
    mtg::DBuf<mtg::u32> gathered_meta;
    bool text_event_recorded = false;   // the start of the last text build was already recorded by the caller (sliced upload)

    mtg::PinnedBuf text_stage[3];     // bitvector / GFA / FASTA bytes, valid until the next call of the same kind
    mtg::PinnedBuf tail_stage[6];     // DMA targets of the host tail (graph arrays for the host-prepared path; walk records,
                                      // slot tables and the initial bitset for the device-prepared path)
    mtg::TailScratch tail_scratch;    // cached working arrays of the host tail

    mtg_search_stats stats{};
    // scratch
    mtg::DBuf<mtg::u8> scratch_a, scratch_b;
};

namespace mtg {

inline dim3 grid_for(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

#define MTG_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
        (ctx)->launches++;                                                      \
        MTG_CUDA(cudaGetLastError());                                           \
    } while (0)

// ---- primitives (prims.cu) ----
// out[i] = sum_{j<i} in[j]; returns nothing; `total` (device pointer, may be null) receives the grand total.
void exclusive_sum_u32(mtg_ctx* ctx, const u32* in, u32* out, size_t n, u32* d_total);
void exclusive_sum_u8(mtg_ctx* ctx, const u8* in, u32* out, size_t n, u32* d_total);
void exclusive_sum_u32_to_u64(mtg_ctx* ctx, const u32* in, u64* out, size_t n, u64* d_total);
// out[i] = max_{j<=i} in[j]
void inclusive_max_u32(mtg_ctx* ctx, const u32* in, u32* out, size_t n);
// Stable LSD radix sort of (key words, value) by bits [0, key_bits) of the 64*nwords-bit key (word 0 = least significant).
// Buffers ping-pong; returns 0 if the result is in the *_a buffers, 1 if in *_b.
int radix_sort_pairs(mtg_ctx* ctx, u64* k0_a, u64* k0_b, u64* k1_a, u64* k1_b, u32* v_a, u32* v_b, size_t n,
                     int nwords, int key_bits);
int radix_sort_pairs_u32(mtg_ctx* ctx, u32* k_a, u32* k_b, u32* v_a, u32* v_b, size_t n, int key_bits);

// ---- graph construction (graph.cu) ----
// total_bases: offsets[U] if the caller already knows it (saves a round trip when the offsets live on the device)
constexpr u64 UNKNOWN_TOTAL = ~0ull;
// prepacked: ctx->seq_words / seq_off / total_bases already hold the sequences (device-side parser); seq and offsets are unused
void build_graph_from_sequences(mtg_ctx* ctx, const char* seq, const u64* offsets, u64 U, u32 k, bool on_device,
                                u64 total_bases = UNKNOWN_TOTAL, bool prepacked = false);
void build_graph_from_links(mtg_ctx* ctx, u64 U, const u64* weights, u64 n_links, const u64* a, const u8* sa, const u64* b,
                            const u8* sb, u32 k, const char* seq, const u64* offsets, bool on_device,
                            u64 total_bases = UNKNOWN_TOTAL, bool prepacked = false);
void finish_deferred_graph(mtg_ctx* ctx, u32 k);
// device-side FASTA / bcalm2 record parser feeding the two builders (parse.cu)
void build_graph_from_text(mtg_ctx* ctx, const char* text, u64 len, bool bcalm, u32 k, bool text_on_device);

// Graphs below this node count prepare the sequential tail on the host (fewer round trips than launches).
constexpr u64 TAIL_HOST_PREP_MAX_NODES = 1u << 12;
void stage_tail_inputs(mtg_ctx* ctx);  // host_tail.cpp

// ---- search + matching (dijkstra.cu, match.cu) ----
void dijkstra_candidates(mtg_ctx* ctx, u32 cap, u32 shard_rank, u32 shard_count);
void greedy_match(mtg_ctx* ctx, const u64* d_records_all, const u32* d_meta_all, u32 shard_count);

// ---- outputs (emit.cu) ----
// walk_lo / walk_hi: only the share of the text that belongs to these walks; where[0] = its offset in the whole text, where[1] = whole length
u64 dup_bitvector(mtg_ctx* ctx, char* out, u64 cap, bool size_only, const char** view, u64 walk_lo = 0, u64 walk_hi = ~0ull,
                  u64* where = nullptr);
u64 assemble_tigs(mtg_ctx* ctx, int format, char* out, u64 cap, bool size_only, const char** view, u64 walk_lo = 0, u64 walk_hi = ~0ull,
                  u64* where = nullptr);

// ---- host tail (host_tail.cpp) and its device-side preparation (tail_prep.cu) ----
void finish_walks(mtg_ctx* ctx);
void tail_leftover(mtg_ctx* ctx, TailLeftover& lo);
// Builds the walk records on the device and copies them (plus slot -> edge id, original edge -> slot, node -> handle and the
// initial used-slot bitset) into the page-locked staging buffers of the context.
struct TailRecords {
    u64 n_slots = 0, n_pairs = 0;
    u64 chunk_slots = 0;  // the records arrive in TAIL_DMA_CHUNKS pieces of this many slots; piece c is complete once
                          // ctx->tail_events[c] has fired (the other arrays: once the stream is idle)
    WalkRec* recs = nullptr;
    u32 *slot_edge = nullptr, *slot_of_edge = nullptr, *handle = nullptr;
    u64* used0 = nullptr;
};
void tail_build_records(mtg_ctx* ctx, const u32* breaking_pairs, u64 n_break, TailRecords* out);

}  // namespace mtg
