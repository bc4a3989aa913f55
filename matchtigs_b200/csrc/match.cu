// match.cu -- step 3: the source-ordered greedy matching (K5), exact and parallel.
//
// Reference semantics (greedytigs/mod.rs:301-523 at --threads 1, restated in SURVEY.md 3.3):
// sources are visited in ascending node id; each takes the nearest still-open in-nodes of its
// Dijkstra list and updates up to four node multiplicities.  In the single-threaded reference
// `in_node_map[v]` is set exactly while `mult[v] > 0`, so the whole state is the multiplicity
// array, targets only ever close, and a Dijkstra call with target_amount m+1 against the
// current map equals "the first m+1 entries of the precomputed list L(src) with mult > 0".
// With m+1 candidates at most one (the source's own mirror) is skipped, so the reference's
// `while` loop runs exactly once per source.
//
// Parallelisation = dataflow over per-target wait lists.  For every in-type node x we build the
// sorted list rev[x] of the sources that might ever take x (x is mirror(src) or an open entry of
// L(src)).  holder(x) is the lowest-indexed source of rev[x] that is not done yet.  A source
// commits as soon as it is the holder of everything its commit reads and writes in the current
// state -- mirror(src) and the first m+1 open entries -- i.e. as soon as every lower-indexed
// source that could still touch those nodes is done.  Only in-type nodes need guarding: every
// out-type node a commit touches (src, mirror of an accepted entry) is shared only with sources
// that have its mirror in their own set.  Unheld entries may be read racily: they can only flip
// open -> closed, through a lower-indexed source, which is what the sequential order shows.
//
// There are no rounds and no grid barriers: one persistent kernel hands sources out in ascending
// index order (atomic counter), each thread spins until its source is the holder of what it needs,
// commits, and publishes `done` with release semantics.  A waiting source only ever waits for
// lower-indexed sources, all of which were handed out earlier to resident threads, and the lowest
// unfinished source never waits -- so the kernel cannot deadlock and its critical path is the
// longest dependency chain, not (#rounds x barrier latency).
//
// Capped lists: a source whose list is truncated and runs dry is "insufficient".  Everything
// from the smallest insufficient index j* on is discarded, the multiplicities are rebuilt from
// the triples of sources < j*, truncated lists of sources >= j* are searched again (4x cap,
// against the current open map) and matching resumes at j*.
#include <algorithm>
#include <memory>

#include "mtg_internal.cuh"

namespace mtg {

void run_searches(mtg_ctx* ctx, const u32* bitmap, const u32* work_list, u64 n_work, u32 shard_rank, u32 shard_count, u32 cap,
                  u64* records, u32* meta);

namespace {

constexpr int TB = 256;
#ifndef MTG_MATCH_SLEEP_NS
#define MTG_MATCH_SLEEP_NS 100  // pause between two attempts of a blocked source (A/B: -DMTG_MATCH_SLEEP_NS=0|20|50)
#endif
constexpr u32 META_TRUNC = 0x80000000u;
constexpr u32 META_COUNT = 0x00FFFFFFu;
constexpr u32 NO_INDEX = 0xFFFFFFFFu;
constexpr unsigned long long MATCH_SPIN_LIMIT = 1ull << 26;  // blocked attempts of one source (each >= 100 ns) before giving up: minutes

struct MatchArgs {
    const u32* sources;
    const u32* mirror;
    i32* mult;
    const u64* list_addr;   // device address of the first record of every source's list
    const u32* list_meta;   // count | truncated << 31
    const u32* pend;        // pending sources of this phase, ascending
    u32 n_pend;
    const u32* rev_ptr;     // [N+1] wait lists per node
    const u32* rev;         // [P] source indices, ascending inside every list
    u32* cur;               // [N] first possibly-unfinished position of every wait list
    u32* done;              // [S] 1 once a source has committed (or was dropped)
    const u32* trip_off;    // [S] first triple slot of every source
    u32* trip_cnt;          // [S]
    u32* trip_slots;        // [3 * total]
    u32* min_insufficient;  // [1]
    unsigned long long* retries;  // [1] blocked attempts (diagnostic)
    u32* error;             // [1] invariant violations
    unsigned long long* counter;  // [1] hand-out position in `pend`
};

__device__ __forceinline__ u32 ld_acquire(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(u32* p, u32 v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// mult[] is rewritten by other SMs while the kernel runs: always go through L2.
__device__ __forceinline__ i32 ldm(const i32* mult, u32 x) { return __ldcg(&mult[x]); }
__device__ __forceinline__ void addm(i32* mult, u32 x, i32 d) { __stcg(&mult[x], __ldcg(&mult[x]) + d); }

// Lowest-indexed unfinished source waiting on x.  The acquire loads of `done` order the caller's later reads of
// mult[] after the commits of the sources it skips.
__device__ __forceinline__ u32 holder(const MatchArgs& a, u32 x) {
    const u32 end = a.rev_ptr[x + 1];
    u32 c = __ldcg(&a.cur[x]);
    const u32 c0 = c;
    while (c < end && ld_acquire(&a.done[a.rev[c]])) c++;
    if (c != c0) atomicMax(&a.cur[x], c);
    return c < end ? a.rev[c] : NO_INDEX;
}

enum TryResult { TRY_BLOCKED = 0, TRY_DONE = 1, TRY_INSUFFICIENT = 2 };

// Applies the reference's matching rules (greedytigs/mod.rs:301-523) to source i if every lower-indexed source
// that could touch what it reads or writes is done.
__device__ TryResult try_source(const MatchArgs& a, u32 i) {
    const u32 out_node = a.sources[i];
    const u32 M = a.mirror[out_node];
    if (ldm(a.mult, M) == 0) {  // :318-320; 0 is final (in-type multiplicities never grow), no need to wait for anyone
        a.trip_cnt[i] = 0;
        return TRY_DONE;
    }
    if (holder(a, M) != i) return TRY_BLOCKED;
    const bool out_self = M == out_node;
    i32 m = ldm(a.mult, M);  // :306-311, stable now
    if (m == 0) {
        a.trip_cnt[i] = 0;
        return TRY_DONE;
    }
    const u64* list = reinterpret_cast<const u64*>(a.list_addr[i]);
    const u32 meta = a.list_meta[i];
    const u32 count = meta & META_COUNT;
    const bool truncated = (meta & META_TRUNC) != 0;
    // The Dijkstra call (:324-335) == the first m+1 entries of the list that are open now.
    const u32 target_amount = (u32)m + 1;
    u32 found = 0, last_pos = 0;
    for (u32 p = 0; p < count && found < target_amount; p++) {
        const u32 x = (u32)list[p];
        if (x != M) {
            if (ldm(a.mult, x) <= 0) continue;            // closed entries never reopen
            if (holder(a, x) != i) return TRY_BLOCKED;    // a lower-indexed source may still take x
            if (ldm(a.mult, x) <= 0) continue;            // re-read after the acquire: closed by a source that just finished
        }
        found++;
        last_pos = p;
    }
    if (found == 0) {  // distances.is_empty() :338-346
        a.trip_cnt[i] = 0;
        return truncated ? TRY_INSUFFICIENT : TRY_DONE;
    }
    const bool abort_after_this = found < target_amount;  // :348
    u32* slots = a.trip_slots + 3ull * a.trip_off[i];
    u32 emitted = 0;
    for (u32 p = 0; p <= last_pos; p++) {  // :350
        const u64 rec = list[p];
        const u32 in_node = (u32)rec;
        // entries of `distances` = open at call time; M was open (m > 0), every other node is touched only by its own iteration
        if (in_node != M && ldm(a.mult, in_node) <= 0) continue;
        bool sme = false;
        if (in_node == M) {  // :352-358
            if (m < 2) continue;
            sme = true;
        }
        m = out_self ? ldm(a.mult, out_node) : -ldm(a.mult, out_node);  // :401-410
        if (m == 0) break;                                               // :412-414
        const u32 in_mirror = a.mirror[in_node];
        const i32 r = sme ? 2 : 1;
        slots[3 * emitted + 0] = out_node;  // :461
        slots[3 * emitted + 1] = in_node;
        slots[3 * emitted + 2] = (u32)(rec >> 32);
        emitted++;
        if (out_self) {  // :463-473
            addm(a.mult, out_node, -1);
        } else {
            addm(a.mult, out_node, r);
            addm(a.mult, M, -r);
        }
        m = -ldm(a.mult, out_node);  // :474
        if (!sme) {                  // :476-491
            addm(a.mult, in_node, -1);
            if (in_mirror != in_node) addm(a.mult, in_mirror, 1);
        }
    }
    a.trip_cnt[i] = emitted;
    if (m > 0) {
        // With m+1 candidates at most one (mirror(src)) is skipped, so m reaches 0 unless the list ran out (:504-511).
        if (!abort_after_this) atomicExch(a.error, 1u);  // would be a second Dijkstra call: impossible at --threads 1
        if (truncated) return TRY_INSUFFICIENT;           // the real call would have returned more entries
    }
    return TRY_DONE;
}

__global__ void __launch_bounds__(TB) match_dataflow_kernel(MatchArgs a) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long retries = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.counter, 32ull);  // ascending hand-out: whoever we wait for is already running
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= a.n_pend) break;
        const unsigned long long idx = base + lane;
        if (idx >= a.n_pend) continue;
        const u32 i = a.pend[idx];
        // The release store that publishes this source's commit sits INSIDE the loop, on the path that leaves it: lanes of
        // one warp wait for each other here, so the store must be issued before any point where the compiler could make
        // the finished lanes wait for the spinning ones.  The spin is bounded: a dependency that never resolves (an
        // invariant violation, not a legal state) raises the error flag instead of hanging the device.
        for (unsigned long long spins = 0;; spins++) {
            TryResult r = TRY_DONE;
            if (i <= ((volatile u32*)a.min_insufficient)[0]) {  // everything above the smallest insufficient source is redone
                r = try_source(a, i);
                if (r == TRY_INSUFFICIENT) atomicMin(a.min_insufficient, i);
                if (r == TRY_BLOCKED && spins > MATCH_SPIN_LIMIT) {
                    atomicExch(a.error, 2u);
                    r = TRY_DONE;
                }
            }
            if (r != TRY_BLOCKED) {
                st_release(&a.done[i], 1u);  // publishes the commit (release) to the sources waiting behind this one
                break;
            }
            retries++;
            if (MTG_MATCH_SLEEP_NS) __nanosleep(MTG_MATCH_SLEEP_NS);
        }
    }
    for (int o = 16; o > 0; o >>= 1) retries += __shfl_down_sync(0xffffffffu, retries, o);
    if (lane == 0 && retries) atomicAdd(a.retries, retries);
}

// ---------------- per-phase set-up ----------------
__global__ void __launch_bounds__(TB)
    init_lists(const u32* __restrict__ sources, const u32* __restrict__ mirror, const i32* __restrict__ imbalance,
               const u64* __restrict__ records_all, const u32* __restrict__ meta_all, u64 S, u32 shard_count, u64 padded, u32 cap,
               u64* __restrict__ list_addr, u32* __restrict__ list_meta, u32* __restrict__ max_trip) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= S) return;
    u64 slot = (i % shard_count) * padded + i / shard_count;
    u32 meta = meta_all[slot];
    list_addr[i] = reinterpret_cast<u64>(records_all + slot * cap);
    list_meta[i] = meta;
    i32 m0 = imbalance[mirror[sources[i]]];
    max_trip[i] = ((meta & META_COUNT) && m0 > 0) ? (u32)m0 : 0u;
}

__global__ void __launch_bounds__(TB) flag_pending(const u32* __restrict__ list_meta, u64 S, u64 lo, u32* __restrict__ flag) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < S) flag[i] = (i >= lo && (list_meta[i] & META_COUNT)) ? 1u : 0u;
}
__global__ void __launch_bounds__(TB) flag_requery(const u32* __restrict__ list_meta, u64 S, u64 lo, u32* __restrict__ flag) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < S) flag[i] = (i >= lo && (list_meta[i] & META_TRUNC)) ? 1u : 0u;
}
__global__ void __launch_bounds__(TB) compact_indices(const u32* __restrict__ flag, const u32* __restrict__ pos, u64 n, u32* __restrict__ out) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = (u32)i;
}
// Wait lists by counting: a pending source waits on mirror(src) and on every entry of its list that is open at the start
// of the phase.  Pass 1 counts per source (the 64-bit guard of the total) and per node (the list lengths); after the scan
// of the lengths pass 2 drops every source into its nodes' lists through per-node cursors -- in arbitrary order -- and pass
// 3 sorts every list in place (lists hold a handful of sources; the few long ones get a CTA each).  This replaces a
// three-pass radix sort of all (node, source) pairs.
__global__ void __launch_bounds__(TB)
    count_waits(const u32* __restrict__ pend, const u32* __restrict__ n_pend_dev, u64 S, const u64* __restrict__ list_addr,
                const u32* __restrict__ list_meta, const i32* __restrict__ mult, const u32* __restrict__ sources,
                const u32* __restrict__ mirror, u32* __restrict__ cnt, u32* __restrict__ deg) {
    u32 r = blockIdx.x * TB + threadIdx.x;
    if (r >= S) return;
    if (r >= *n_pend_dev) {  // launched over all S slots: the host learns n_pend together with the wait-list total
        cnt[r] = 0;
        return;
    }
    const u32 i = pend[r];
    const u64* list = reinterpret_cast<const u64*>(list_addr[i]);
    const u32 count = list_meta[i] & META_COUNT;
    u32 c = 1;
    atomicAdd(&deg[mirror[sources[i]]], 1u);
    for (u32 p = 0; p < count; p++) {
        const u32 x = (u32)list[p];
        if (mult[x] > 0) {
            atomicAdd(&deg[x], 1u);
            c++;
        }
    }
    cnt[r] = c;
}
__global__ void __launch_bounds__(TB)
    fill_waits(const u32* __restrict__ pend, u32 n_pend, const u64* __restrict__ list_addr, const u32* __restrict__ list_meta,
               const i32* __restrict__ mult, const u32* __restrict__ sources, const u32* __restrict__ mirror, u32* __restrict__ cursor,
               u32* __restrict__ rev) {
    u32 r = blockIdx.x * TB + threadIdx.x;
    if (r >= n_pend) return;
    const u32 i = pend[r];
    const u64* list = reinterpret_cast<const u64*>(list_addr[i]);
    const u32 count = list_meta[i] & META_COUNT;
    rev[atomicAdd(&cursor[mirror[sources[i]]], 1u)] = i;
    for (u32 p = 0; p < count; p++) {
        const u32 x = (u32)list[p];
        if (mult[x] > 0) rev[atomicAdd(&cursor[x], 1u)] = i;
    }
}
constexpr u32 WAIT_LONG = 48;  // lists longer than this are sorted by a CTA
// ascending source index inside every list (equal entries -- mirror(src) that is also an entry of its own list -- stay
// next to each other, which is all the order the matching needs)
__global__ void __launch_bounds__(TB)
    sort_wait_lists(const u32* __restrict__ rev_ptr, u64 N, u32* __restrict__ rev, u32* __restrict__ long_nodes, u32* __restrict__ n_long) {
    const u64 x = (u64)blockIdx.x * TB + threadIdx.x;
    if (x >= N) return;
    const u32 b = rev_ptr[x], e = rev_ptr[x + 1];
    const u32 len = e - b;
    if (len < 2) return;
    if (len > WAIT_LONG) {
        long_nodes[atomicAdd(n_long, 1u)] = (u32)x;
        return;
    }
    for (u32 p = b + 1; p < e; p++) {  // insertion sort, in place
        const u32 v = rev[p];
        u32 q = p;
        while (q > b && rev[q - 1] > v) {
            rev[q] = rev[q - 1];
            q--;
        }
        rev[q] = v;
    }
}
// one CTA per long list: rank by counting (ties by position), through a scratch copy
__global__ void __launch_bounds__(TB)
    sort_long_wait_lists(const u32* __restrict__ rev_ptr, const u32* __restrict__ long_nodes, const u32* __restrict__ n_long,
                         u32* __restrict__ rev, u32* __restrict__ scratch) {
    for (u32 w = blockIdx.x; w < *n_long; w += gridDim.x) {
        const u32 x = long_nodes[w];
        const u32 b = rev_ptr[x], e = rev_ptr[x + 1];
        for (u32 p = b + threadIdx.x; p < e; p += TB) scratch[p] = rev[p];
        __syncthreads();
        for (u32 p = b + threadIdx.x; p < e; p += TB) {
            const u32 v = scratch[p];
            u32 rank = 0;
            for (u32 q = b; q < e; q++) {
                const u32 o = scratch[q];
                rank += (o < v) || (o == v && q < p);
            }
            rev[b + rank] = v;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(TB) final_counts(const u32* __restrict__ trip_cnt, u64 S, u64 lo, u64 hi, u32* __restrict__ out) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < S) out[i] = (i >= lo && i < hi) ? trip_cnt[i] : 0u;
}
__global__ void __launch_bounds__(TB)
    copy_final(const u32* __restrict__ cnt, const u32* __restrict__ dst_off, const u32* __restrict__ trip_off,
               const u32* __restrict__ trip_slots, u64 S, u64 base, u32* __restrict__ triples) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= S) return;
    u32 c = cnt[i];
    const u32* src = trip_slots + 3ull * trip_off[i];
    u32* dst = triples + 3ull * (base + dst_off[i]);
    for (u32 j = 0; j < 3 * c; j++) dst[j] = src[j];
}
// mult = imbalance + effect of all final triples (the updates commute)
__global__ void __launch_bounds__(TB) apply_triples(const u32* __restrict__ triples, u64 n, const u32* __restrict__ mirror, i32* mult) {
    u64 t = (u64)blockIdx.x * TB + threadIdx.x;
    if (t >= n) return;
    u32 out_node = triples[3 * t], in_node = triples[3 * t + 1];
    u32 M = mirror[out_node];
    bool out_self = M == out_node;
    bool sme = !out_self && in_node == M;
    i32 r = sme ? 2 : 1;
    if (out_self) atomicAdd(&mult[out_node], -1);
    else {
        atomicAdd(&mult[out_node], r);
        atomicAdd(&mult[M], -r);
    }
    if (!sme) {
        atomicAdd(&mult[in_node], -1);
        u32 im = mirror[in_node];
        if (im != in_node) atomicAdd(&mult[im], 1);
    }
}
__global__ void __launch_bounds__(TB) open_bits_from_mult(const i32* __restrict__ mult, u64 N, u32* __restrict__ bits) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    bool open = v < N && mult[v] > 0;
    unsigned b = __ballot_sync(0xffffffffu, open);
    if ((threadIdx.x & 31) == 0 && (v & ~31ull) < N) bits[v >> 5] = b;
}
__global__ void __launch_bounds__(TB)
    adopt_requery(const u32* __restrict__ work_list, u64 n, const u64* __restrict__ pool, const u32* __restrict__ pool_meta, u32 cap,
                  u64* __restrict__ list_addr, u32* __restrict__ list_meta) {
    u64 t = (u64)blockIdx.x * TB + threadIdx.x;
    if (t >= n) return;
    u32 i = work_list[t];
    list_addr[i] = reinterpret_cast<u64>(pool + t * cap);
    list_meta[i] = pool_meta[t];
}

__global__ void __launch_bounds__(1024) sum_u32_to_u64(const u32* __restrict__ in, u64 n, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long part[32];
    unsigned long long acc = 0;
    for (u64 i = threadIdx.x; i < n; i += 1024) acc += in[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 32; w++) t += part[w];
        *out = t;
    }
}

}  // namespace

void greedy_match(mtg_ctx* ctx, const u64* d_records_all, const u32* d_meta_all, u32 shard_count) {
    MTG_REQUIRE(ctx->have_graph, MTG_ERR_INVALID, "no graph resident");
    cudaStream_t s = ctx->stream;
    const u64 S = ctx->S, N = ctx->N;
    u32 cap = ctx->cap;
    if (!d_records_all) {
        MTG_REQUIRE(ctx->have_cand && ctx->shard_count == 1, MTG_ERR_INVALID,
                    "no gathered candidates given and the local result does not cover all sources");
        d_records_all = ctx->cand.p;
        d_meta_all = ctx->cand_meta.p;
        shard_count = 1;
    }
    MTG_REQUIRE(cap >= 1, MTG_ERR_INVALID, "mtg_dijkstra_candidates has not run");
    ctx->n_triples = 0;
    ctx->h_triples.clear();
    ctx->stats.match_rounds = 0;
    ctx->stats.requery_phases = 0;
    ctx->stats.preextended_sources = 0;
    ctx->final_mult.resize(N, s);
    if (N) MTG_CUDA(cudaMemcpyAsync(ctx->final_mult.p, ctx->imbalance.p, N * sizeof(i32), cudaMemcpyDeviceToDevice, s));
    if (S == 0) {
        MTG_CUDA(cudaStreamSynchronize(s));
        ctx->have_triples = true;
        ctx->stats.matched = 0;
        return;
    }
    const u64 padded = (S + shard_count - 1) / shard_count;
    MTG_CUDA(cudaEventRecord(ctx->ev0, s));
    DBuf<i32>& mult = ctx->final_mult;  // ends up as the leftover imbalance the tail starts from
    DBuf<u64> list_addr;
    DBuf<u32> list_meta, max_trip, trip_off, trip_cnt, trip_slots, flag, pos, pend, small, work_list, open_bits, pool_meta;
    DBuf<u32> wait_cnt, rev, rev_scratch, long_nodes, rev_ptr, cur, done;
    DBuf<unsigned long long> big;  // [0] hand-out counter, [1] retries
    std::vector<DBuf<u64>> pools;
    list_addr.resize(S, s);
    list_meta.resize(S, s);
    max_trip.resize(S, s);
    trip_off.resize(S, s);
    trip_cnt.resize(S, s);
    flag.resize(S, s);
    pos.resize(S, s);
    pend.resize(S, s);
    done.resize(S, s);
    wait_cnt.resize(S, s);
    rev_ptr.resize(N + 1, s);
    cur.resize(N + 1, s);
    small.resize(16, s);  // [0] pending count, [3] min_insufficient, [5..7] scan totals, [8] error
    small.zero(s);
    big.resize(3, s);
    MTG_LAUNCH(ctx, init_lists, grid_for(S, TB), TB, 0, ctx->sources.p, ctx->mirror.p, ctx->imbalance.p, d_records_all, d_meta_all, S,
               shard_count, padded, cap, list_addr.p, list_meta.p, max_trip.p);
    exclusive_sum_u32(ctx, max_trip.p, trip_off.p, S, small.p + 5);
    // every match closes one unit of some target's multiplicity: the graph build's total bounds the triple slots
    const u64 slot_bound = std::max<u64>(ctx->target_mult_total, 1);
    MTG_REQUIRE(slot_bound < 0xFFFFFFFFull / 3, MTG_ERR_UNSUPPORTED, "too many potential matches");
    trip_slots.resize(3 * slot_bound, s);
    ctx->triples.resize(3 * slot_bound, s);

    int occ = 0;
    MTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, match_dataflow_kernel, TB, 0));
    if (occ < 1) occ = 1;

    // Truncated lists get a deeper second search up front (8x cap, against the initial target map, so the result is a
    // longer prefix of the same list).  They are a few percent of the sources, so this costs a fraction of the main
    // search -- and it makes "a truncated list ran dry", which throws away everything from that source on and costs a
    // whole extra search + matching phase, a rare event instead of the rule on repeat-rich graphs.
    if (cap < 4096 && !getenv("MTG_NO_PREEXTEND")) {
        MTG_LAUNCH(ctx, flag_requery, grid_for(S, TB), TB, 0, list_meta.p, S, (u64)0, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 7);
        work_list.resize(S, s);
        MTG_LAUNCH(ctx, compact_indices, grid_for(S, TB), TB, 0, flag.p, pos.p, S, work_list.p);
        u32 n_trunc = 0;
        MTG_CUDA(cudaMemcpyAsync(&n_trunc, small.p + 7, sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        if (n_trunc && (u64)n_trunc * 4 <= S) {
            const u32 cap2 = std::min<u32>(cap * 8, 4096);
            pools.emplace_back();
            pools.back().resize((u64)n_trunc * cap2, s);
            pool_meta.resize(n_trunc, s);
            run_searches(ctx, ctx->target_bits.p, work_list.p, n_trunc, 0, 1, cap2, pools.back().p, pool_meta.p);
            MTG_LAUNCH(ctx, adopt_requery, grid_for(n_trunc, TB), TB, 0, work_list.p, (u64)n_trunc, pools.back().p, pool_meta.p, cap2, list_addr.p,
                       list_meta.p);
            cap = cap2;
            ctx->stats.preextended_sources = n_trunc;
        }
    }

    u64 lo = 0, n_final = 0, retries_total = 0;
    bool copied_out = false;  // triples + counters already fetched with the last phase's round trip
    for (int phase = 0;; phase++) {
        MTG_REQUIRE(phase < 16, MTG_ERR_INTERNAL, "matching did not converge");
        // pending = sources >= lo with a non-empty list, ascending
        MTG_LAUNCH(ctx, flag_pending, grid_for(S, TB), TB, 0, list_meta.p, S, lo, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 0);
        MTG_LAUNCH(ctx, compact_indices, grid_for(S, TB), TB, 0, flag.p, pos.p, S, pend.p);
        // wait lists: lengths per node counted here, filled and sorted below (ascending sources per node)
        MTG_CUDA(cudaMemsetAsync(cur.p, 0, (N + 1) * sizeof(u32), s));  // used as the length histogram first
        MTG_LAUNCH(ctx, count_waits, grid_for(S, TB), TB, 0, pend.p, small.p + 0, S, list_addr.p, list_meta.p, mult.p, ctx->sources.p,
                   ctx->mirror.p, wait_cnt.p, cur.p);
        exclusive_sum_u32(ctx, cur.p, rev_ptr.p, N + 1, small.p + 6);  // total = rev_ptr[N]
        // the 32-bit offsets are safe if even the worst case fits; otherwise the same total is taken in 64 bits
        const bool need_guard = (u64)S * ((u64)cap + 1) >= 0xFFFFFFF0ull;
        if (need_guard) MTG_LAUNCH(ctx, sum_u32_to_u64, 1, 1024, 0, wait_cnt.p, S, big.p + 2);
        else MTG_CUDA(cudaMemsetAsync(big.p + 2, 0, sizeof(unsigned long long), s));
        u32 h_np[7];
        unsigned long long h_total64 = 0;
        MTG_CUDA(cudaMemcpyAsync(h_np, small.p, sizeof(h_np), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaMemcpyAsync(&h_total64, big.p + 2, sizeof(h_total64), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        MTG_REQUIRE(h_total64 < 0xFFFFFFF0ull, MTG_ERR_UNSUPPORTED, "wait lists exceed 2^32 entries (candidate lists too deep for this many sources)");
        const u32 n_pend = h_np[0], P = h_np[6];
        trip_cnt.zero(s);
        u32 min_insuff = NO_INDEX;
        if (n_pend) {
            rev.resize(P, s);
            MTG_CUDA(cudaMemcpyAsync(cur.p, rev_ptr.p, (N + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, s));  // fill cursors
            MTG_LAUNCH(ctx, fill_waits, grid_for(n_pend, TB), TB, 0, pend.p, n_pend, list_addr.p, list_meta.p, mult.p, ctx->sources.p,
                       ctx->mirror.p, cur.p, rev.p);
            long_nodes.resize(P / WAIT_LONG + 2, s);  // at most P / WAIT_LONG lists are longer than WAIT_LONG
            MTG_CUDA(cudaMemsetAsync(small.p + 9, 0, sizeof(u32), s));
            MTG_LAUNCH(ctx, sort_wait_lists, grid_for(N, TB), TB, 0, rev_ptr.p, N, rev.p, long_nodes.p, small.p + 9);
            rev_scratch.resize(P, s);
            MTG_LAUNCH(ctx, sort_long_wait_lists, (u32)std::min<u64>(P / WAIT_LONG + 1, (u64)ctx->num_sms * 4), TB, 0, rev_ptr.p, long_nodes.p,
                       small.p + 9, rev.p, rev_scratch.p);
            MTG_CUDA(cudaMemcpyAsync(cur.p, rev_ptr.p, (N + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, s));  // first unfinished position per list
            done.zero(s);
            u32 init_small[2] = {NO_INDEX, 0};
            MTG_CUDA(cudaMemcpyAsync(small.p + 3, init_small, sizeof(u32), cudaMemcpyHostToDevice, s));
            big.zero(s);
            MatchArgs a{};
            a.sources = ctx->sources.p;
            a.mirror = ctx->mirror.p;
            a.mult = mult.p;
            a.list_addr = list_addr.p;
            a.list_meta = list_meta.p;
            a.pend = pend.p;
            a.n_pend = n_pend;
            a.rev_ptr = rev_ptr.p;
            a.rev = rev.p;
            a.cur = cur.p;
            a.done = done.p;
            a.trip_off = trip_off.p;
            a.trip_cnt = trip_cnt.p;
            a.trip_slots = trip_slots.p;
            a.min_insufficient = small.p + 3;
            a.retries = big.p + 1;
            a.error = small.p + 8;
            a.counter = big.p;
            const u32 grid = (u32)std::min<u64>(((u64)n_pend + TB - 1) / TB, (u64)ctx->num_sms * occ);
            MTG_CUDA(cudaEventRecord(ctx->ev2, s));
            MTG_LAUNCH(ctx, match_dataflow_kernel, grid, TB, 0, a);
            MTG_CUDA(cudaEventRecord(ctx->ev3, s));
            u32 h_small[9];
            unsigned long long h_big[2];
            MTG_CUDA(cudaMemcpyAsync(h_small, small.p, sizeof(h_small), cudaMemcpyDeviceToHost, s));
            MTG_CUDA(cudaMemcpyAsync(h_big, big.p, sizeof(h_big), cudaMemcpyDeviceToHost, s));
            MTG_CUDA(cudaStreamSynchronize(s));
            MTG_REQUIRE(h_small[8] != 2, MTG_ERR_INTERNAL, "matching did not make progress (a source waited for a dependency that never resolved)");
            MTG_REQUIRE(h_small[8] == 0, MTG_ERR_INTERNAL, "matching invariant violated (second Dijkstra call for one source)");
            if (phase == 0) MTG_CUDA(cudaEventElapsedTime(&ctx->stats.match_kernel_ms, ctx->ev2, ctx->ev3));
            min_insuff = h_small[3];
            retries_total += h_big[1];
        }
        const u64 jstar = std::min<u64>(min_insuff, S);
        // finalise sources [lo, jstar)
        MTG_LAUNCH(ctx, final_counts, grid_for(S, TB), TB, 0, trip_cnt.p, S, lo, jstar, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 5);
        MTG_LAUNCH(ctx, copy_final, grid_for(S, TB), TB, 0, flag.p, pos.p, trip_off.p, trip_slots.p, S, n_final, ctx->triples.p);
        u32 added = 0;
        MTG_CUDA(cudaMemcpyAsync(&added, small.p + 5, sizeof(u32), cudaMemcpyDeviceToHost, s));
        if (jstar >= S && slot_bound <= (1u << 20)) {
            // last phase of a small job: the triples (up to their bound), the counters and the count share one round trip
            ctx->h_triples.resize(3 * slot_bound);
            MTG_CUDA(cudaMemcpyAsync(ctx->h_triples.data(), ctx->triples.p, 3 * slot_bound * sizeof(u32), cudaMemcpyDeviceToHost, s));
            MTG_CUDA(cudaMemcpyAsync(&ctx->h_dstats, ctx->dstats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
            MTG_CUDA(cudaEventRecord(ctx->ev1, s));
            copied_out = true;
        }
        MTG_CUDA(cudaStreamSynchronize(s));
        n_final += added;
        if (jstar >= S) break;
        // ---- requery phase ----
        ctx->stats.requery_phases++;
        MTG_REQUIRE(cap < 4096, MTG_ERR_UNSUPPORTED, "a candidate list deeper than 4096 entries was needed");
        cap = std::min<u32>(cap * 4, 4096);
        MTG_CUDA(cudaMemcpyAsync(mult.p, ctx->imbalance.p, N * sizeof(i32), cudaMemcpyDeviceToDevice, s));
        if (n_final) MTG_LAUNCH(ctx, apply_triples, grid_for(n_final, TB), TB, 0, ctx->triples.p, n_final, ctx->mirror.p, mult.p);
        open_bits.resize((N + 31) / 32 + 1, s);
        MTG_LAUNCH(ctx, open_bits_from_mult, grid_for((N + 31) / 32 * 32, TB), TB, 0, mult.p, N, open_bits.p);
        MTG_LAUNCH(ctx, flag_requery, grid_for(S, TB), TB, 0, list_meta.p, S, jstar, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 7);
        work_list.resize(S, s);
        MTG_LAUNCH(ctx, compact_indices, grid_for(S, TB), TB, 0, flag.p, pos.p, S, work_list.p);
        u32 n_req = 0;
        MTG_CUDA(cudaMemcpyAsync(&n_req, small.p + 7, sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        MTG_REQUIRE(n_req > 0, MTG_ERR_INTERNAL, "insufficient source without a truncated list");
        pools.emplace_back();
        pools.back().resize((u64)n_req * cap, s);
        pool_meta.resize(n_req, s);
        run_searches(ctx, open_bits.p, work_list.p, n_req, 0, 1, cap, pools.back().p, pool_meta.p);
        MTG_LAUNCH(ctx, adopt_requery, grid_for(n_req, TB), TB, 0, work_list.p, (u64)n_req, pools.back().p, pool_meta.p, cap, list_addr.p,
                   list_meta.p);
        lo = jstar;
    }
    ctx->n_triples = n_final;
    ctx->stats.matched = n_final;
    ctx->stats.match_rounds = retries_total;
    ctx->h_triples.resize(3 * n_final);
    if (!copied_out) {
        if (n_final) MTG_CUDA(cudaMemcpyAsync(ctx->h_triples.data(), ctx->triples.p, 3 * n_final * sizeof(u32), cudaMemcpyDeviceToHost, s));
        // searches of requery phases added to the device counters
        MTG_CUDA(cudaMemcpyAsync(&ctx->h_dstats, ctx->dstats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaEventRecord(ctx->ev1, s));
        MTG_CUDA(cudaStreamSynchronize(s));
    }
    float ms = 0;
    MTG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.match_ms = ms;
    const DevStats& h = ctx->h_dstats;
    ctx->stats.settled_nodes = h.settled;
    ctx->stats.relaxed_edges = h.relaxed;
    ctx->stats.overflow_sources = h.overflow;
    ctx->stats.labelled_nodes = h.labels;
    ctx->stats.max_labelled_nodes = h.max_labels;
    ctx->stats.max_open_nodes = h.max_open;
    for (auto& p : pools) p.release(s);
    list_addr.release(s);
    big.release(s);
    for (DBuf<u32>* b : {&list_meta, &max_trip, &trip_off, &trip_cnt, &trip_slots, &flag, &pos, &pend, &small, &work_list, &open_bits,
                         &pool_meta, &wait_cnt, &rev, &rev_scratch, &long_nodes, &rev_ptr, &cur, &done})
        b->release(s);
    ctx->have_triples = true;
    ctx->have_walks = false;
}

}  // namespace mtg
