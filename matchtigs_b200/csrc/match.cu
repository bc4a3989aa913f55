// match.cu -- step 3: the source-ordered greedy matching (K5), exact and parallel.
//
// Reference semantics (greedytigs/mod.rs:301-523 at --threads 1, restated in SURVEY.md 3.3):
// sources are visited in ascending node id; each takes the nearest still-open in-nodes of its
// Dijkstra list and updates up to four node multiplicities.  In the single-threaded reference
// `in_node_map[v]` is set exactly while `mult[v] > 0`, so the whole state is the multiplicity
// array, and a Dijkstra call with target_amount t against the current map equals "the first t
// entries of the precomputed list L(src) whose multiplicity is still positive".
//
// Parallelisation = deterministic reservations.  A source's commit reads/writes only its
// touch set {src, mirror(src)} + {x, mirror(x) : x in L(src)}.  Each round every pending source
// writes its index into res[x] with atomicMin for all x in its touch set, and a source commits
// iff it holds every reservation, i.e. no lower-indexed pending source shares a node with it.
// Commits of one round touch disjoint state, and the lowest pending source always commits, so
// the result equals the sequential pass.  One cooperative launch runs all rounds (grid.sync()).
//
// Capped lists: a source that runs out of known open candidates while its list is truncated is
// "insufficient".  Everything from the smallest insufficient index j* on is discarded, the
// multiplicities are rebuilt from the triples of sources < j*, truncated lists of sources >= j*
// are searched again (4x cap, against the current open map) and matching resumes at j*.
#include <cooperative_groups.h>

#include <algorithm>
#include <memory>

#include "mtg_internal.cuh"

namespace cg = cooperative_groups;

namespace mtg {

void run_searches(mtg_ctx* ctx, const u32* bitmap, const u32* work_list, u64 n_work, u32 shard_rank, u32 shard_count, u32 cap,
                  u64* records, u32* meta);

namespace {

constexpr int TB = 256;
constexpr u32 META_TRUNC = 0x80000000u;
constexpr u32 META_COUNT = 0x00FFFFFFu;
constexpr u32 NO_INDEX = 0xFFFFFFFFu;
constexpr u32 MATCH_HIST = 48;

struct MatchArgs {
    const u32* sources;
    const u32* mirror;
    i32* mult;
    const u64* list_addr;  // device address of the first record of every source's list
    const u32* list_meta;  // count | truncated << 31
    unsigned long long* res;  // [N] reservation words: (~round) << 32 | source index
    u32* pend[3];
    u32* counts;           // [3]
    const u32* trip_off;   // [S] first triple slot of every source
    u32* trip_cnt;         // [S]
    u32* trip_slots;       // [3 * total]
    u32* min_insufficient; // [1]
    u32* rounds;           // [1]
    u32* hist;             // [MATCH_HIST] pending sources at the start of each of the first rounds (diagnostic)
    u32* error;            // [1] invariant violations
    u32 round_base;
};

__device__ __forceinline__ void reserve(unsigned long long* res, u32 x, unsigned long long word) { atomicMin(&res[x], word); }
// mult[] and res[] are rewritten by other SMs between rounds: read them through L2 (ld.global.cg).
__device__ __forceinline__ i32 ldm(const i32* mult, u32 x) { return __ldcg(&mult[x]); }
__device__ __forceinline__ void addm(i32* mult, u32 x, i32 d) { __stcg(&mult[x], __ldcg(&mult[x]) + d); }
__device__ __forceinline__ unsigned long long ldr(const unsigned long long* res, u32 x) { return __ldcg(&res[x]); }

// Reservations are asymmetric.  A pending source RESERVES mirror(src) and every still-open entry of its
// whole list (anything it might ever take: closed entries never reopen, so they need no protection), but to
// COMMIT it only has to HOLD what it reads and writes in the current state: mirror(src) and the first m+1
// open entries.  Only in-type nodes are reserved: every access a commit makes to an out-type node v
// (v = src, or v = mirror(x) of an accepted entry x) is shared only with sources that have mirror(v) in
// their own set.  Holding x means no lower-indexed pending source lists x, so its value is final for this
// source; a higher-indexed source can never hold x while this one is pending.  Unheld entries may be read
// racily: they can only flip open -> closed, by a lower-indexed source, which is what the sequential order
// would have shown; an unheld entry read as open blocks the commit.
__device__ __forceinline__ void reserve_source(const MatchArgs& a, u32 i, unsigned long long word) {
    reserve(a.res, a.mirror[a.sources[i]], word);
    const u64* list = reinterpret_cast<const u64*>(a.list_addr[i]);
    const u32 count = a.list_meta[i] & META_COUNT;
    for (u32 p = 0; p < count; p++) {
        const u32 x = (u32)list[p];
        if (ldm(a.mult, x) > 0) reserve(a.res, x, word);
    }
}

enum TryResult { TRY_BLOCKED = 0, TRY_DONE = 1, TRY_INSUFFICIENT = 2 };

// Applies the reference's matching rules (greedytigs/mod.rs:301-523) to source i if it holds its reservations.
__device__ TryResult try_source(const MatchArgs& a, u32 i, unsigned long long word) {
    const u32 out_node = a.sources[i];
    const u32 M = a.mirror[out_node];
    if (ldr(a.res, M) != word) return TRY_BLOCKED;
    const bool out_self = M == out_node;
    i32 m = ldm(a.mult, M);  // :306-311
    if (m == 0) {            // :318-320
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
            if (ldm(a.mult, x) <= 0) continue;
            if (ldr(a.res, x) != word) return TRY_BLOCKED;
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

constexpr u32 LOCAL_MODE_MAX = 2048;  // at most this many pending sources: finish inside one CTA (no grid barriers)

__global__ void __launch_bounds__(TB) match_rounds_kernel(MatchArgs a) {
    cg::grid_group grid = cg::this_grid();
    u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 nthreads = (u64)gridDim.x * blockDim.x;
    bool local = false;
    u32 round = 0;
    for (;; round++) {
        const u32 cur = round % 3, nxt = (round + 1) % 3, spare = (round + 2) % 3;
        const u32 n = ((volatile u32*)a.counts)[cur];
        if (n == 0) break;
        if (!local && n <= LOCAL_MODE_MAX) {  // uniform decision: every CTA reads the same n after the barrier
            if (blockIdx.x != 0) return;
            local = true;
            tid = threadIdx.x;
            nthreads = blockDim.x;
        }
        if (tid == 0 && round < MATCH_HIST) a.hist[round] = n;
        const unsigned long long tag = (unsigned long long)(0xFFFFFFFFu - (a.round_base + round)) << 32;
        const u32* pend = a.pend[cur];
        for (u64 idx = tid; idx < n; idx += nthreads) {
            const u32 i = __ldcg(&pend[idx]);
            if (i > ((volatile u32*)a.min_insufficient)[0]) continue;
            reserve_source(a, i, tag | i);
        }
        if (tid == 0) a.counts[spare] = 0;
        if (local) {
            __threadfence();
            __syncthreads();
        } else {
            grid.sync();
        }
        for (u64 idx = tid; idx < n; idx += nthreads) {
            const u32 i = __ldcg(&pend[idx]);
            if (i > ((volatile u32*)a.min_insufficient)[0]) continue;  // will be discarded anyway
            const TryResult r = try_source(a, i, tag | i);
            if (r == TRY_BLOCKED) a.pend[nxt][atomicAdd(&a.counts[nxt], 1u)] = i;
            else if (r == TRY_INSUFFICIENT) atomicMin(a.min_insufficient, i);
        }
        if (local) {
            __threadfence();
            __syncthreads();
        } else {
            grid.sync();
        }
    }
    if (tid == 0) *a.rounds = round;
}

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
    if (S == 0) {
        ctx->have_triples = true;
        ctx->stats.matched = 0;
        return;
    }
    const u64 padded = (S + shard_count - 1) / shard_count;
    MTG_CUDA(cudaEventRecord(ctx->ev0, s));
    DBuf<i32> mult;
    DBuf<unsigned long long> res;
    DBuf<u64> list_addr;
    DBuf<u32> list_meta, max_trip, trip_off, trip_cnt, trip_slots, flag, pos, pend0, pend1, pend2, small, work_list, open_bits, pool_meta;
    std::vector<DBuf<u64>> pools;
    mult.resize(N, s);
    MTG_CUDA(cudaMemcpyAsync(mult.p, ctx->imbalance.p, N * sizeof(i32), cudaMemcpyDeviceToDevice, s));
    res.resize(N, s);
    res.fill_ff(s);
    list_addr.resize(S, s);
    list_meta.resize(S, s);
    max_trip.resize(S, s);
    trip_off.resize(S, s);
    trip_cnt.resize(S, s);
    flag.resize(S, s);
    pos.resize(S, s);
    pend0.resize(S, s);
    pend1.resize(S, s);
    pend2.resize(S, s);
    small.resize(16 + MATCH_HIST, s);  // [0..2] counts, [3] min_insufficient, [4] rounds, [5] scan total, [6] scan total 2, [16..] hist
    small.zero(s);
    MTG_LAUNCH(ctx, init_lists, grid_for(S, TB), TB, 0, ctx->sources.p, ctx->mirror.p, ctx->imbalance.p, d_records_all, d_meta_all, S,
               shard_count, padded, cap, list_addr.p, list_meta.p, max_trip.p);
    exclusive_sum_u32(ctx, max_trip.p, trip_off.p, S, small.p + 5);
    u32 total_slots = 0;
    MTG_CUDA(cudaMemcpyAsync(&total_slots, small.p + 5, sizeof(u32), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    trip_slots.resize(3ull * std::max<u32>(total_slots, 1), s);
    ctx->triples.resize(3ull * std::max<u32>(total_slots, 1), s);

    int dev_blocks = 0;
    MTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dev_blocks, match_rounds_kernel, TB, 0));
    MTG_REQUIRE(dev_blocks >= 1, MTG_ERR_CUDA, "matching kernel does not fit on an SM");
    const u32 coop_grid = (u32)ctx->num_sms * (u32)std::min(dev_blocks, 8);

    u64 lo = 0, n_final = 0;
    u32 round_base = 0;
    for (int phase = 0;; phase++) {
        MTG_REQUIRE(phase < 16, MTG_ERR_INTERNAL, "matching did not converge");
        // pending = sources >= lo with a non-empty list
        MTG_LAUNCH(ctx, flag_pending, grid_for(S, TB), TB, 0, list_meta.p, S, lo, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 0);
        MTG_LAUNCH(ctx, compact_indices, grid_for(S, TB), TB, 0, flag.p, pos.p, S, pend0.p);
        u32 init_small[5] = {0, 0, 0, NO_INDEX, 0};
        MTG_CUDA(cudaMemcpyAsync(small.p + 1, init_small + 1, 4 * sizeof(u32), cudaMemcpyHostToDevice, s));
        trip_cnt.zero(s);
        MatchArgs a{};
        a.sources = ctx->sources.p;
        a.mirror = ctx->mirror.p;
        a.mult = mult.p;
        a.list_addr = list_addr.p;
        a.list_meta = list_meta.p;
        a.res = res.p;
        a.pend[0] = pend0.p;
        a.pend[1] = pend1.p;
        a.pend[2] = pend2.p;
        a.counts = small.p;
        a.trip_off = trip_off.p;
        a.trip_cnt = trip_cnt.p;
        a.trip_slots = trip_slots.p;
        a.min_insufficient = small.p + 3;
        a.rounds = small.p + 4;
        a.hist = small.p + 16;
        a.error = small.p + 7;
        a.round_base = round_base;
        void* kargs[] = {&a};
        MTG_CUDA(cudaLaunchCooperativeKernel((void*)match_rounds_kernel, dim3(coop_grid), dim3(TB), kargs, 0, s));
        ctx->launches++;
        u32 h_small[8];
        MTG_CUDA(cudaMemcpyAsync(h_small, small.p, sizeof(h_small), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        MTG_REQUIRE(h_small[7] == 0, MTG_ERR_INTERNAL, "matching invariant violated (second Dijkstra call for one source)");
        const u32 rounds = h_small[4];
        if (phase == 0) MTG_CUDA(cudaMemcpyAsync(ctx->match_hist, small.p + 16, sizeof(ctx->match_hist), cudaMemcpyDeviceToHost, s));
        round_base += rounds + 1;
        ctx->stats.match_rounds += rounds;
        const u64 jstar = std::min<u64>(h_small[3], S);
        // finalise sources [lo, jstar)
        MTG_LAUNCH(ctx, final_counts, grid_for(S, TB), TB, 0, trip_cnt.p, S, lo, jstar, flag.p);
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 5);
        MTG_LAUNCH(ctx, copy_final, grid_for(S, TB), TB, 0, flag.p, pos.p, trip_off.p, trip_slots.p, S, n_final, ctx->triples.p);
        u32 added = 0;
        MTG_CUDA(cudaMemcpyAsync(&added, small.p + 5, sizeof(u32), cudaMemcpyDeviceToHost, s));
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
        exclusive_sum_u32(ctx, flag.p, pos.p, S, small.p + 6);
        work_list.resize(S, s);
        MTG_LAUNCH(ctx, compact_indices, grid_for(S, TB), TB, 0, flag.p, pos.p, S, work_list.p);
        u32 n_req = 0;
        MTG_CUDA(cudaMemcpyAsync(&n_req, small.p + 6, sizeof(u32), cudaMemcpyDeviceToHost, s));
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
    ctx->h_triples.resize(3 * n_final);
    if (n_final) MTG_CUDA(cudaMemcpyAsync(ctx->h_triples.data(), ctx->triples.p, 3 * n_final * sizeof(u32), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaEventRecord(ctx->ev1, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    MTG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.match_ms = ms;
    // searches of requery phases added to the device counters
    DevStats h{};
    MTG_CUDA(cudaMemcpyAsync(&h, ctx->dstats.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->stats.settled_nodes = h.settled;
    ctx->stats.relaxed_edges = h.relaxed;
    ctx->stats.overflow_sources = h.overflow;
    for (auto& p : pools) p.release(s);
    mult.release(s);
    res.release(s);
    list_addr.release(s);
    for (DBuf<u32>* b : {&list_meta, &max_trip, &trip_off, &trip_cnt, &trip_slots, &flag, &pos, &pend0, &pend1, &pend2, &small, &work_list,
                         &open_bits, &pool_meta})
        b->release(s);
    ctx->have_triples = true;
    ctx->have_walks = false;
}

}  // namespace mtg
