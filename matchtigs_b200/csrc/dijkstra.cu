// dijkstra.cu -- step 2: many-source bounded Dijkstra (K4).
//
// Replaces Dijkstra::shortest_path_lens as called from greedytigs/mod.rs:324-335 (semantics:
// SURVEY.md A.1): weights are positive integers and the search stops at distance k-1, so the
// heap degenerates to <= k distance levels ("Dial").  Settle order is the total order
// (dist, node id); the kernel therefore produces, per source, the targets of each level sorted
// by node id, level after level, until `cap` records exist.
//
// Tier 0 -- one THREAD per source (dijkstra_thread_kernel): almost every search labels a few dozen
//   nodes, so each lane runs its own search over a 48-entry lane-interleaved slice of shared memory
//   (extract-min = scan for the smallest (dist, node), relaxation = linear search); a warp then
//   carries 32 independent gather chains instead of one.
// Tier 1 -- one warp per source for the searches that outgrow tier 0, persistent CTAs pulling
//   work from an atomic counter:
//   * distance labels live in a per-warp open-addressing table in shared memory
//     (key = node id, value = tentative distance, atomicMin relaxations);
//   * the frontier is the insertion-ordered slot list of that table; a 64-bit mask of
//     non-empty distance levels (warp OR-reduction) replaces the priority queue;
//   * lanes settle the nodes of the current level in parallel: one row_ptr pair + target-bit
//     probe + the (col, weight) row per settled node are the only global loads.
// Tier 2 -- searches that outgrow the shared-memory table (or a level with > 64 targets)
//   rerun with one CTA per source against dense global-memory labels.
#include <cooperative_groups.h>

#include <algorithm>

#include "mtg_internal.cuh"

namespace cg = cooperative_groups;

namespace mtg {

namespace {

constexpr int DJ_WARPS = 4;
constexpr int DJ_THREADS = DJ_WARPS * 32;
constexpr int DJ_TABLE = 512;         // slots per warp
constexpr int DJ_TABLE_BITS = 9;
constexpr int DJ_MAX_ENTRIES = 320;   // distinct labelled nodes per search before tier 2
constexpr int DJ_LVL_TARGETS = 64;    // targets of one distance level kept for sorting
constexpr int DJ_CHUNK = 1;           // work items fetched per atomic (tier 1 only sees the overflow of tier 0)
constexpr u32 EMPTY = 0xFFFFFFFFu;
constexpr u32 META_TRUNC = 0x80000000u;
constexpr u32 META_OVERFLOW = 0x40000000u;

struct WarpState {
    u32 key[DJ_TABLE];
    u32 dist[DJ_TABLE];
    u16 slots[DJ_MAX_ENTRIES];
    u32 lvl[DJ_LVL_TARGETS];
    u32 n;         // labelled nodes
    u32 overflow;  // table or level buffer exceeded
};

struct SearchArgs {
    const u32* row_s;
    const u32* col_s;
    const u8* w_s;
    const u32* bitmap;     // target bits (initial map or the current open map)
    const u32* sources;    // node id of every source
    const u32* work_list;  // optional: global source index per work item
    u64 n_work;
    const u32* todo;       // optional: the work items this launch processes (tier overflow lists)
    u64 n_items;           // number of work items of this launch
    u32 shard_rank, shard_count;
    u32 cap, max_weight;
    u32 tie_flip;          // 0: ties inside a distance level go to the smaller node id (P1); 0xFFFFFFFF: to the larger
    u64* records;          // [n_work * cap]
    u32* meta;             // [n_work]
    u32* overflow_list;    // work items for tier 2
    u32* overflow_count;
    DevStats* stats;
    unsigned long long* work_counter;
};

__device__ __forceinline__ u32 hash_slot(u32 v) { return (v * 0x9E3779B1u) >> (32 - DJ_TABLE_BITS); }

// Returns true if `u` received a smaller label.  Sets *ovf on table exhaustion.
__device__ __forceinline__ bool table_insert_min(WarpState* ws, u32 u, u32 nw) {
    u32 slot = hash_slot(u);
    for (int probes = 0; probes < DJ_TABLE; probes++) {
        u32 cur = ((volatile u32*)ws->key)[slot];
        if (cur == EMPTY) {
            u32 prev = atomicCAS(&ws->key[slot], EMPTY, u);
            if (prev == EMPTY) {
                atomicMin(&ws->dist[slot], nw);
                u32 idx = atomicAdd(&ws->n, 1u);
                if (idx < DJ_MAX_ENTRIES) ws->slots[idx] = (u16)slot;
                else ws->overflow = 1;
                return true;
            }
            cur = prev;
        }
        if (cur == u) return atomicMin(&ws->dist[slot], nw) > nw;
        slot = (slot + 1) & (DJ_TABLE - 1);
    }
    ws->overflow = 1;
    return false;
}

__device__ __forceinline__ u64 warp_or64(u64 v) {
    u32 lo = __reduce_or_sync(0xffffffffu, (u32)v);
    u32 hi = __reduce_or_sync(0xffffffffu, (u32)(v >> 32));
    return ((u64)hi << 32) | lo;
}

// ---------------- tier 0: one THREAD per source ----------------
// Most searches label a few dozen nodes and settle one node per distance level, so a whole warp per
// source idles 31 lanes and pays every per-level warp operation for nothing.  Here every lane runs its own
// search; 32 independent load chains per warp hide the L2 latency that a single chain cannot.  Labels live in a
// per-thread slice of shared memory (lane-interleaved, so equal indices never conflict):
//   [0, ns)  settled labels, [ns, n)  open labels -- extract-min scans the open ones only (the frontier of these
//   searches is a handful of nodes) for the smallest (dist, node id), exactly the heap's pop order, and swaps the
//   winner to position ns;
//   a relaxed edge first asks a 64-bit signature of the labelled node ids (one bit per hash class): most relaxed
//   edges lead to a node not seen before and skip the linear search of the labels altogether.
// Searches that outgrow T0_ENTRIES labels go to the warp tier.
#ifndef MTG_T0_THREADS
#define MTG_T0_THREADS 128
#endif
constexpr int T0_THREADS = MTG_T0_THREADS;
#ifndef MTG_T0_ENTRIES
#define MTG_T0_ENTRIES 48
#endif
constexpr int T0_ENTRIES = MTG_T0_ENTRIES;
constexpr int T0_HASH = 64;           // hash slots per thread (power of two, > T0_ENTRIES)
static_assert(T0_ENTRIES < T0_HASH && T0_ENTRIES < 255, "label indices are bytes and the table must keep free slots");
__device__ __forceinline__ u32 t0_hash(u32 v) { return (v * 0x9E3779B1u) >> 26; }

// PACKED (graphs with at most 2^26 nodes; distances are < 64): a label is ONE word, distance << 26 | node id, so the
// extract-min compares 32-bit words from one array and the distance array (and its shared memory: one more CTA per SM)
// goes away.  The unpacked variant serves bigger graphs.
constexpr u32 T0_ID_BITS = 26, T0_ID_MASK = (1u << T0_ID_BITS) - 1u;
template <bool PACKED>
__global__ void __launch_bounds__(T0_THREADS) dijkstra_thread_kernel(SearchArgs a) {
    __shared__ u32 s_key[T0_ENTRIES][T0_THREADS];   // labelled node ids, in insertion order (labels never move)
    __shared__ u8 s_dist[PACKED ? 1 : T0_ENTRIES][T0_THREADS];   // tentative distance of an open label, 0 once settled
    __shared__ u8 s_open[T0_ENTRIES][T0_THREADS];   // indices of the open labels (unordered)
    __shared__ u32 s_hash32[T0_HASH / 4][T0_THREADS];  // node id -> label index + 1 (open addressing, 0 = empty): one byte per
                                                       // slot, four slots of a thread per word so that a table is cleared
                                                       // with 16 stores; every access of a lane stays in its own bank
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const u32 flip = PACKED ? (a.tie_flip & T0_ID_MASK) : a.tie_flip;
    auto hash_slot_of = [&](u32 h) -> u8& { return reinterpret_cast<u8*>(&s_hash32[h >> 2][tid])[h & 3]; };
    unsigned long long st_settled = 0, st_relaxed = 0, st_cand = 0, st_searched = 0, st_trunc = 0, st_ovf = 0, st_labels = 0;
    u32 st_max_labels = 0, st_max_open = 0;
    // One flat loop: every iteration a lane either fetches its next source or settles ONE node of its current search.
    // Lanes whose search ends refill on the next iteration instead of idling until the largest search of the warp
    // is done (a nested search loop would reconverge only after all 32 searches).
    bool active = false;
    u64 t = 0;
    u32 src = 0, n = 0, n_open = 0, n_settled = 0, emitted = 0, relaxed = 0, max_open = 0;
    for (;;) {
        if (!active) {
            u64 q;
            {
                cg::coalesced_group g = cg::coalesced_threads();  // the lanes that refill together share one atomic
                unsigned long long base = 0;
                if (g.thread_rank() == 0) base = atomicAdd(a.work_counter, (unsigned long long)g.size());
                q = g.shfl(base, 0) + g.thread_rank();
            }
            if (q >= a.n_items) break;
            t = a.todo ? (u64)a.todo[q] : q;
            const u64 gi = a.work_list ? (u64)a.work_list[t] : t * a.shard_count + a.shard_rank;
            src = a.sources[gi];
            if (a.row_s[src] == a.row_s[src + 1]) {  // no traversable out-edge: the search settles the source only
                a.meta[t] = 0;
                st_settled++;
            } else {
                st_searched++;
                n = 1;
                n_open = 1;
                n_settled = 0;
                emitted = relaxed = 0;
                max_open = 1;
#pragma unroll
                for (int h = 0; h < T0_HASH / 4; h++) s_hash32[h][tid] = 0;
                s_key[0][tid] = src;
                if (!PACKED) s_dist[0][tid] = 0;
                s_open[0][tid] = 0;
                hash_slot_of(t0_hash(src)) = 1;
                active = true;
            }
        }
        if (active) {
            // extract-min over the open labels: total order (dist, node id)
            unsigned long long best = ~0ull;
            u32 bi = 0;
            for (u32 i = 0; i < n_open; i++) {
                const u32 l = s_open[i][tid];
                const unsigned long long val = PACKED ? (unsigned long long)(s_key[l][tid] ^ flip)
                                                      : ((unsigned long long)s_dist[l][tid] << 32) | (s_key[l][tid] ^ flip);
                if (val < best) {
                    best = val;
                    bi = i;
                }
            }
            const u32 d = PACKED ? (u32)best >> T0_ID_BITS : (u32)(best >> 32);
            const u32 v = PACKED ? ((u32)best ^ flip) & T0_ID_MASK : (u32)best ^ flip;
            {  // the winner leaves the open list and is marked settled (distance 0: no relaxation, nw >= 1, compares smaller)
                const u32 l = s_open[bi][tid];
                s_open[bi][tid] = s_open[--n_open][tid];
                if (PACKED) s_key[l][tid] = v;
                else s_dist[l][tid] = 0;
            }
            n_settled++;
            if (v != src && ((a.bitmap[v >> 5] >> (v & 31)) & 1u)) a.records[t * a.cap + emitted++] = (u64)v | ((u64)d << 32);
            bool overflow = false;
            const u32 e0 = a.row_s[v], e1 = a.row_s[v + 1];
            auto relax = [&](u32 u, u32 w) {
                const u32 nw = d + w;
                relaxed++;
                if (nw > a.max_weight || overflow) return;
                // label of u: a few probes of the per-thread hash table instead of a scan over all labels
                u32 h = t0_hash(u), j;
                while ((j = hash_slot_of(h)) != 0 && (PACKED ? s_key[j - 1][tid] & T0_ID_MASK : s_key[j - 1][tid]) != u) h = (h + 1) & (T0_HASH - 1);
                if (j) {  // open label: decrease-key in place (settled ones hold distance 0)
                    if (PACKED) {
                        if (nw < (s_key[j - 1][tid] >> T0_ID_BITS)) s_key[j - 1][tid] = (nw << T0_ID_BITS) | u;
                    } else if (nw < s_dist[j - 1][tid]) {
                        s_dist[j - 1][tid] = (u8)nw;
                    }
                } else if (n < T0_ENTRIES) {
                    s_key[n][tid] = PACKED ? (nw << T0_ID_BITS) | u : u;
                    if (!PACKED) s_dist[n][tid] = (u8)nw;
                    s_open[n_open++][tid] = (u8)n;
                    hash_slot_of(h) = (u8)(++n);
                } else {
                    overflow = true;
                }
            };
            // A node of a de Bruijn graph has at most four out-edges: their targets and weights are requested together,
            // so the row costs one round trip to L2 instead of two per edge (weight, then target behind the bound check).
            u32 eu[4], ew[4];
#pragma unroll
            for (u32 q = 0; q < 4; q++) {
                const bool in_row = e0 + q < e1;
                eu[q] = in_row ? a.col_s[e0 + q] : 0u;
                ew[q] = in_row ? (u32)a.w_s[e0 + q] : 0u;
            }
#pragma unroll
            for (u32 q = 0; q < 4; q++)
                if (e0 + q < e1) relax(eu[q], ew[q]);
            for (u32 e = e0 + 4; e < e1 && !overflow; e++) relax(a.col_s[e], (u32)a.w_s[e]);
            max_open = max(max_open, n_open);
            if (overflow) {
                a.meta[t] = META_OVERFLOW;
                a.overflow_list[atomicAdd(a.overflow_count, 1u)] = (u32)t;
                st_ovf++;
                active = false;
            } else if (n_open == 0 || emitted == a.cap) {
                // a full list is complete only if nothing is left to settle (checked after v's own relaxation, because
                // the last target may be the only way to further ones)
                const bool truncated = n_open != 0;
                a.meta[t] = emitted | (truncated ? META_TRUNC : 0u);
                st_settled += n_settled;
                st_relaxed += relaxed;
                st_cand += emitted;
                st_trunc += truncated;
                st_labels += n;
                st_max_labels = max(st_max_labels, n);
                st_max_open = max(st_max_open, max_open);
                active = false;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        st_settled += __shfl_down_sync(0xffffffffu, st_settled, o);
        st_relaxed += __shfl_down_sync(0xffffffffu, st_relaxed, o);
        st_cand += __shfl_down_sync(0xffffffffu, st_cand, o);
        st_searched += __shfl_down_sync(0xffffffffu, st_searched, o);
        st_trunc += __shfl_down_sync(0xffffffffu, st_trunc, o);
        st_ovf += __shfl_down_sync(0xffffffffu, st_ovf, o);
        st_labels += __shfl_down_sync(0xffffffffu, st_labels, o);
        st_max_labels = max(st_max_labels, __shfl_down_sync(0xffffffffu, st_max_labels, o));
        st_max_open = max(st_max_open, __shfl_down_sync(0xffffffffu, st_max_open, o));
    }
    if (lane == 0) {
        if (st_searched) atomicAdd(&a.stats->sources_searched, st_searched);
        if (st_settled) atomicAdd(&a.stats->settled, st_settled);
        if (st_relaxed) atomicAdd(&a.stats->relaxed, st_relaxed);
        if (st_cand) atomicAdd(&a.stats->candidates, st_cand);
        if (st_trunc) atomicAdd(&a.stats->truncated, st_trunc);
        if (st_ovf) atomicAdd(&a.stats->overflow, st_ovf);
        if (st_labels) atomicAdd(&a.stats->labels, st_labels);
        if (st_max_labels) atomicMax(&a.stats->max_labels, (unsigned long long)st_max_labels);
        if (st_max_open) atomicMax(&a.stats->max_open, (unsigned long long)st_max_open);
    }
}

__global__ void __launch_bounds__(DJ_THREADS) dijkstra_warp_kernel(SearchArgs a) {
    __shared__ WarpState state[DJ_WARPS];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpState* ws = &state[warp];
    for (int i = lane; i < DJ_TABLE; i += 32) {
        ws->key[i] = EMPTY;
        ws->dist[i] = EMPTY;
    }
    if (lane == 0) {
        ws->n = 0;
        ws->overflow = 0;
    }
    __syncwarp();
    unsigned long long st_settled = 0, st_relaxed = 0, st_cand = 0, st_searched = 0, st_trunc = 0, st_ovf = 0;

    for (;;) {
        unsigned long long chunk0 = 0;
        if (lane == 0) chunk0 = atomicAdd(a.work_counter, (unsigned long long)DJ_CHUNK);
        chunk0 = __shfl_sync(0xffffffffu, chunk0, 0);
        if (chunk0 >= a.n_items) break;
        for (int c = 0; c < DJ_CHUNK; c++) {
            const u64 q = chunk0 + c;
            if (q >= a.n_items) break;
            const u64 t = a.todo ? (u64)a.todo[q] : q;
            const u64 gi = a.work_list ? (u64)a.work_list[t] : t * a.shard_count + a.shard_rank;
            const u32 src = a.sources[gi];
            u32 r0 = 0, r1 = 0;
            if (lane == 0) {
                r0 = a.row_s[src];
                r1 = a.row_s[src + 1];
            }
            r0 = __shfl_sync(0xffffffffu, r0, 0);
            r1 = __shfl_sync(0xffffffffu, r1, 0);
            if (r0 == r1) {  // no traversable out-edge: the search settles the source only
                if (lane == 0) a.meta[t] = 0;
                st_settled += (lane == 0);
                continue;
            }
            if (lane == 0) {
                u32 slot = hash_slot(src);
                ws->key[slot] = src;
                ws->dist[slot] = 0;
                ws->slots[0] = (u16)slot;
                ws->n = 1;
            }
            __syncwarp();
            u64 pending = 1ull;
            u32 emitted = 0;
            bool truncated = false;
            u64* out = a.records + t * a.cap;
            unsigned long long src_settled = 0, src_relaxed = 0;
            while (pending) {
                const u32 d = (u32)__ffsll((long long)pending) - 1;
                pending &= pending - 1;
                const u32 n_now = min(((volatile u32*)&ws->n)[0], (u32)DJ_MAX_ENTRIES);
                u32 lvl_cnt = 0;
                u64 newmask = 0;
                for (u32 i0 = 0; i0 < n_now; i0 += 32) {
                    const u32 i = i0 + lane;
                    bool active = false;
                    u32 v = 0;
                    if (i < n_now) {
                        u32 slot = ((volatile u16*)ws->slots)[i];
                        active = ((volatile u32*)ws->dist)[slot] == d;
                        v = ((volatile u32*)ws->key)[slot];
                    }
                    bool is_t = false;
                    u32 e0 = 0, e1 = 0;
                    if (active) {
                        e0 = a.row_s[v];
                        e1 = a.row_s[v + 1];
                        is_t = v != src && ((a.bitmap[v >> 5] >> (v & 31)) & 1u);
                        src_settled++;
                    }
                    const unsigned tb = __ballot_sync(0xffffffffu, is_t);
                    if (is_t) {
                        u32 pos = lvl_cnt + __popc(tb & ((1u << lane) - 1u));
                        if (pos < DJ_LVL_TARGETS) ws->lvl[pos] = v;
                    }
                    lvl_cnt += __popc(tb);
                    if (active) {
                        for (u32 e = e0; e < e1; e++) {
                            u32 nw = d + a.w_s[e];
                            src_relaxed++;
                            if (nw <= a.max_weight && table_insert_min(ws, a.col_s[e], nw)) newmask |= 1ull << nw;
                        }
                    }
                }
                pending |= warp_or64(newmask);
                __syncwarp();
                if (lvl_cnt > DJ_LVL_TARGETS) ws->overflow = 1;
                __syncwarp();
                if (((volatile u32*)&ws->overflow)[0]) break;
                if (lvl_cnt) {
                    // settle order inside a level is ascending node id: rank by counting (ids are distinct)
                    for (u32 e = lane; e < lvl_cnt; e += 32) {
                        u32 x = ws->lvl[e], rank = 0;
                        for (u32 j = 0; j < lvl_cnt; j++) rank += (ws->lvl[j] ^ a.tie_flip) < (x ^ a.tie_flip);
                        if (emitted + rank < a.cap) out[emitted + rank] = (u64)x | ((u64)d << 32);
                    }
                    emitted += lvl_cnt;
                    __syncwarp();
                    if (emitted >= a.cap) {
                        truncated = emitted > a.cap || pending != 0;
                        emitted = a.cap;
                        break;
                    }
                }
            }
            const bool ovf = ((volatile u32*)&ws->overflow)[0] != 0;
            if (ovf) {
                if (lane == 0) {
                    a.meta[t] = META_OVERFLOW;
                    a.overflow_list[atomicAdd(a.overflow_count, 1u)] = (u32)t;
                }
                st_ovf += (lane == 0);
                __syncwarp();
                for (int i = lane; i < DJ_TABLE; i += 32) {
                    ws->key[i] = EMPTY;
                    ws->dist[i] = EMPTY;
                }
            } else {
                st_settled += src_settled;
                st_relaxed += src_relaxed;
                if (lane == 0) a.meta[t] = emitted | (truncated ? META_TRUNC : 0u);
                st_cand += (lane == 0) ? emitted : 0;
                st_trunc += (lane == 0 && truncated);
                const u32 n_end = ((volatile u32*)&ws->n)[0];
                if (lane == 0) {
                    atomicAdd(&a.stats->labels, (unsigned long long)n_end);
                    atomicMax(&a.stats->max_labels, (unsigned long long)n_end);
                }
                for (u32 i = lane; i < n_end; i += 32) {
                    u32 slot = ws->slots[i];
                    ws->key[slot] = EMPTY;
                    ws->dist[slot] = EMPTY;
                }
            }
            __syncwarp();
            if (lane == 0) {
                ws->n = 0;
                ws->overflow = 0;
            }
            __syncwarp();
        }
    }
    // one atomic per warp and counter
    for (int o = 16; o > 0; o >>= 1) {
        st_settled += __shfl_down_sync(0xffffffffu, st_settled, o);
        st_relaxed += __shfl_down_sync(0xffffffffu, st_relaxed, o);
    }
    if (lane == 0) {
        if (st_searched) atomicAdd(&a.stats->sources_searched, st_searched);
        if (st_settled) atomicAdd(&a.stats->settled, st_settled);
        if (st_relaxed) atomicAdd(&a.stats->relaxed, st_relaxed);
        if (st_cand) atomicAdd(&a.stats->candidates, st_cand);
        if (st_trunc) atomicAdd(&a.stats->truncated, st_trunc);
        if (st_ovf) atomicAdd(&a.stats->overflow, st_ovf);
    }
}

// ---------------- tier 2: one CTA per source, dense labels in global memory ----------------
constexpr int BIG_THREADS = 256;

struct BigArgs {
    SearchArgs s;
    const u32* todo;   // work items
    u32 n_todo;
    u32* labels;       // [slots][N] tentative distances, all EMPTY between searches
    u32* visited;      // [slots][vis_cap] labelled nodes in insertion order
    u32* lvl_targets;  // [slots][vis_cap]
    u64 N, vis_cap;
    int* err;
};

__global__ void __launch_bounds__(BIG_THREADS) dijkstra_cta_kernel(BigArgs b) {
    __shared__ u32 s_n, s_lvl, s_fail;
    __shared__ unsigned long long s_pending;
    const SearchArgs& a = b.s;
    u32* lab = b.labels + (u64)blockIdx.x * b.N;
    u32* vis = b.visited + (u64)blockIdx.x * b.vis_cap;
    u32* lt = b.lvl_targets + (u64)blockIdx.x * b.vis_cap;
    unsigned long long st_settled = 0, st_relaxed = 0;
    for (u32 w = blockIdx.x; w < b.n_todo; w += gridDim.x) {
        const u64 t = b.todo[w];
        const u64 gi = a.work_list ? (u64)a.work_list[t] : t * a.shard_count + a.shard_rank;
        const u32 src = a.sources[gi];
        if (threadIdx.x == 0) {
            lab[src] = 0;
            vis[0] = src;
            s_n = 1;
            s_pending = 1ull;
            s_fail = 0;
        }
        __syncthreads();
        u32 emitted = 0;
        bool truncated = false;
        u64* out = a.records + t * a.cap;
        for (;;) {
            unsigned long long pending = s_pending;
            if (!pending) break;
            const u32 d = (u32)__ffsll((long long)pending) - 1;
            const u32 n_now = s_n;
            __syncthreads();
            if (threadIdx.x == 0) {
                s_pending = pending & (pending - 1);
                s_lvl = 0;
            }
            __syncthreads();
            unsigned long long newmask = 0;
            for (u32 i = threadIdx.x; i < n_now; i += BIG_THREADS) {
                u32 v = vis[i];
                if (lab[v] != d) continue;
                st_settled++;
                if (v != src && ((a.bitmap[v >> 5] >> (v & 31)) & 1u)) lt[atomicAdd(&s_lvl, 1u)] = v;
                u32 e0 = a.row_s[v], e1 = a.row_s[v + 1];
                for (u32 e = e0; e < e1; e++) {
                    u32 nw = d + a.w_s[e];
                    st_relaxed++;
                    if (nw > a.max_weight) continue;
                    u32 u = a.col_s[e];
                    u32 old = atomicMin(&lab[u], nw);
                    if (old > nw) {
                        newmask |= 1ull << nw;
                        if (old == EMPTY) {
                            u32 idx = atomicAdd(&s_n, 1u);
                            if (idx < b.vis_cap) vis[idx] = u;
                            else s_fail = 1;
                        }
                    }
                }
            }
            if (newmask) atomicOr(&s_pending, newmask);
            __syncthreads();
            if (s_fail) break;
            const u32 lvl_cnt = s_lvl;
            if (lvl_cnt) {
                for (u32 e = threadIdx.x; e < lvl_cnt; e += BIG_THREADS) {
                    u32 x = lt[e], rank = 0;
                    for (u32 j = 0; j < lvl_cnt && rank + emitted < a.cap; j++) rank += (lt[j] ^ a.tie_flip) < (x ^ a.tie_flip);
                    if (emitted + rank < a.cap) out[emitted + rank] = (u64)x | ((u64)d << 32);
                }
                emitted += lvl_cnt;
                if (emitted >= a.cap) {
                    truncated = emitted > a.cap || s_pending != 0;
                    emitted = a.cap;
                    break;
                }
            }
        }
        __syncthreads();
        if (s_fail) atomicExch(b.err, 1);
        const u32 n_end = min(s_n, (u32)b.vis_cap);
        for (u32 i = threadIdx.x; i < n_end; i += BIG_THREADS) lab[vis[i]] = EMPTY;
        if (threadIdx.x == 0) {
            a.meta[t] = emitted | (truncated ? META_TRUNC : 0u);
            atomicAdd(&a.stats->candidates, (unsigned long long)emitted);
            if (truncated) atomicAdd(&a.stats->truncated, 1ull);
            atomicAdd(&a.stats->labels, (unsigned long long)n_end);
            atomicMax(&a.stats->max_labels, (unsigned long long)n_end);
        }
        __syncthreads();
    }
    if (st_settled) atomicAdd(&b.s.stats->settled, st_settled);
    if (st_relaxed) atomicAdd(&b.s.stats->relaxed, st_relaxed);
}

}  // namespace

// Runs both tiers for `n_work` work items.  work_list == nullptr: item t is global source t*shard_count+shard_rank.
void run_searches(mtg_ctx* ctx, const u32* bitmap, const u32* work_list, u64 n_work, u32 shard_rank, u32 shard_count, u32 cap,
                  u64* records, u32* meta) {
    if (n_work == 0) return;
    cudaStream_t s = ctx->stream;
    MTG_REQUIRE(n_work < 0xFFFFFFFFull, MTG_ERR_UNSUPPORTED, "too many sources");
    // owning buffers: released on every exit path, including the throwing ones
    DBuf<u32> list0_buf, list1_buf, counts_buf;  // work items leaving tier 0 / tier 1; [0] |list0|, [1] |list1|
    DBuf<unsigned long long> work_counter_buf;
    list0_buf.resize(n_work, s);
    counts_buf.resize(2, s);
    counts_buf.zero(s);
    work_counter_buf.resize(1, s);
    work_counter_buf.zero(s);
    u32 *list0 = list0_buf.p, *list1 = nullptr, *counts = counts_buf.p;
    unsigned long long* work_counter = work_counter_buf.p;
    SearchArgs a{};
    a.row_s = ctx->row_s.p;
    a.col_s = ctx->col_s.p;
    a.w_s = ctx->w_s.p;
    a.bitmap = bitmap;
    a.sources = ctx->sources.p;
    a.work_list = work_list;
    a.n_work = n_work;
    a.shard_rank = shard_rank;
    a.shard_count = shard_count;
    a.cap = cap;
    a.max_weight = ctx->k - 1 - (ctx->opt.p1_exclusive_bound ? 1u : 0u);
    a.tie_flip = ctx->opt.p1_tie_desc ? 0xFFFFFFFFu : 0u;
    a.records = records;
    a.meta = meta;
    a.stats = ctx->dstats.p;
    a.work_counter = work_counter;
    // ---- tier 0: thread per source ----
    a.todo = nullptr;
    a.n_items = n_work;
    a.overflow_list = list0;
    a.overflow_count = counts;
    int occ = 0;
    // one-word labels whenever node ids fit 26 bits (MTG_T0_UNPACKED=1 forces the general kernel: tests)
    const bool packed = ctx->N <= (u64(1) << T0_ID_BITS) && a.max_weight < 64 && !getenv("MTG_T0_UNPACKED");
    if (packed) MTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dijkstra_thread_kernel<true>, T0_THREADS, 0));
    else MTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dijkstra_thread_kernel<false>, T0_THREADS, 0));
    if (occ < 1) occ = 1;
    u32 grid = (u32)std::min<u64>((n_work + T0_THREADS - 1) / T0_THREADS, (u64)ctx->num_sms * occ);  // persistent CTAs
    MTG_CUDA(cudaEventRecord(ctx->ev2, s));
    if (packed) MTG_LAUNCH(ctx, dijkstra_thread_kernel<true>, grid, T0_THREADS, 0, a);
    else MTG_LAUNCH(ctx, dijkstra_thread_kernel<false>, grid, T0_THREADS, 0, a);
    MTG_CUDA(cudaEventRecord(ctx->ev3, s));
    u32 h_counts[2] = {0, 0};
    MTG_CUDA(cudaMemcpyAsync(h_counts, counts, sizeof(u32), cudaMemcpyDeviceToHost, s));
    // the counters ride along: if no search left tier 0 (the common case) they are final and nobody has to ask again
    MTG_CUDA(cudaMemcpyAsync(&ctx->h_dstats, ctx->dstats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->h_dstats_final = h_counts[0] == 0;
    MTG_CUDA(cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev2, ctx->ev3));
    // ---- tier 1: warp per source, for searches with more than T0_ENTRIES labelled nodes ----
    if (h_counts[0]) {
        const u64 n1 = h_counts[0];
        list1_buf.resize(n1, s);
        list1 = list1_buf.p;
        MTG_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s));
        a.todo = list0;
        a.n_items = n1;
        a.overflow_list = list1;
        a.overflow_count = counts + 1;
        MTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dijkstra_warp_kernel, DJ_THREADS, 0));
        if (occ < 1) occ = 1;
        grid = (u32)std::min<u64>((n1 + DJ_WARPS - 1) / DJ_WARPS, (u64)ctx->num_sms * occ);
        MTG_LAUNCH(ctx, dijkstra_warp_kernel, grid, DJ_THREADS, 0, a);
        MTG_CUDA(cudaMemcpyAsync(h_counts + 1, counts + 1, sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
    }
    // ---- tier 2: CTA per source with dense labels in global memory ----
    if (h_counts[1]) {
        const u32 n_ovf = h_counts[1];
        const u64 N = ctx->N;
        size_t free_b = 0, total_b = 0;
        MTG_CUDA(cudaMemGetInfo(&free_b, &total_b));
        u64 per_slot = N * 12;  // labels + visited + level targets
        u64 slots = std::min<u64>(std::min<u64>((u64)n_ovf, (u64)ctx->num_sms * 2), std::max<u64>(1, (free_b / 2) / std::max<u64>(per_slot, 1)));
        DBuf<u32> labels_buf, visited_buf, lvl_buf;
        DBuf<int> err_buf;
        labels_buf.resize(slots * N, s);
        visited_buf.resize(slots * N, s);
        lvl_buf.resize(slots * N, s);
        err_buf.resize(1, s);
        labels_buf.fill_ff(s);
        err_buf.zero(s);
        u32 *labels = labels_buf.p, *visited = visited_buf.p, *lvl = lvl_buf.p;
        int* d_err = err_buf.p;
        BigArgs b{};
        b.s = a;
        b.todo = list1;
        b.n_todo = n_ovf;
        b.labels = labels;
        b.visited = visited;
        b.lvl_targets = lvl;
        b.N = N;
        b.vis_cap = N;
        b.err = d_err;
        MTG_LAUNCH(ctx, dijkstra_cta_kernel, (u32)slots, BIG_THREADS, 0, b);
        int h_err = 0;
        MTG_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        MTG_REQUIRE(h_err == 0, MTG_ERR_INTERNAL, "tier-2 search exceeded its visited list");
    }
}

void dijkstra_candidates(mtg_ctx* ctx, u32 cap, u32 shard_rank, u32 shard_count) {
    MTG_REQUIRE(ctx->have_graph, MTG_ERR_INVALID, "no graph resident: call mtg_build_graph_* first");
    MTG_REQUIRE(cap >= 1 && cap <= 4096, MTG_ERR_INVALID, "cap must be in [1, 4096]");
    MTG_REQUIRE(shard_count >= 1 && shard_rank < shard_count, MTG_ERR_INVALID, "bad shard");
    MTG_REQUIRE(ctx->k - 1 <= 63, MTG_ERR_UNSUPPORTED, "k - 1 must be <= 63 for the level mask");
    cudaStream_t s = ctx->stream;
    ctx->cap = cap;
    ctx->shard_rank = shard_rank;
    ctx->shard_count = shard_count;
    ctx->S_local = ctx->S > shard_rank ? (ctx->S - shard_rank + shard_count - 1) / shard_count : 0;
    // all ranks allocate the same padded slice so that a plain all-gather lines the slices up
    u64 padded = (ctx->S + shard_count - 1) / shard_count;
    ctx->cand.resize(std::max<u64>(padded, 1) * cap, s);
    ctx->cand_meta.resize(std::max<u64>(padded, 1), s);
    ctx->cand_meta.zero(s);
    ctx->dstats.resize(1, s);
    ctx->dstats.zero(s);
    MTG_CUDA(cudaEventRecord(ctx->ev0, s));
    run_searches(ctx, ctx->target_bits.p, nullptr, ctx->S_local, shard_rank, shard_count, cap, ctx->cand.p, ctx->cand_meta.p);
    float ms = 0;
    if (ctx->S_local == 0) {
        ctx->h_dstats = DevStats{};
        ctx->last_kernel_ms = 0;
    } else if (ctx->h_dstats_final) {  // tier 0 was everything: its synchronisation already brought the counters back
        MTG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev3));
    } else {
        MTG_CUDA(cudaEventRecord(ctx->ev1, s));
        MTG_CUDA(cudaMemcpyAsync(&ctx->h_dstats, ctx->dstats.p, sizeof(DevStats), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        MTG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    }
    const DevStats& h = ctx->h_dstats;
    ctx->stats = mtg_search_stats{};
    ctx->stats.sources_searched = h.sources_searched;
    ctx->stats.settled_nodes = h.settled;
    ctx->stats.relaxed_edges = h.relaxed;
    ctx->stats.candidates = h.candidates;
    ctx->stats.truncated_sources = h.truncated;
    ctx->stats.overflow_sources = h.overflow;
    ctx->stats.labelled_nodes = h.labels;
    ctx->stats.max_labelled_nodes = h.max_labels;
    ctx->stats.max_open_nodes = h.max_open;
    ctx->stats.dijkstra_ms = ms;
    ctx->stats.dijkstra_kernel_ms = ctx->last_kernel_ms;
    ctx->have_cand = true;
    ctx->have_triples = ctx->have_walks = false;
}

}  // namespace mtg
