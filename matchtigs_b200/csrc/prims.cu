// prims.cu -- device-wide scan and stable LSD radix sort, hand-written for sm_100a.
//
// Both are HBM-streaming kernels: tiles are sized so that a grid covers the 148 SMs several
// times over, every global access is a full 32-byte sector per thread or a coalesced row per
// warp, and all intra-tile work happens in registers / shared memory.
#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct OpAdd {
    template <class T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
    template <class T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

template <class T, class Op>
__device__ __forceinline__ T warp_inclusive(T v, Op op) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v = op(o, v);
    }
    return v;
}

// Exclusive prefix of one value per thread across the block; `total` = block aggregate.
template <class T, class Op>
__device__ __forceinline__ T block_exclusive(T v, Op op, T* total) {
    __shared__ T warp_tot[SCAN_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_inclusive(v, op);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    T prefix = 0, all = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        T t = warp_tot[w];
        if ((unsigned)w < warp) prefix = op(prefix, t);
        all = op(all, t);
    }
    T exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0;
    __syncthreads();
    *total = all;
    return op(prefix, exc);
}

template <class TIn, class TOut, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_tiles(const TIn* __restrict__ in, TOut* __restrict__ agg, size_t n) {
    Op op;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    TOut acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) acc = op(acc, (TOut)in[base + i]);
    TOut total;
    block_exclusive(acc, op, &total);
    if (threadIdx.x == 0) agg[blockIdx.x] = total;
}

template <class TIn, class TOut, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_tiles(const TIn* __restrict__ in, TOut* __restrict__ out, const TOut* __restrict__ tile_prefix, size_t n, TOut* d_total) {
    Op op;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? (TOut)in[base + i] : (TOut)0;
        acc = op(acc, v[i]);
    }
    TOut total;
    TOut run = block_exclusive(acc, op, &total);
    if (tile_prefix) run = op(run, tile_prefix[blockIdx.x]);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        TOut next = op(run, v[i]);
        if (base + i < n) {
            out[base + i] = INCLUSIVE ? next : run;
            if (d_total && base + i == n - 1) *d_total = next;
        }
        run = next;
    }
}

// Small inputs: one 1024-thread CTA, one launch instead of three -- what matters when the whole step is launch-bound
// (E. coli-sized inputs).  Every thread owns a span of consecutive elements: it reduces the span, the 1024 partial results
// are scanned across the block (two barriers in all), and the span is read again (from L1 / L2) to write its prefixes.
constexpr int SCAN1_THREADS = 1024;
constexpr size_t SCAN1_MAX = 64 * SCAN1_THREADS;

template <class TIn, class TOut, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN1_THREADS) scan_single_cta(const TIn* in, TOut* out, size_t n, TOut* d_total) {
    __shared__ TOut warp_tot[SCAN1_THREADS / 32];
    Op op;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t span = (n + SCAN1_THREADS - 1) / SCAN1_THREADS;
    const size_t lo = min((size_t)threadIdx.x * span, n), hi = min(lo + span, n);
    TOut acc = 0;
    for (size_t i = lo; i < hi; i++) acc = op(acc, (TOut)in[i]);
    const TOut inc = warp_inclusive(acc, op);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    TOut prefix = 0;
    for (unsigned w = 0; w < warp; w++) prefix = op(prefix, warp_tot[w]);
    TOut exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0;
    TOut run = op(prefix, exc);
    // in-place use (in == out) is safe: a thread only ever reads and writes its own span, reading each element before writing it
    for (size_t i = lo; i < hi; i++) {
        const TOut next = op(run, (TOut)in[i]);
        out[i] = INCLUSIVE ? next : run;
        run = next;
    }
    if (d_total && hi == n && lo < n) *d_total = run;
}

template <class TIn, class TOut, class Op, bool INCLUSIVE>
void scan_impl(mtg_ctx* ctx, const TIn* in, TOut* out, size_t n, TOut* d_total) {
    if (n == 0) {
        if (d_total) MTG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(TOut), ctx->stream));
        return;
    }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        MTG_LAUNCH(ctx, (scan_tiles<TIn, TOut, Op, INCLUSIVE>), 1, SCAN_THREADS, 0, in, out, (const TOut*)nullptr, n, d_total);
        return;
    }
    if (n <= SCAN1_MAX) {
        MTG_LAUNCH(ctx, (scan_single_cta<TIn, TOut, Op, INCLUSIVE>), 1, SCAN1_THREADS, 0, in, out, n, d_total);
        return;
    }
    DBuf<TOut> agg_buf;
    agg_buf.resize(tiles, ctx->stream);
    TOut* agg = agg_buf.p;
    MTG_LAUNCH(ctx, (scan_reduce_tiles<TIn, TOut, Op>), (unsigned)tiles, SCAN_THREADS, 0, in, agg, n);
    scan_impl<TOut, TOut, Op, false>(ctx, agg, agg, tiles, nullptr);  // in-place exclusive scan of the tile aggregates
    MTG_LAUNCH(ctx, (scan_tiles<TIn, TOut, Op, INCLUSIVE>), (unsigned)tiles, SCAN_THREADS, 0, in, out, (const TOut*)agg, n, d_total);
}

}  // namespace

void exclusive_sum_u32(mtg_ctx* ctx, const u32* in, u32* out, size_t n, u32* d_total) {
    scan_impl<u32, u32, OpAdd, false>(ctx, in, out, n, d_total);
}
void exclusive_sum_u8(mtg_ctx* ctx, const u8* in, u32* out, size_t n, u32* d_total) {
    scan_impl<u8, u32, OpAdd, false>(ctx, in, out, n, d_total);
}
void exclusive_sum_u32_to_u64(mtg_ctx* ctx, const u32* in, u64* out, size_t n, u64* d_total) {
    scan_impl<u32, u64, OpAdd, false>(ctx, in, out, n, d_total);
}
void inclusive_max_u32(mtg_ctx* ctx, const u32* in, u32* out, size_t n) {
    scan_impl<u32, u32, OpMax, true>(ctx, in, out, n, nullptr);
}

// =====================================================================================
// Stable LSD radix sort, 8-bit digits.  Per pass:
//   1. radix_hist:    per-tile digit histogram -> hist[digit][tile]
//   2. exclusive sum over hist (digit-major == output order) -> global base of every (digit, tile)
//   3. radix_scatter: stable rank of every element inside its tile (warp match_any ranking,
//                     per-warp digit counters in shared memory) + scatter.
// A tile is 8 warps x 8 rows x 32 lanes = 2048 consecutive elements; each warp owns a
// contiguous 256-element chunk so that rows are coalesced and order is preserved.
// =====================================================================================
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROWS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROWS;
constexpr u32 RS_RAW_TILES = 64;  // up to this many tiles (128 K elements) the scatter CTAs scan the count matrix themselves

template <class KW>
__global__ void __launch_bounds__(RS_THREADS) radix_hist(const KW* __restrict__ keyword, u32* __restrict__ hist, size_t n, int shift,
                                                         u32 tiles) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROWS; r++) {
        size_t i = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(u32)(keyword[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// RAW: `base_of` still holds the per-(digit, tile) counts of radix_hist; every CTA derives its own bases from them
// (used while the count matrix is small: saves the scan launch of the pass).
template <class KW, int NW, bool RAW>
__global__ void __launch_bounds__(RS_THREADS)
    radix_scatter(const KW* __restrict__ k0_in, KW* __restrict__ k0_out, const KW* __restrict__ k1_in, KW* __restrict__ k1_out,
                  const u32* __restrict__ v_in, u32* __restrict__ v_out, const u32* __restrict__ base_of, size_t n, int word,
                  int shift, u32 tiles) {
    __shared__ u32 cnt[RS_WARPS][256];
    __shared__ u32 warp_sum[RS_WARPS];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    u32 my_base = 0;  // global base of (digit == threadIdx.x, this tile)
    if (RAW) {
        static_assert(RS_THREADS == 256, "one thread per digit");
        const u32* row = base_of + (size_t)threadIdx.x * tiles;
        u32 tot = 0, before = 0;
        for (u32 t = 0; t < tiles; t++) {
            const u32 c = row[t];
            if (t < blockIdx.x) before += c;
            tot += c;
        }
        // exclusive scan of the digit totals over the 256 threads
        u32 inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        u32 pre = 0;
        for (unsigned w = 0; w < warp; w++) pre += warp_sum[w];
        my_base = pre + inc - tot + before;
    } else {
        my_base = base_of[(size_t)threadIdx.x * tiles + blockIdx.x];
    }
    __syncthreads();
    const size_t chunk = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (RS_ROWS * 32);
    KW k0[RS_ROWS], k1[RS_ROWS];
    u32 val[RS_ROWS], rank[RS_ROWS];
    u32 dig[RS_ROWS];
#pragma unroll
    for (int r = 0; r < RS_ROWS; r++) {
        size_t i = chunk + (size_t)r * 32 + lane;
        bool valid = i < n;
        unsigned vmask = __ballot_sync(0xffffffffu, valid);
        k0[r] = 0;
        k1[r] = 0;
        val[r] = 0;
        dig[r] = 0;
        rank[r] = 0;
        if (valid) {
            k0[r] = k0_in[i];
            if (NW == 2) k1[r] = k1_in[i];
            val[r] = v_in[i];
            KW kw = (NW == 2 && word == 1) ? k1[r] : k0[r];
            u32 d = (u32)(kw >> shift) & 255u;
            dig[r] = d;
            unsigned peers = __match_any_sync(vmask, d);
            unsigned lt = peers & ((1u << lane) - 1u);
            int leader = __ffs(peers) - 1;
            u32 b = 0;
            if (lt == 0) {  // lowest lane of the peer group owns the counter update
                b = cnt[warp][d];
                cnt[warp][d] = b + __popc(peers);
            }
            b = __shfl_sync(vmask, b, leader);
            rank[r] = b + __popc(lt);
        }
        __syncwarp();
    }
    __syncthreads();
    if constexpr (NW == 1) {
        // One-word keys: the tile is put in output order in shared memory first (tile-local position = first position of the
        // digit inside the tile + the warp's offset inside the digit + rank), then written out by consecutive threads:
        // elements of one digit go to consecutive global addresses, so a warp's store covers whole sectors instead of up to
        // 32 different ones.
        __shared__ KW s_key[RS_TILE];
        __shared__ u32 s_val[RS_TILE];
        __shared__ u32 lbase[256], gbase[256];
        {
            u32 run = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                u32 c = cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = run;  // offset of warp w inside the digit, tile-local
                run += c;
            }
            u32 total;
            lbase[threadIdx.x] = block_exclusive<u32>(run, OpAdd(), &total);
            gbase[threadIdx.x] = my_base;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ROWS; r++) {
            size_t i = chunk + (size_t)r * 32 + lane;
            if (i < n) {
                const u32 p = lbase[dig[r]] + cnt[warp][dig[r]] + rank[r];
                s_key[p] = k0[r];
                s_val[p] = val[r];
            }
        }
        __syncthreads();
        const size_t tile_base = (size_t)blockIdx.x * RS_TILE;
        const u32 n_tile = (u32)min((size_t)RS_TILE, n - tile_base);
        for (u32 p = threadIdx.x; p < n_tile; p += RS_THREADS) {
            const KW kw = s_key[p];
            const u32 d = (u32)(kw >> shift) & 255u;
            const size_t pos = (size_t)gbase[d] + (p - lbase[d]);
            k0_out[pos] = kw;
            v_out[pos] = s_val[p];
        }
    } else {
        {  // exclusive prefix over warps for digit == threadIdx.x, seeded with the global base of (digit, tile)
            u32 run = my_base;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                u32 c = cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ROWS; r++) {
            size_t i = chunk + (size_t)r * 32 + lane;
            if (i < n) {
                size_t pos = (size_t)cnt[warp][dig[r]] + rank[r];
                k0_out[pos] = k0[r];
                if (NW == 2) k1_out[pos] = k1[r];
                v_out[pos] = val[r];
            }
        }
    }
}

template <class KW, int NW>
int radix_sort_impl(mtg_ctx* ctx, KW* k0_a, KW* k0_b, KW* k1_a, KW* k1_b, u32* v_a, u32* v_b, size_t n, int key_bits) {
    if (n == 0 || key_bits <= 0) return 0;
    MTG_REQUIRE(n < (size_t)0xFFFFFFFFu, MTG_ERR_UNSUPPORTED, "radix sort: more than 2^32-1 elements");
    const int word_bits = (int)sizeof(KW) * 8;
    u32 tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    DBuf<u32> hist_buf;
    hist_buf.resize((size_t)256 * tiles, ctx->stream);
    u32* hist = hist_buf.p;
    int cur = 0;
    for (int bit = 0; bit < key_bits; bit += 8) {
        int word = bit / word_bits, shift = bit % word_bits;
        KW *ki0 = cur ? k0_b : k0_a, *ko0 = cur ? k0_a : k0_b;
        KW *ki1 = cur ? k1_b : k1_a, *ko1 = cur ? k1_a : k1_b;
        u32 *vi = cur ? v_b : v_a, *vo = cur ? v_a : v_b;
        const KW* digit_src = (NW == 2 && word == 1) ? ki1 : ki0;
        MTG_LAUNCH(ctx, (radix_hist<KW>), tiles, RS_THREADS, 0, digit_src, hist, n, shift, tiles);
        if (tiles <= RS_RAW_TILES) {
            MTG_LAUNCH(ctx, (radix_scatter<KW, NW, true>), tiles, RS_THREADS, 0, ki0, ko0, ki1, ko1, vi, vo, hist, n, word, shift, tiles);
        } else {
            exclusive_sum_u32(ctx, hist, hist, (size_t)256 * tiles, nullptr);
            MTG_LAUNCH(ctx, (radix_scatter<KW, NW, false>), tiles, RS_THREADS, 0, ki0, ko0, ki1, ko1, vi, vo, hist, n, word, shift, tiles);
        }
        cur ^= 1;
    }
    return cur;
}

}  // namespace

int radix_sort_pairs(mtg_ctx* ctx, u64* k0_a, u64* k0_b, u64* k1_a, u64* k1_b, u32* v_a, u32* v_b, size_t n, int nwords,
                     int key_bits) {
    if (nwords == 1) return radix_sort_impl<u64, 1>(ctx, k0_a, k0_b, nullptr, nullptr, v_a, v_b, n, key_bits);
    return radix_sort_impl<u64, 2>(ctx, k0_a, k0_b, k1_a, k1_b, v_a, v_b, n, key_bits);
}

int radix_sort_pairs_u32(mtg_ctx* ctx, u32* k_a, u32* k_b, u32* v_a, u32* v_b, size_t n, int key_bits) {
    return radix_sort_impl<u32, 1>(ctx, k_a, k_b, nullptr, nullptr, v_a, v_b, n, key_bits);
}

}  // namespace mtg
