// emit.cu -- step 3b: duplicate-k-mer bitvector (K6) and tig string assembly from the 2-bit store.
//
//   bitvector: write_duplication_bitvector, src/implementation/mod.rs:671-702
//   strings:   write_walks_gfa src/bin.rs:667-818, write_walks_fasta src/bin.rs:466-606
//
// Both are prefix-sum + fill: every walk edge becomes a segment of the output text, an exclusive
// scan of segment lengths gives its byte offset, and each thread then produces 16 consecutive
// output bytes (one 128-bit store): the segments of a CTA's 4 KiB of output are looked up once and staged in
// shared memory, bases are decoded from 64-bit reads of the 2-bit store.
#include <algorithm>
#include <memory>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;
constexpr int CHUNK = 16;

std::string gfa_header(const mtg_ctx* ctx) { return "H\tKL:Z:" + std::to_string(ctx->k) + "\n"; }  // src/bin.rs:688-693

__device__ __forceinline__ u32 dec_digits(u64 v) {
    u32 d = 1;
    while (v >= 10) {
        v /= 10;
        d++;
    }
    return d;
}

struct WalkView {
    const u32* edges;     // [W]
    const u64* limits;    // [T] end offsets
    const u32* dummy_w;   // weights of dummy edges
    const u32* unitig_w;  // [U]
    const u64* seq_off;   // [U+1]
    u64 W, T, E;          // E = 2U: ids >= E are dummies
    u32 k;
};

__device__ __forceinline__ u64 tig_of(const WalkView& w, u64 j) {  // first t with limits[t] > j
    u64 lo = 0, hi = w.T;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (w.limits[mid] > j) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// mode 0: bitvector, 1: GFA, 2: FASTA.  seg_len[j] = bytes walk edge j contributes (incl. header / newline).
__global__ void __launch_bounds__(TB) segment_lengths(WalkView w, int mode, u32* __restrict__ seg_len, u32* __restrict__ seg_tig) {
    u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= w.W) return;
    const u64 t = tig_of(w, j);
    const u64 start = t ? w.limits[t - 1] : 0, end = w.limits[t];
    const u32 e = w.edges[j];
    const bool dummy = e >= w.E;
    u32 len;
    if (mode == 0) {
        len = dummy ? w.dummy_w[e - w.E] : w.unitig_w[e >> 1];
    } else if (dummy) {
        len = 0;
    } else {
        const u32 full = w.unitig_w[e >> 1] + w.k - 1;
        if (j == start) {
            len = full + (mode == 1 ? 3 : 2) + dec_digits(t + 1);  // "S\t<i>\t" / "><i>\n"
        } else {
            const u32 pe = w.edges[j - 1];
            const u32 skip = pe >= w.E ? w.k - 1 - w.dummy_w[pe - w.E] : w.k - 1;  // src/bin.rs:745-749
            len = full - skip;
        }
    }
    if (j + 1 == end) len += 1;  // '\n'
    seg_len[j] = len;
    seg_tig[j] = (u32)t;
}

__device__ __forceinline__ u64 segment_of(const u64* __restrict__ seg_off, u64 W, u64 q) {  // last j with seg_off[j] <= q
    u64 lo = 0, hi = W;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (seg_off[mid] <= q) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

// up to 32 consecutive bases starting at `pos`, base i at bits [2i, 2i+2)
__device__ __forceinline__ u64 read_bases64(const u64* __restrict__ words, u64 pos) {
    const u64 w = pos >> 5;
    const u32 s = (u32)(pos & 31) * 2;
    u64 x = words[w] >> s;
    if (s) x |= words[w + 1] << (64 - s);
    return x;
}
__device__ __forceinline__ char base_char(u32 c) { return (char)((0x47544341u >> (8 * (c & 3u))) & 0xFFu); }  // "ACTG"

constexpr int SEG_CACHE = 512;  // segments of one CTA's output range whose description is staged in shared memory

// Everything fill_text needs to know about one segment (= one walk edge's share of the text), gathered once per segment.
struct SegMeta {
    u64 s0;     // first output byte
    u64 first;  // forward: position of the first body character in the 2-bit store; backward: one past the position of it
    u32 len;    // bytes incl. header and trailing newline
    u32 tig;    // 0-based index of the tig (header digits)
    u32 flags;  // bits 0-7: header bytes; 8: dummy edge; 9: backward edge; 10: ends with '\n'
};
__device__ __forceinline__ SegMeta seg_meta(const WalkView& w, int mode, const u64* __restrict__ seg_off, const u32* __restrict__ seg_len,
                                            const u32* __restrict__ seg_tig, u64 j) {
    SegMeta m;
    m.s0 = seg_off[j];
    m.len = seg_len[j];
    const u32 e = w.edges[j];
    const bool dummy = e >= w.E;
    const u64 t = seg_tig[j];
    m.tig = (u32)t;
    const bool last = (j + 1 == w.limits[t]);
    u32 hdr = 0, skip = 0;
    if (mode != 0 && !dummy) {
        const u64 tstart = t ? w.limits[t - 1] : 0;
        if (j == tstart) hdr = (mode == 1 ? 3 : 2) + dec_digits(t + 1);
        else {
            const u32 pe = w.edges[j - 1];
            skip = pe >= w.E ? w.k - 1 - w.dummy_w[pe - w.E] : w.k - 1;  // src/bin.rs:745-749
        }
    }
    m.first = 0;
    if (mode != 0 && !dummy) m.first = (e & 1) ? w.seq_off[(e >> 1) + 1] - skip : w.seq_off[e >> 1] + skip;
    m.flags = hdr | (dummy ? 256u : 0u) | ((e & 1) ? 512u : 0u) | (last ? 1024u : 0u);
    return m;
}

// first[b] = the segment that holds the first output byte of CTA b of fill_text (byte q_base + b * TB * CHUNK): every
// non-empty segment announces itself to the CTAs that start inside it (almost always none or one), so that no CTA has to
// binary-search the offsets -- a chain of ~20 dependent global loads in front of a barrier, which used to BE the kernel.
__global__ void __launch_bounds__(TB)
    cta_first_segments(const u64* __restrict__ seg_off, const u32* __restrict__ seg_len, u64 W, u64 q_base, u64 q_end, u32* __restrict__ first) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= W) return;
    const u64 s0 = seg_off[j], s1 = s0 + seg_len[j];
    if (s1 <= s0 || s1 <= q_base || s0 >= q_end) return;
    const u64 span = (u64)TB * CHUNK;
    u64 b = s0 <= q_base ? 0 : (s0 - q_base + span - 1) / span;
    for (; q_base + b * span < s1 && q_base + b * span < q_end; b++) first[b] = (u32)j;
}

// Every CTA produces TB * CHUNK consecutive output bytes.  The segments that intersect that range (from the first segment
// of this CTA to the first segment of the next) are described once per segment into shared memory (edge, tig, header / overlap lengths, position in
// the 2-bit store: a dozen dependent global loads per segment instead of per 16 output bytes); a thread then locates its
// first segment there and decodes its 16 bytes from 64-bit reads of the 2-bit store (up to 32 bases per read).
__global__ void __launch_bounds__(TB)
    fill_text(WalkView w, int mode, const u64* __restrict__ seg_off, const u32* __restrict__ seg_len, const u32* __restrict__ seg_tig,
              const u32* __restrict__ first, const u64* __restrict__ words, u64 q_base, u64 total, char* __restrict__ out) {
    // produces bytes [q_base, total) of the text into out[0 ..): a rank's share of the output, or all of it
    __shared__ SegMeta s_meta[SEG_CACHE];
    // four bases (one byte of the 2-bit store) -> four characters: forward, and reverse complement (characters reversed)
    __shared__ u32 s_fwd[256], s_rc[256];
    static_assert(TB == 256, "one table entry per thread");
    {
        const u32 b = threadIdx.x;
        u32 f = 0, r = 0;
#pragma unroll
        for (u32 i = 0; i < 4; i++) {
            const u32 c = (b >> (2 * i)) & 3u;
            f |= (u32)(u8)base_char(c) << (8 * i);
            r |= (u32)(u8)base_char(c ^ 2u) << (8 * (3 - i));
        }
        s_fwd[b] = f;
        s_rc[b] = r;
    }
    const u64 cta_q0 = q_base + (u64)blockIdx.x * TB * CHUNK;
    const u64 j0 = first[blockIdx.x];
    const u64 j1 = blockIdx.x + 1 < gridDim.x ? (u64)first[blockIdx.x + 1] : w.W - 1;
    const u32 n_seg = (u32)min((u64)SEG_CACHE + 1, j1 - j0 + 1);
    const bool cached = n_seg <= SEG_CACHE;
    if (cached)
        for (u32 i = threadIdx.x; i < n_seg; i += TB) s_meta[i] = seg_meta(w, mode, seg_off, seg_len, seg_tig, j0 + i);
    __syncthreads();
    const u64 q0 = cta_q0 + (u64)threadIdx.x * CHUNK;
    if (q0 >= total) return;
    u64 j;
    if (cached) {  // last cached segment with offset <= q0
        u32 lo = 0, hi = n_seg;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (s_meta[mid].s0 <= q0) lo = mid + 1;
            else hi = mid;
        }
        j = j0 + lo - 1;
    } else {
        j = segment_of(seg_off, w.W, q0);
    }
    alignas(16) char buf[CHUNK];
    u64 q = q0;
    const u64 qend = min(q0 + (u64)CHUNK, total);
    while (q < qend) {
        SegMeta m = cached ? s_meta[j - j0] : seg_meta(w, mode, seg_off, seg_len, seg_tig, j);
        while (m.s0 + m.len <= q) {  // skips empty segments
            j++;
            m = cached ? s_meta[j - j0] : seg_meta(w, mode, seg_off, seg_len, seg_tig, j);
        }
        const u64 s0 = m.s0;
        const u32 len = m.len, hdr = m.flags & 255u;
        const bool dummy = (m.flags & 256u) != 0, backward = (m.flags & 512u) != 0, last = (m.flags & 1024u) != 0;
        const u64 stop = min(qend, s0 + len);             // this thread's bytes of the segment: [q, stop)
        const u32 body_end = len - (last ? 1u : 0u);      // segment-relative index of the trailing '\n', if any
        if (CHUNK == 16 && q == q0 && qend - q0 == CHUNK && (u32)(q - s0) >= hdr && q0 + CHUNK <= s0 + body_end) {
            // the common case: all 16 bytes are body characters of one segment -- four table look-ups, one 16-byte store
            uint4 v;
            if (mode == 0) {
                const u32 c = dummy ? 0x30303030u : 0x31313131u;
                v = make_uint4(c, c, c, c);
            } else if (!backward) {
                const u64 bits = read_bases64(words, m.first + ((u32)(q - s0) - hdr));
                v = make_uint4(s_fwd[bits & 255u], s_fwd[(bits >> 8) & 255u], s_fwd[(bits >> 16) & 255u], s_fwd[(bits >> 24) & 255u]);
            } else {
                const u64 bits = read_bases64(words, m.first - ((u32)(q - s0) - hdr) - CHUNK);
                v = make_uint4(s_rc[(bits >> 24) & 255u], s_rc[(bits >> 16) & 255u], s_rc[(bits >> 8) & 255u], s_rc[bits & 255u]);
            }
            *reinterpret_cast<uint4*>(out + (q0 - q_base)) = v;
            return;
        }
        // header characters: "S\t<i>\t" (GFA) / "><i>\n" (FASTA)
        for (; q < stop && (u32)(q - s0) < hdr; q++) {
            const u32 c = (u32)(q - s0);
            char ch;
            if (c == 0) ch = mode == 1 ? 'S' : '>';
            else if (c == hdr - 1) ch = mode == 1 ? '\t' : '\n';
            else if (mode == 1 && c == 1) ch = '\t';
            else {
                u64 v = (u64)m.tig + 1;
                for (u32 r = hdr - 2 - c; r > 0; r--) v /= 10;
                ch = (char)('0' + v % 10);
            }
            buf[q - q0] = ch;
        }
        // body: '1' / '0' of the bitvector, or bases of the oriented unitig
        if (q < stop && (u32)(q - s0) < body_end) {
            const u32 c0 = (u32)(q - s0);
            const u32 nb = (u32)(min(stop, s0 + body_end) - q);  // 1..16 body characters
            if (mode == 0) {
                const char ch = dummy ? '0' : '1';
                for (u32 i = 0; i < nb; i++) buf[q - q0 + i] = ch;
            } else if (!backward) {
                u64 bits = read_bases64(words, m.first + (c0 - hdr));
                for (u32 i = 0; i < nb; i++, bits >>= 2) buf[q - q0 + i] = base_char((u32)bits);
            } else {
                // reverse complement: oriented index b <-> stored position first - 1 - b, complement = code ^ 2
                u64 bits = read_bases64(words, m.first - (c0 - hdr) - nb);  // from the stored position of the LAST character of this run
                for (u32 i = 0; i < nb; i++, bits >>= 2) buf[q - q0 + (nb - 1 - i)] = base_char((u32)bits ^ 2u);
            }
            q += nb;
        }
        if (q < stop) {  // only the newline is left
            buf[q - q0] = '\n';
            q++;
        }
    }
    if (qend - q0 == CHUNK) {
        *reinterpret_cast<uint4*>(out + (q0 - q_base)) = *reinterpret_cast<const uint4*>(buf);
    } else {
        for (u64 i = 0; i < qend - q0; i++) out[q0 - q_base + i] = buf[i];
    }
}

// byte range of the walks [lo, hi): bounds[0] = first byte, bounds[1] = one past the last, bounds[2] = length of the whole text
__global__ void range_bounds(const u64* __restrict__ limits, const u64* __restrict__ seg_off, u64 W, u64 lo, u64 hi, u64* __restrict__ bounds) {
    bounds[0] = lo ? seg_off[limits[lo - 1]] : 0;
    bounds[1] = hi ? seg_off[limits[hi - 1]] : 0;
    bounds[2] = seg_off[W];
}

// Produces the text of `mode` -- or the share of it that belongs to the walks [walk_lo, walk_hi) -- into the context's
// pinned staging buffer (full-speed DMA) and returns the length of what was produced.  out != nullptr: additionally
// copied to the caller's buffer.  size_only: nothing is materialised.  `where` (optional): [0] = offset of the produced
// bytes inside the whole text, [1] = length of the whole text (both including `prefix`).
u64 emit(mtg_ctx* ctx, int mode, const char* prefix, size_t prefix_len, char* out, u64 cap, bool size_only, const char** view,
         u64 walk_lo = 0, u64 walk_hi = ~0ull, u64* where = nullptr) {
    MTG_REQUIRE(ctx->have_walks, MTG_ERR_INVALID, "no walks: call mtg_finish_walks first");
    MTG_REQUIRE(mode == 0 || ctx->have_seqs, MTG_ERR_INVALID, "sequences were not supplied: tig strings cannot be assembled");
    cudaStream_t s = ctx->stream;
    WalkView w{};
    w.edges = ctx->d_walk_edges.p;
    w.limits = ctx->d_walk_limits.p;
    w.dummy_w = ctx->d_dummy_w.p;
    w.unitig_w = ctx->unitig_w.p;
    w.seq_off = ctx->seq_off.p;
    w.W = ctx->n_walk_edges_dev;
    w.T = ctx->n_walks_dev;
    w.E = ctx->E;
    w.k = ctx->k;
    walk_hi = std::min<u64>(walk_hi, w.T);
    MTG_REQUIRE(walk_lo <= walk_hi, MTG_ERR_INVALID, "bad walk range");
    if (walk_lo != 0) prefix_len = 0;  // the header line belongs to the share that starts the text
    PinnedBuf& stage = ctx->text_stage[mode];
    const double t_begin = wall_ms();
    double t_scan = t_begin, t_alloc = t_begin, t_launched = t_begin;
    u64 h_bounds[3] = {0, 0, 0};
    DBuf<u32> seg_len, seg_tig;
    DBuf<u64> seg_off, bounds;
    if (w.W) {
        seg_len.resize(w.W, s);
        seg_tig.resize(w.W, s);
        seg_off.resize(w.W + 1, s);
        bounds.resize(3, s);
        MTG_LAUNCH(ctx, segment_lengths, grid_for(w.W, TB), TB, 0, w, mode, seg_len.p, seg_tig.p);
        exclusive_sum_u32_to_u64(ctx, seg_len.p, seg_off.p, w.W, seg_off.p + w.W);
        MTG_LAUNCH(ctx, range_bounds, 1, 1, 0, w.limits, seg_off.p, w.W, walk_lo, walk_hi, bounds.p);
        MTG_CUDA(cudaMemcpyAsync(h_bounds, bounds.p, sizeof(h_bounds), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        t_scan = wall_ms();
    }
    const u64 b0 = h_bounds[0], b1 = h_bounds[1], part = b1 - b0;
    if (where) {
        where[0] = walk_lo ? b0 + (mode == 1 ? gfa_header(ctx).size() : 0) : 0;
        where[1] = h_bounds[2] + (mode == 1 ? gfa_header(ctx).size() : 0);
    }
    if (!size_only) {
        MTG_REQUIRE(!out || cap >= part + prefix_len, MTG_ERR_INVALID, "output buffer too small");
        stage.ensure(part + prefix_len + 1);
        if (prefix_len) memcpy(stage.p, prefix, prefix_len);
        if (part) {
            DBuf<char> out_buf;
            out_buf.resize(part + CHUNK, s);
            char* d_out = out_buf.p;
            u64 chunks = (part + CHUNK - 1) / CHUNK;
            const dim3 grid = grid_for(chunks, TB);
            DBuf<u32> first;
            first.resize(grid.x + 1, s);
            t_alloc = wall_ms();
            MTG_LAUNCH(ctx, cta_first_segments, grid_for(w.W, TB), TB, 0, seg_off.p, seg_len.p, w.W, b0, b1, first.p);
            MTG_LAUNCH(ctx, fill_text, grid, TB, 0, w, mode, seg_off.p, seg_len.p, seg_tig.p, first.p, ctx->seq_words.p, b0, b1, d_out);
            MTG_CUDA(cudaMemcpyAsync(stage.p + prefix_len, d_out, part, cudaMemcpyDeviceToHost, s));
            t_launched = wall_ms();
            MTG_CUDA(cudaStreamSynchronize(s));
            if (trace_slow_calls() && wall_ms() - t_begin > 50.0)
                fprintf(stderr, "[mtg trace] emit mode %d, %llu bytes: scan %.1f ms, allocations %.1f ms, launches %.1f ms, wait %.1f ms\n", mode,
                        (unsigned long long)part, t_scan - t_begin, t_alloc - t_scan, t_launched - t_alloc, wall_ms() - t_launched);
        }
        if (out && part + prefix_len) memcpy(out, stage.p, part + prefix_len);
        if (view) *view = stage.p;
    }
    return part + prefix_len;
}

}  // namespace

u64 dup_bitvector(mtg_ctx* ctx, char* out, u64 cap, bool size_only, const char** view, u64 walk_lo, u64 walk_hi, u64* where) {
    return emit(ctx, 0, nullptr, 0, out, cap, size_only, view, walk_lo, walk_hi, where);
}

u64 assemble_tigs(mtg_ctx* ctx, int format, char* out, u64 cap, bool size_only, const char** view, u64 walk_lo, u64 walk_hi, u64* where) {
    MTG_REQUIRE(format == MTG_FORMAT_GFA || format == MTG_FORMAT_FASTA, MTG_ERR_INVALID, "unknown text format");
    if (format == MTG_FORMAT_GFA) {
        const std::string header = gfa_header(ctx);
        return emit(ctx, 1, header.data(), header.size(), out, cap, size_only, view, walk_lo, walk_hi, where);
    }
    return emit(ctx, 2, nullptr, 0, out, cap, size_only, view, walk_lo, walk_hi, where);
}

}  // namespace mtg
