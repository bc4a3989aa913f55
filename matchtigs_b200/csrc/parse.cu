// parse.cu -- device-side FASTA / bcalm2 record parser.
//
// The reference parses records one after the other on the host (genome-graph readers, call sites
// src/bin.rs:896-899, :907-910).  Here one thread owns 32 bytes of the file:
//   1. one read of the text: per chunk the state of its last line start (header line or not) -> inclusive max-scan over
//      the chunks = the state every chunk starts in (a two-state automaton, composed by "the latest line start wins"),
//      together with the chunk's record / sequence-byte / `L:` counts for either starting state;
//   2. the counts for the actual starting state -> three exclusive sums over the chunks;
//   3. a second read writes the record offsets, the parsed links and the bases in file order -- the bases already
//      packed to 2 bits and OR-ed into the 2-bit store, every thread counting on from its chunk's bases.
// All byte classification is done on 32-bit masks (one bit per byte of the chunk, SIMD-in-register compares), not in
// loops over the bytes.  The text is read twice, the ASCII sequence never exists in HBM, and the work arrays are ~1 byte
// per text byte.  Outputs stay on the device and feed build_graph_from_sequences / build_graph_from_links directly.
#include <algorithm>

#include "mtg_internal.cuh"

namespace mtg {

// 64-bit totals of per-chunk count arrays: the parser's offsets are 32-bit sums, this is their overflow guard.  (Kept outside
// the anonymous namespace below, whose kernels tests/test_parse_emulation.py compiles as host code.)
static __global__ void __launch_bounds__(256) sum_counts_u64(const u32* __restrict__ a, const u32* __restrict__ b, const u32* __restrict__ c, u64 n,
                                                      unsigned long long* __restrict__ out) {
    unsigned long long sa = 0, sb = 0, sc = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        sa += a[i];
        sb += b[i];
        if (c) sc += c[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_down_sync(0xffffffffu, sa, o);
        sb += __shfl_down_sync(0xffffffffu, sb, o);
        sc += __shfl_down_sync(0xffffffffu, sc, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sa) atomicAdd(out, sa);
        if (sb) atomicAdd(out + 1, sb);
        if (sc) atomicAdd(out + 2, sc);
    }
}

namespace {

constexpr int TB = 256;
constexpr int CHUNK = 32;  // text bytes per thread

__device__ __forceinline__ bool is_blank(char c) { return c == ' ' || c == '\t'; }
__device__ __forceinline__ bool is_eol(char c) { return c == '\n' || c == '\r'; }

// The chunk's bytes as eight 32-bit words in registers (two 16-byte loads where the chunk is complete and the text
// 16-byte aligned); bytes behind the end of the text read as '\n'.  Everything below works on 32-bit masks with one bit per
// byte of the chunk instead of looping over the bytes.
struct Chunk {
    u32 w[8];
    u32 valid;  // bit i: byte i exists
};
__device__ __forceinline__ Chunk load_chunk(const char* __restrict__ text, u64 L, u64 base, bool aligned) {
    Chunk c;
    const u32 n = (u32)min((u64)CHUNK, L - base);
    c.valid = n == 32 ? 0xFFFFFFFFu : (1u << n) - 1u;
    if (aligned && n == CHUNK) {
        const uint4* p = reinterpret_cast<const uint4*>(text + base);
        const uint4 a = p[0], b = p[1];
        c.w[0] = a.x, c.w[1] = a.y, c.w[2] = a.z, c.w[3] = a.w;
        c.w[4] = b.x, c.w[5] = b.y, c.w[6] = b.z, c.w[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u32 x = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u32 o = 4 * i + j;
                x |= (u32)(u8)(o < n ? text[base + o] : '\n') << (8 * j);
            }
            c.w[i] = x;
        }
    }
    return c;
}
// bit i of the result: byte i of the chunk equals `ch`
__device__ __forceinline__ u32 eq_mask(const Chunk& c, char ch) {
    const u32 bc = 0x01010101u * (u32)(u8)ch;
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 y = __vcmpeq4(c.w[i], bc) & 0x01010101u;   // 1 per equal byte, at bits 0, 8, 16, 24
        r |= (((y * 0x00204081u) >> 21) & 0xFu) << (4 * i);  // the four bits side by side (the partial products never collide)
    }
    return r;
}
__device__ __forceinline__ u32 bits_below(u32 p) { return p >= 32 ? 0xFFFFFFFFu : (1u << p) - 1u; }
__device__ __forceinline__ u32 bit_range(u32 a, u32 b) { return bits_below(b) & ~bits_below(a); }  // [a, b)

// A line is a header line iff it starts with '>'.  in_header(): bit i set iff byte i belongs to a header line, given
// whether the chunk starts inside one (`hdr0`); lines are long, so the loop over the line starts of a chunk runs 0-2 times.
struct Lines {
    u32 starts, hdr_starts;
};
__device__ __forceinline__ Lines line_starts(const Chunk& c, u32 nl, bool prev_is_nl) {
    Lines l;
    l.starts = ((nl << 1) | (prev_is_nl ? 1u : 0u)) & c.valid;
    l.hdr_starts = l.starts & eq_mask(c, '>');
    return l;
}
__device__ __forceinline__ u32 in_header(const Lines& l, bool hdr0) {
    u32 m = 0, s = l.starts, prev = 0;
    bool state = hdr0;
    while (s) {
        const u32 p = __ffs(s) - 1;
        s &= s - 1;
        if (state) m |= bit_range(prev, p);
        state = (l.hdr_starts >> p) & 1u;
        prev = p;
    }
    if (state) m |= bit_range(prev, 32);
    return m;
}
// `L:` tokens of at least 7 characters in a header line, preceded by a blank (same rules as the host reader, csrc/reader.cpp)
__device__ __forceinline__ u32 link_mask(const Chunk& c, const char* __restrict__ text, u64 L, u64 base, u32 hdr, char prev) {
    const u32 blank = eq_mask(c, ' ') | eq_mask(c, '\t');
    u32 cand = eq_mask(c, 'L') & hdr & ((blank << 1) | (is_blank(prev) ? 1u : 0u)) & c.valid;
    u32 out = 0;
    while (cand) {
        const u32 p = __ffs(cand) - 1;
        cand &= cand - 1;
        const u64 pos = base + p;
        if (pos + 6 >= L || text[pos + 1] != ':') continue;
        bool ok = true;
        for (int q = 2; q < 7; q++) {
            const char t = text[pos + q];
            if (is_blank(t) || is_eol(t)) ok = false;
        }
        if (ok) out |= 1u << p;
    }
    return out;
}

// Pass 1, one read of the text: per chunk the code of its LAST line start (0 none, 1 sequence line, 2 header line) tagged
// with the chunk index -- an inclusive max-scan carries the latest one forward, which tells every chunk whether it starts
// inside a header line -- and its record / sequence-byte / link counts, split into the part that does not depend on that
// (everything from the first line start of the chunk on) and the part in front of the first line start, which counts as
// sequence if the chunk starts inside a sequence line and may hold links if it starts inside a header line.
//   packed = records | seq_after << 6 | seq_before << 12 | links_after << 18 | links_before << 21
__global__ void __launch_bounds__(TB)
    chunk_scan_text(const char* __restrict__ text, u64 L, u64 n_chunks, bool aligned, int bcalm, u32* __restrict__ key, u32* __restrict__ packed) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    const u64 base = j * CHUNK;
    const Chunk c = load_chunk(text, L, base, aligned);
    const char prev = base ? text[base - 1] : '\n';
    const u32 nl = eq_mask(c, '\n');
    const Lines l = line_starts(c, nl, prev == '\n');
    u32 code = 0;
    if (l.starts) code = ((l.hdr_starts >> (31 - __clz(l.starts))) & 1u) ? 2u : 1u;
    key[j] = code ? (u32)((j + 1) << 2) | code : 0u;
    const u32 first = l.starts ? (u32)__ffs(l.starts) - 1 : 32u;
    const u32 not_eol = ~(nl | eq_mask(c, '\r')) & c.valid;
    const u32 before = bits_below(first);
    const u32 hdr_after = in_header(l, false);  // exact from the first line start on
    u32 links_after = 0, links_before = 0;
    if (bcalm) {
        // a chunk that starts inside a header line: everything in front of its first line start is header
        const u32 lm = link_mask(c, text, L, base, hdr_after | before, prev);
        links_after = __popc(lm & ~before);
        links_before = __popc(lm & before);
    }
    packed[j] = (u32)__popc(l.hdr_starts) | ((u32)__popc(not_eol & ~hdr_after & ~before) << 6) | ((u32)__popc(not_eol & before) << 12) |
                (links_after << 18) | (links_before << 21);
}

// counts per chunk once the state every chunk starts in is known (key_scan = inclusive max-scan of the keys)
__global__ void __launch_bounds__(TB)
    resolve_counts(const u32* __restrict__ key_scan, const u32* __restrict__ packed, u64 n_chunks, u32* __restrict__ n_rec,
                   u32* __restrict__ n_seq, u32* __restrict__ n_link) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    const bool hdr0 = j ? (key_scan[j - 1] & 3u) == 2u : false;
    const u32 p = packed[j];
    n_rec[j] = p & 63u;
    n_seq[j] = ((p >> 6) & 63u) + (hdr0 ? 0u : (p >> 12) & 63u);
    n_link[j] = ((p >> 18) & 7u) + (hdr0 ? (p >> 21) & 7u : 0u);
}

// Pass 2, the second and last read of the text: record offsets, links, and the bases -- packed to 2 bits right here
// (A0 C1 T2 G3, the code is (c >> 1) & 3) and OR-ed into the 2-bit store at the chunk's base offset, so the ASCII sequence
// never exists in HBM.
__global__ void __launch_bounds__(TB)
    chunk_scatter(const char* __restrict__ text, u64 L, u64 n_chunks, bool aligned, const u32* __restrict__ key_scan, int bcalm,
                  const u32* __restrict__ rbase, const u32* __restrict__ sbase, const u32* __restrict__ lbase,
                  unsigned long long* __restrict__ seq_words, u64* __restrict__ offsets, u64* __restrict__ link_a, u8* __restrict__ strand_a,
                  u64* __restrict__ link_b, u8* __restrict__ strand_b, int* __restrict__ err) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    const u64 base = j * CHUNK;
    const Chunk c = load_chunk(text, L, base, aligned);
    const char prev = base ? text[base - 1] : '\n';
    const bool hdr0 = j ? (key_scan[j - 1] & 3u) == 2u : false;
    const u32 nl = eq_mask(c, '\n');
    const Lines l = line_starts(c, nl, prev == '\n');
    const u32 hdr = in_header(l, hdr0);
    const u32 seq = ~hdr & ~(nl | eq_mask(c, '\r')) & c.valid;
    const u32 r0 = rbase[j], s0 = sbase[j];
    if (seq && r0 == 0 && !(l.hdr_starts && (u32)__ffs(l.hdr_starts) - 1 < (u32)__ffs(seq) - 1)) atomicExch(err, 1);  // sequence data before the first header
    // ---- records ----
    for (u32 h = l.hdr_starts, idx = 0; h; h &= h - 1, idx++) {
        const u32 p = __ffs(h) - 1;
        const u32 r = r0 + idx;
        offsets[r] = (u64)s0 + __popc(seq & bits_below(p));
        if (bcalm) {  // the record id must equal its position
            u64 q = base + p + 1, id = 0;
            bool any = false;
            while (q < L && text[q] >= '0' && text[q] <= '9') {
                id = id * 10 + (u64)(text[q] - '0');
                q++;
                any = true;
            }
            if (!any || id != r) atomicExch(err, 2);
        }
    }
    // ---- links: L:<+/->:<id>:<+/-> ----
    if (bcalm && hdr) {
        u32 lk = lbase[j];
        for (u32 m = link_mask(c, text, L, base, hdr, prev); m; m &= m - 1) {
            const u32 p = __ffs(m) - 1;
            const u64 i = base + p;
            const char s = text[i + 2];
            u64 q = i + 4, n = 0;
            bool digits = false;
            while (q < L && text[q] >= '0' && text[q] <= '9') {
                n = n * 10 + (u64)(text[q] - '0');
                q++;
                digits = true;
            }
            const bool shape = text[i + 3] == ':' && digits && q + 1 < L && text[q] == ':' && !is_blank(text[q + 1]) && !is_eol(text[q + 1]);
            const char t = shape ? text[q + 1] : '?';
            if (!shape) atomicExch(err, 3);
            else if ((s != '+' && s != '-') || (t != '+' && t != '-')) atomicExch(err, 4);
            link_a[lk] = (u64)r0 + __popc(l.hdr_starts & bits_below(p)) - 1;  // the record of this header line
            strand_a[lk] = s == '+';
            link_b[lk] = n;
            strand_b[lk] = t == '+';
            lk++;
        }
    }
    // ---- bases ----
    if (seq) {
        const u32 acgt = eq_mask(c, 'A') | eq_mask(c, 'C') | eq_mask(c, 'G') | eq_mask(c, 'T');
        if (seq & ~acgt) atomicExch(err, 6);
        u64 codes = 0;  // 2-bit code of every byte of the chunk, byte i at bits [2i, 2i+2)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 x = (c.w[i] >> 1) & 0x03030303u;
            const u32 t = x | (x >> 6);
            codes |= (u64)((t | (t >> 12)) & 0xFFu) << (8 * i);
        }
        // squeeze out the non-sequence bytes run by run (a chunk holds one or two runs)
        u64 acc = 0;
        u32 cnt = 0;
        for (u32 m = seq; m;) {
            const u32 a = __ffs(m) - 1;
            const u32 inv = ~(m >> a);
            const u32 len = inv ? (u32)__ffs(inv) - 1 : 32u - a;
            const u64 run = (codes >> (2 * a)) & (len >= 32 ? ~0ull : (1ull << (2 * len)) - 1ull);
            acc |= run << (2 * cnt);
            cnt += len;
            m &= ~bit_range(a, a + len);
        }
        const u64 pos = s0;
        const u32 sh = 2 * (u32)(pos & 31);
        atomicOr(&seq_words[pos >> 5], (unsigned long long)(acc << sh));
        if (sh && cnt > 32 - (u32)(pos & 31)) atomicOr(&seq_words[(pos >> 5) + 1], (unsigned long long)(acc >> (64 - sh)));
    }
}

__global__ void __launch_bounds__(TB) weights_from_offsets(const u64* __restrict__ offsets, u64 U, u32 k, u64* __restrict__ w, int* __restrict__ err) {
    u64 u = (u64)blockIdx.x * TB + threadIdx.x;
    if (u >= U) return;
    const u64 len = offsets[u + 1] - offsets[u];
    if (len < k) {
        atomicExch(err, 5);
        w[u] = 1;
    } else {
        w[u] = len + 1 - k;
    }
}

}  // namespace

void build_graph_from_text(mtg_ctx* ctx, const char* text, u64 L, bool bcalm, u32 k, bool text_on_device) {
    MTG_REQUIRE(L == 0 || text, MTG_ERR_INVALID, "null text");
    // Text positions are 64-bit throughout; what is 32-bit are the chunk tags of the line-state scan (chunk index << 2) and
    // the per-chunk offsets (records, bases, links) -- the latter guarded by 64-bit totals below.
    MTG_REQUIRE(L < (u64(1) << 35) - 64, MTG_ERR_UNSUPPORTED, "text of 32 GiB or more: parse it in pieces with the host reader");
    cudaStream_t s = ctx->stream;
    if (!ctx->text_event_recorded) MTG_CUDA(cudaEventRecord(ctx->ev_build[0], s));
    ctx->text_event_recorded = false;
    auto& ws = ctx->parse_ws;
    DBuf<u32>& totals = ws.totals;
    const char* d_text = text;
    if (!text_on_device && L) {
        ws.text.resize(L, s);
        MTG_CUDA(cudaMemcpyAsync(ws.text.p, text, L, cudaMemcpyHostToDevice, s));
        d_text = ws.text.p;
    }
    DBuf<u8> strand_a, strand_b;
    DBuf<u64> link_a, link_b, weights;
    DBuf<int> err;
    err.resize(1, s);
    err.zero(s);
    totals.resize(4, s);
    totals.zero(s);
    u32 h_tot[3] = {0, 0, 0};
    const u64 n_chunks = (L + CHUNK - 1) / CHUNK;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_text) & 15u) == 0;
    if (L) {
        for (DBuf<u32>* b : {&ws.key, &ws.n_rec, &ws.n_seq, &ws.n_link, &ws.rbase, &ws.sbase, &ws.lbase}) b->resize(n_chunks, s);
        // ws.lbase doubles as the packed per-chunk counts until the scans have run
        MTG_LAUNCH(ctx, chunk_scan_text, grid_for(n_chunks, TB), TB, 0, d_text, L, n_chunks, aligned, (int)bcalm, ws.key.p, ws.lbase.p);
        inclusive_max_u32(ctx, ws.key.p, ws.key.p, n_chunks);
        MTG_LAUNCH(ctx, resolve_counts, grid_for(n_chunks, TB), TB, 0, ws.key.p, ws.lbase.p, n_chunks, ws.n_rec.p, ws.n_seq.p, ws.n_link.p);
        exclusive_sum_u32(ctx, ws.n_rec.p, ws.rbase.p, n_chunks, totals.p + 0);
        exclusive_sum_u32(ctx, ws.n_seq.p, ws.sbase.p, n_chunks, totals.p + 1);
        if (bcalm) exclusive_sum_u32(ctx, ws.n_link.p, ws.lbase.p, n_chunks, totals.p + 2);
        MTG_CUDA(cudaMemcpyAsync(h_tot, totals.p, sizeof(h_tot), cudaMemcpyDeviceToHost, s));
        unsigned long long h_tot64[3] = {0, 0, 0};
        const bool may_overflow = L >= 0xFFFFFFF0ull;  // below 4 GiB of text no count can reach 2^32
        DBuf<unsigned long long> tot64;
        if (may_overflow) {
            tot64.resize(3, s);
            tot64.zero(s);
            MTG_LAUNCH(ctx, sum_counts_u64, (unsigned)std::min<u64>(grid_for(n_chunks, TB).x, 148u * 16u), TB, 0, ws.n_rec.p, ws.n_seq.p,
                       bcalm ? ws.n_link.p : (const u32*)nullptr, n_chunks, tot64.p);
            MTG_CUDA(cudaMemcpyAsync(h_tot64, tot64.p, sizeof(h_tot64), cudaMemcpyDeviceToHost, s));
        }
        MTG_CUDA(cudaStreamSynchronize(s));
        if (may_overflow)
            MTG_REQUIRE(h_tot64[0] == h_tot[0] && h_tot64[1] == h_tot[1] && (!bcalm || h_tot64[2] == h_tot[2]), MTG_ERR_UNSUPPORTED,
                        "more than 2^32 - 1 records, bases or links in one file");
    }
    const u64 U = h_tot[0], B = h_tot[1], NL = bcalm ? h_tot[2] : 0;
    // the bases go straight into the context's 2-bit store (zeroed: the chunks OR their pieces in)
    const u64 nwords = (B + 31) / 32;
    ctx->seq_words.resize(nwords + 2, s);
    ctx->seq_words.zero(s);
    ctx->seq_off.resize(U + 1, s);
    ctx->total_bases = B;
    link_a.resize(NL, s);
    link_b.resize(NL, s);
    strand_a.resize(NL, s);
    strand_b.resize(NL, s);
    if (L)
        MTG_LAUNCH(ctx, chunk_scatter, grid_for(n_chunks, TB), TB, 0, d_text, L, n_chunks, aligned, ws.key.p, (int)bcalm, ws.rbase.p, ws.sbase.p,
                   ws.lbase.p, reinterpret_cast<unsigned long long*>(ctx->seq_words.p), ctx->seq_off.p, link_a.p, strand_a.p, link_b.p,
                   strand_b.p, err.p);
    MTG_CUDA(cudaMemcpyAsync(ctx->seq_off.p + U, &B, sizeof(u64), cudaMemcpyHostToDevice, s));
    if (bcalm) {
        weights.resize(U, s);
        if (U) MTG_LAUNCH(ctx, weights_from_offsets, grid_for(U, TB), TB, 0, ctx->seq_off.p, U, k, weights.p, err.p);
    }
    int h_err = 0;
    MTG_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    auto cleanup = [&] {
        link_a.release(s);
        link_b.release(s);
        strand_a.release(s);
        strand_b.release(s);
        weights.release(s);
        err.release(s);
    };
    if (h_err) {
        cleanup();
        switch (h_err) {
            case 1: throw Error{MTG_ERR_INPUT, "FASTA: expected '>'"};
            case 2: throw Error{MTG_ERR_INPUT, "bcalm: record id != position"};
            case 3: throw Error{MTG_ERR_INPUT, "bcalm: malformed L field"};
            case 4: throw Error{MTG_ERR_INPUT, "bcalm: malformed L sign"};
            case 6: throw Error{MTG_ERR_INPUT, "sequence contains a character other than A, C, G, T"};
            default: throw Error{MTG_ERR_INPUT, "sequence shorter than k"};
        }
    }
    MTG_CUDA(cudaEventRecord(ctx->ev_build[1], s));
    ctx->in_text_build = true;
    try {
        // sequences already packed in the context (seq_words / seq_off): the builders only use them
        if (bcalm) build_graph_from_links(ctx, U, weights.p, NL, link_a.p, strand_a.p, link_b.p, strand_b.p, k, nullptr, nullptr, true, B, true);
        else build_graph_from_sequences(ctx, nullptr, nullptr, U, k, true, B, true);
    } catch (...) {
        ctx->in_text_build = false;
        cleanup();
        throw;
    }
    ctx->in_text_build = false;
    cleanup();
}

}  // namespace mtg
