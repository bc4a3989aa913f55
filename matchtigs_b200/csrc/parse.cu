// parse.cu -- device-side FASTA / bcalm2 record parser.
//
// The reference parses records one after the other on the host (genome-graph readers, call sites
// src/bin.rs:896-899, :907-910).  Here one thread owns 32 bytes of the file:
//   1. per chunk the state of its last line start (header line or not) -> inclusive max-scan over the chunks = the
//      state every chunk starts in (a two-state automaton, composed by "the latest line start wins");
//   2. per chunk the number of record starts, sequence bytes and `L:` fields -> three exclusive sums over the chunks;
//   3. one pass writes the bases, the record offsets and the parsed links in file order, every thread counting on
//      from its chunk's bases.
// The text is read three times and the work arrays are ~1 byte per text byte (a per-byte formulation needs ~19).
// Outputs stay on the device and feed build_graph_from_sequences / build_graph_from_links directly.
#include <algorithm>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;
constexpr int CHUNK = 32;  // text bytes per thread

__device__ __forceinline__ bool is_blank(char c) { return c == ' ' || c == '\t'; }
__device__ __forceinline__ bool is_eol(char c) { return c == '\n' || c == '\r'; }

// The chunk's bytes in registers: two 16-byte loads where the chunk is complete and the text 16-byte aligned.
struct ChunkBytes {
    char b[CHUNK];
    int n;
};
__device__ __forceinline__ ChunkBytes load_chunk(const char* __restrict__ text, u64 L, u64 base, bool aligned) {
    ChunkBytes c;
    c.n = (int)min((u64)CHUNK, L - base);
    if (aligned && c.n == CHUNK) {
        const uint4* p = reinterpret_cast<const uint4*>(text + base);
        *reinterpret_cast<uint4*>(c.b) = p[0];
        *reinterpret_cast<uint4*>(c.b + 16) = p[1];
    } else {
        for (int i = 0; i < CHUNK; i++) c.b[i] = i < c.n ? text[base + i] : '\n';
    }
    return c;
}

// A line is a header line iff it starts with '>'.  Whether the bytes at the start of a chunk belong to a header line
// is decided by the last line start before the chunk: per chunk the code of its LAST line start (0 none, 1 sequence
// line, 2 header line), tagged with the chunk index so that an inclusive max-scan carries the latest one forward.
__global__ void __launch_bounds__(TB) chunk_line_state(const char* __restrict__ text, u64 L, u64 n_chunks, bool aligned, u32* __restrict__ key) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    const u64 base = j * CHUNK;
    const ChunkBytes c = load_chunk(text, L, base, aligned);
    char prev = base ? text[base - 1] : '\n';
    u32 code = 0;
#pragma unroll
    for (int i = 0; i < CHUNK; i++) {
        if (i < c.n) {
            if (prev == '\n') code = c.b[i] == '>' ? 2u : 1u;
            prev = c.b[i];
        }
    }
    key[j] = code ? (u32)((j + 1) << 2) | code : 0u;
}

// Visits the bytes of chunk j in order with their classification (same rules as the host reader, csrc/reader.cpp):
//   rec  -- '>' at a line start;  seq -- any byte of a non-header line except line ends;
//   link -- (bcalm) an `L:` token of at least 7 characters in a header line, preceded by a blank.
template <class F>
__device__ __forceinline__ void visit_chunk(const char* __restrict__ text, u64 L, u64 j, bool aligned, const u32* __restrict__ key_scan,
                                            int bcalm, F&& f) {
    const u64 base = j * CHUNK;
    const ChunkBytes c = load_chunk(text, L, base, aligned);
    bool hdr = j ? (key_scan[j - 1] & 3u) == 2u : false;
    char prev = base ? text[base - 1] : '\n';
    for (int i = 0; i < c.n; i++) {
        const char ch = c.b[i];
        const u64 pos = base + i;
        const bool line_start = prev == '\n';
        if (line_start) hdr = ch == '>';
        bool link = false;
        if (bcalm && hdr && ch == 'L' && is_blank(prev) && pos + 6 < L) {
            const char nx = i + 1 < c.n ? c.b[i + 1] : text[pos + 1];
            if (nx == ':') {
                link = true;  // token end: the host reader only treats tokens of at least 7 characters as links
                for (int q = 2; q < 7; q++) {
                    const char t = text[pos + q];
                    if (is_blank(t) || is_eol(t)) link = false;
                }
            }
        }
        f(pos, ch, line_start && hdr, !hdr && !is_eol(ch), link);
        prev = ch;
    }
}

__global__ void __launch_bounds__(TB)
    chunk_counts(const char* __restrict__ text, u64 L, u64 n_chunks, bool aligned, const u32* __restrict__ key_scan, int bcalm,
                 u32* __restrict__ n_rec, u32* __restrict__ n_seq, u32* __restrict__ n_link) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    u32 r = 0, sq = 0, lk = 0;
    visit_chunk(text, L, j, aligned, key_scan, bcalm, [&](u64, char, bool rec, bool seq, bool link) {
        r += rec;
        sq += seq;
        lk += link;
    });
    n_rec[j] = r;
    n_seq[j] = sq;
    n_link[j] = lk;
}

__global__ void __launch_bounds__(TB)
    chunk_scatter(const char* __restrict__ text, u64 L, u64 n_chunks, bool aligned, const u32* __restrict__ key_scan, int bcalm,
                  const u32* __restrict__ rbase, const u32* __restrict__ sbase, const u32* __restrict__ lbase, char* __restrict__ seq,
                  u64* __restrict__ offsets, u64* __restrict__ link_a, u8* __restrict__ strand_a, u64* __restrict__ link_b,
                  u8* __restrict__ strand_b, int* __restrict__ err) {
    const u64 j = (u64)blockIdx.x * TB + threadIdx.x;
    if (j >= n_chunks) return;
    u32 r = rbase[j], sq = sbase[j], lk = lbase[j];  // records / sequence bytes / links before the current byte
    visit_chunk(text, L, j, aligned, key_scan, bcalm, [&](u64 i, char ch, bool rec, bool is_seq, bool link) {
        if (is_seq) {
            if (r == 0) atomicExch(err, 1);  // sequence data before the first header
            seq[sq++] = ch;
        }
        if (rec) {
            offsets[r] = sq;
            if (bcalm) {  // the record id must equal its position
                u64 p = i + 1, id = 0;
                bool any = false;
                while (p < L && text[p] >= '0' && text[p] <= '9') {
                    id = id * 10 + (u64)(text[p] - '0');
                    p++;
                    any = true;
                }
                if (!any || id != r) atomicExch(err, 2);
            }
            r++;
        }
        if (link) {  // L:<+/->:<id>:<+/->
            const char s = text[i + 2];
            u64 c = i + 4, n = 0;
            bool digits = false;
            while (c < L && text[c] >= '0' && text[c] <= '9') {
                n = n * 10 + (u64)(text[c] - '0');
                c++;
                digits = true;
            }
            const bool shape = text[i + 3] == ':' && digits && c + 1 < L && text[c] == ':' && !is_blank(text[c + 1]) && !is_eol(text[c + 1]);
            const char t = shape ? text[c + 1] : '?';
            if (!shape) atomicExch(err, 3);
            else if ((s != '+' && s != '-') || (t != '+' && t != '-')) atomicExch(err, 4);
            link_a[lk] = r - 1;  // the record of this header line
            strand_a[lk] = s == '+';
            link_b[lk] = n;
            strand_b[lk] = t == '+';
            lk++;
        }
    });
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
    MTG_REQUIRE(L < 0xFFFFFFF0ull, MTG_ERR_UNSUPPORTED, "text of 4 GiB or more: parse it in pieces with the host reader");
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
    DBuf<char> seq;
    DBuf<u64> offsets, link_a, link_b, weights;
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
        MTG_LAUNCH(ctx, chunk_line_state, grid_for(n_chunks, TB), TB, 0, d_text, L, n_chunks, aligned, ws.key.p);
        inclusive_max_u32(ctx, ws.key.p, ws.key.p, n_chunks);
        MTG_LAUNCH(ctx, chunk_counts, grid_for(n_chunks, TB), TB, 0, d_text, L, n_chunks, aligned, ws.key.p, (int)bcalm, ws.n_rec.p, ws.n_seq.p,
                   ws.n_link.p);
        exclusive_sum_u32(ctx, ws.n_rec.p, ws.rbase.p, n_chunks, totals.p + 0);
        exclusive_sum_u32(ctx, ws.n_seq.p, ws.sbase.p, n_chunks, totals.p + 1);
        if (bcalm) exclusive_sum_u32(ctx, ws.n_link.p, ws.lbase.p, n_chunks, totals.p + 2);
        MTG_CUDA(cudaMemcpyAsync(h_tot, totals.p, sizeof(h_tot), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
    }
    const u64 U = h_tot[0], B = h_tot[1], NL = bcalm ? h_tot[2] : 0;
    seq.resize(B + 32, s);
    offsets.resize(U + 1, s);
    link_a.resize(NL, s);
    link_b.resize(NL, s);
    strand_a.resize(NL, s);
    strand_b.resize(NL, s);
    if (L)
        MTG_LAUNCH(ctx, chunk_scatter, grid_for(n_chunks, TB), TB, 0, d_text, L, n_chunks, aligned, ws.key.p, (int)bcalm, ws.rbase.p, ws.sbase.p,
                   ws.lbase.p, seq.p, offsets.p, link_a.p, strand_a.p, link_b.p, strand_b.p, err.p);
    MTG_CUDA(cudaMemcpyAsync(offsets.p + U, &B, sizeof(u64), cudaMemcpyHostToDevice, s));
    if (bcalm) {
        weights.resize(U, s);
        if (U) MTG_LAUNCH(ctx, weights_from_offsets, grid_for(U, TB), TB, 0, offsets.p, U, k, weights.p, err.p);
    }
    int h_err = 0;
    MTG_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    auto cleanup = [&] {
        seq.release(s);
        offsets.release(s);
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
            default: throw Error{MTG_ERR_INPUT, "sequence shorter than k"};
        }
    }
    MTG_CUDA(cudaEventRecord(ctx->ev_build[1], s));
    ctx->in_text_build = true;
    try {
        if (bcalm) build_graph_from_links(ctx, U, weights.p, NL, link_a.p, strand_a.p, link_b.p, strand_b.p, k, seq.p, offsets.p, true, B);
        else build_graph_from_sequences(ctx, seq.p, offsets.p, U, k, true, B);
    } catch (...) {
        ctx->in_text_build = false;
        cleanup();
        throw;
    }
    ctx->in_text_build = false;
    cleanup();
}

}  // namespace mtg
