// parse.cu -- device-side FASTA / bcalm2 record parser.
//
// The reference parses records one after the other on the host (genome-graph readers, call sites
// src/bin.rs:896-899, :907-910).  Here the whole file is classified in parallel, one thread per byte:
//   1. line starts -> inclusive max-scan = start of the line every byte belongs to;
//   2. a byte is header text iff its line starts with '>'; everything else except line ends is sequence;
//   3. exclusive scans over the three flag arrays (record starts, sequence bytes, `L:` fields) give every
//      record its index, every base its position in the concatenated sequence and every link its slot;
//   4. one scatter pass writes the bases, the record offsets and the parsed links in file order.
// Outputs stay on the device and feed build_graph_from_sequences / build_graph_from_links directly.
#include <algorithm>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;

__global__ void __launch_bounds__(TB) mark_line_starts(const char* __restrict__ text, u64 L, u32* __restrict__ ls) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < L) ls[i] = (i == 0 || text[i - 1] == '\n') ? (u32)(i + 1) : 0u;
}

__device__ __forceinline__ bool is_blank(char c) { return c == ' ' || c == '\t'; }
__device__ __forceinline__ bool is_eol(char c) { return c == '\n' || c == '\r'; }

__global__ void __launch_bounds__(TB)
    classify_bytes(const char* __restrict__ text, u64 L, const u32* __restrict__ ls, int bcalm, u8* __restrict__ rec_flag,
                   u8* __restrict__ seq_flag, u8* __restrict__ link_flag) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= L) return;
    const char c = text[i];
    const u64 start = ls[i] - 1;
    const bool hdr = text[start] == '>';
    rec_flag[i] = (hdr && start == i) ? 1 : 0;
    seq_flag[i] = (!hdr && !is_eol(c)) ? 1 : 0;
    u8 lf = 0;
    if (bcalm && hdr && c == 'L' && i > start && is_blank(text[i - 1]) && i + 1 < L && text[i + 1] == ':') {
        u64 q = i;  // token end: the host reader only treats tokens of at least 7 characters as links
        while (q < L && !is_blank(text[q]) && !is_eol(text[q])) q++;
        lf = (q - i >= 7) ? 1 : 0;
    }
    link_flag[i] = lf;
}

__global__ void __launch_bounds__(TB)
    scatter_records(const char* __restrict__ text, u64 L, const u32* __restrict__ ls, int bcalm, const u8* __restrict__ rec_flag,
                    const u8* __restrict__ seq_flag, const u8* __restrict__ link_flag, const u32* __restrict__ rscan,
                    const u32* __restrict__ sscan, const u32* __restrict__ lscan, char* __restrict__ seq, u64* __restrict__ offsets,
                    u64* __restrict__ link_a, u8* __restrict__ strand_a, u64* __restrict__ link_b, u8* __restrict__ strand_b,
                    int* __restrict__ err) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= L) return;
    if (seq_flag[i]) {
        if (rscan[i] == 0) atomicExch(err, 1);  // sequence data before the first header
        seq[sscan[i]] = text[i];
    }
    if (rec_flag[i]) {
        const u32 r = rscan[i];
        offsets[r] = sscan[i];
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
    }
    if (link_flag[i]) {  // L:<+/->:<id>:<+/->
        const u32 slot = lscan[i];
        const char s = text[i + 2];
        u64 c = i + 4, j = 0;
        bool digits = false;
        while (c < L && text[c] >= '0' && text[c] <= '9') {
            j = j * 10 + (u64)(text[c] - '0');
            c++;
            digits = true;
        }
        const bool shape = text[i + 3] == ':' && digits && c + 1 < L && text[c] == ':' && !is_blank(text[c + 1]) && !is_eol(text[c + 1]);
        const char t = shape ? text[c + 1] : '?';
        if (!shape) atomicExch(err, 3);
        else if ((s != '+' && s != '-') || (t != '+' && t != '-')) atomicExch(err, 4);
        link_a[slot] = rscan[ls[i] - 1];
        strand_a[slot] = s == '+';
        link_b[slot] = j;
        strand_b[slot] = t == '+';
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
    MTG_REQUIRE(L < 0xFFFFFFF0ull, MTG_ERR_UNSUPPORTED, "text of 4 GiB or more: parse it in pieces with the host reader");
    cudaStream_t s = ctx->stream;
    auto& ws = ctx->parse_ws;
    DBuf<u32>&ls = ws.ls, &rscan = ws.rscan, &sscan = ws.sscan, &lscan = ws.lscan, &totals = ws.totals;
    DBuf<u8>&rec_flag = ws.rec_flag, &seq_flag = ws.seq_flag, &link_flag = ws.link_flag;
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
    if (L) {
        ls.resize(L, s);
        for (DBuf<u32>* b : {&rscan, &sscan, &lscan}) b->resize(L, s);
        for (DBuf<u8>* b : {&rec_flag, &seq_flag, &link_flag}) b->resize(L, s);
        MTG_LAUNCH(ctx, mark_line_starts, grid_for(L, TB), TB, 0, d_text, L, ls.p);
        inclusive_max_u32(ctx, ls.p, ls.p, L);
        MTG_LAUNCH(ctx, classify_bytes, grid_for(L, TB), TB, 0, d_text, L, ls.p, (int)bcalm, rec_flag.p, seq_flag.p, link_flag.p);
        exclusive_sum_u8(ctx, rec_flag.p, rscan.p, L, totals.p + 0);
        exclusive_sum_u8(ctx, seq_flag.p, sscan.p, L, totals.p + 1);
        if (bcalm) exclusive_sum_u8(ctx, link_flag.p, lscan.p, L, totals.p + 2);
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
        MTG_LAUNCH(ctx, scatter_records, grid_for(L, TB), TB, 0, d_text, L, ls.p, (int)bcalm, rec_flag.p, seq_flag.p, link_flag.p, rscan.p,
                   sscan.p, lscan.p, seq.p, offsets.p, link_a.p, strand_a.p, link_b.p, strand_b.p, err.p);
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
    try {
        if (bcalm) build_graph_from_links(ctx, U, weights.p, NL, link_a.p, strand_a.p, link_b.p, strand_b.p, k, seq.p, offsets.p, true, B);
        else build_graph_from_sequences(ctx, seq.p, offsets.p, U, k, true, B);
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace mtg
