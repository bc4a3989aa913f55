// graph.cu -- step 1 of the hot path: the bidirected, edge-centric unitig overlap graph.
//
//   K1  pack_sequences      ASCII -> 2-bit store (32 bases per 64-bit word), validation
//       extract_end_keys    canonical (k-1)-mer key of both ends of every unitig + orientation
//   K2  radix_sort_pairs    (prims.cu) stable sort of (key, lookup position)
//   K3  mark_groups / assign_nodes / make_edges   first-seen node numbering identical to the
//                           reference reader (assumption P5), mirror table
//       count_degrees / classify_nodes / short-edge CSR / sources / target bitmap
//
// and the link-driven variant (src/clib.rs:135-259): lock-free hooking finds the components of
// the 4U end slots, then one thread per component replays its unions in call order with
// union-by-rank so that representatives -- and therefore node ids -- equal the reference's.
#include <algorithm>
#include <cstdlib>
#include <memory>

#include "mtg_internal.cuh"

namespace mtg {

namespace {

constexpr int TB = 256;

// ---------------- K1: pack ----------------
// code = (c >> 1) & 3 maps A,C,T,G -> 0,1,2,3; complement is code ^ 2.
__device__ __forceinline__ u32 base_code(u32 c, u32* bad) {
    u32 code = (c >> 1) & 3u;
    const u32 back = 0x47544341u;  // 'A','C','T','G' little endian
    if (((back >> (8 * code)) & 0xFFu) != c) *bad = 1;
    return code;
}

__global__ void __launch_bounds__(TB) pack_sequences(const char* __restrict__ seq, u64* __restrict__ words, u64 total, u64 nwords,
                                                     int* __restrict__ err) {
    u64 w = (u64)blockIdx.x * TB + threadIdx.x;
    if (w >= nwords) return;
    u64 base = w * 32;
    u64 out = 0;
    u32 bad = 0;
    if (base + 32 <= total) {
        const uint4* p = reinterpret_cast<const uint4*>(seq + base);
        uint4 a = p[0], b = p[1];
        u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                u32 ch = (c[q] >> (8 * j)) & 0xFFu;
                out |= (u64)base_code(ch, &bad) << (2 * (q * 4 + j));
            }
        }
    } else {
        for (u64 j = 0; base + j < total; j++) out |= (u64)base_code((u8)seq[base + j], &bad) << (2 * j);
    }
    words[w] = out;
    if (bad) atomicExch(err, 1);
}

__device__ __forceinline__ u64 read_bases64(const u64* __restrict__ words, u64 pos) {
    u64 w = pos >> 5;
    u32 s = (u32)(pos & 31) * 2;
    u64 x = words[w] >> s;
    if (s) x |= words[w + 1] << (64 - s);
    return x;
}
__device__ __forceinline__ u64 low_bases_mask(u32 nbases) { return nbases >= 32 ? ~0ull : ((1ull << (2 * nbases)) - 1ull); }
__device__ __forceinline__ u64 rev2_64(u64 x) {
    u64 r = __brevll(x);
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}
// reverse complement of an L-base string held little-endian in (lo, hi)
__device__ __forceinline__ void revcomp_key(u64 lo, u64 hi, u32 L, u64* rlo, u64* rhi) {
    u64 nlo = rev2_64(hi), nhi = rev2_64(lo);  // base j -> position 63 - j
    u32 sh = 2 * (64 - L);                    // bring base L-1 down to position 0
    u64 a, b;
    if (sh == 0) {
        a = nlo;
        b = nhi;
    } else if (sh < 64) {
        a = (nlo >> sh) | (nhi << (64 - sh));
        b = nhi >> sh;
    } else {
        a = nhi >> (sh - 64);
        b = 0;
    }
    a ^= 0xAAAAAAAAAAAAAAAAull & low_bases_mask(L);
    b ^= L > 32 ? (0xAAAAAAAAAAAAAAAAull & low_bases_mask(L - 32)) : 0ull;
    *rlo = a;
    *rhi = b;
}

// One thread per lookup position p = 2u (prefix) / 2u+1 (suffix), the order in which the reference
// reader consults its (k-1)-mer map.
__global__ void __launch_bounds__(TB)
    extract_end_keys(const u64* __restrict__ words, const u64* __restrict__ off, u64 U, u32 k, u64* __restrict__ key_lo,
                     u64* __restrict__ key_hi, u32* __restrict__ val, u32* __restrict__ unitig_w, int* __restrict__ err) {
    u64 p = (u64)blockIdx.x * TB + threadIdx.x;
    if (p >= 2 * U) return;
    u64 u = p >> 1;
    u64 s = off[u], e = off[u + 1];
    u64 len = e - s;
    const u32 L = k - 1;
    if (len < k) {
        atomicExch(err, 2);
        key_lo[p] = 0;
        if (key_hi) key_hi[p] = 0;
        val[p] = (u32)p;
        if (!(p & 1)) unitig_w[u] = 0;
        return;
    }
    if (!(p & 1)) {
        u64 w = len + 1 - k;  // compute_edge_weights, src/bin.rs:369
        if (w > 0x7FFFFFFFull) atomicExch(err, 3);
        unitig_w[u] = (u32)w;
    }
    u64 start = (p & 1) ? e - L : s;
    u64 lo = read_bases64(words, start) & low_bases_mask(L);
    u64 hi = L > 32 ? (read_bases64(words, start + 32) & low_bases_mask(L - 32)) : 0ull;
    u64 rlo, rhi;
    revcomp_key(lo, hi, L, &rlo, &rhi);
    bool flip = (rhi < hi) || (rhi == hi && rlo < lo);
    key_lo[p] = flip ? rlo : lo;
    if (key_hi) key_hi[p] = flip ? rhi : hi;
    val[p] = (u32)p | (flip ? 0x80000000u : 0u);
}

// ---------------- K3: numbering ----------------
__global__ void __launch_bounds__(TB)
    mark_groups(const u64* __restrict__ key_lo, const u64* __restrict__ key_hi, const u32* __restrict__ val, u64 n, u32 L,
                u32* __restrict__ head_idx, u32* __restrict__ created) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    u64 lo = key_lo[i], hi = key_hi ? key_hi[i] : 0ull;
    bool head = (i == 0) || key_lo[i - 1] != lo || (key_hi && key_hi[i - 1] != hi);
    head_idx[i] = head ? (u32)i : 0u;
    if (head) {
        u64 rlo, rhi;
        revcomp_key(lo, hi, L, &rlo, &rhi);
        bool pal = (rlo == lo) && (rhi == hi);
        created[val[i] & 0x7FFFFFFFu] = pal ? 1u : 2u;  // nodes the reference creates at this lookup
    }
}

__global__ void __launch_bounds__(TB)
    assign_nodes(const u32* __restrict__ val, const u32* __restrict__ head_idx, const u32* __restrict__ created,
                 const u32* __restrict__ base, u64 n, u32* __restrict__ node_ep, u32* __restrict__ mirror_ep) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    u32 v = val[i], vh = val[head_idx[i]];
    u32 p = v & 0x7FFFFFFFu, p0 = vh & 0x7FFFFFFFu;
    bool same = (v >> 31) == (vh >> 31);
    bool pal = created[p0] == 1u;
    u32 b = base[p0];
    u32 node = b + ((same || pal) ? 0u : 1u);
    u32 mir = pal ? b : b + (same ? 1u : 0u);
    node_ep[p] = node;
    mirror_ep[p] = mir;
}

__global__ void __launch_bounds__(TB) make_edges(const u32* __restrict__ node_ep, const u32* __restrict__ mirror_ep, u64 U,
                                                 u32* __restrict__ edge_from, u32* __restrict__ edge_to, u32* __restrict__ mirror) {
    u64 u = (u64)blockIdx.x * TB + threadIdx.x;
    if (u >= U) return;
    u32 a = node_ep[2 * u], b = node_ep[2 * u + 1], am = mirror_ep[2 * u], bm = mirror_ep[2 * u + 1];
    edge_from[2 * u] = a;
    edge_to[2 * u] = b;
    edge_from[2 * u + 1] = bm;
    edge_to[2 * u + 1] = am;
    mirror[a] = am;
    mirror[am] = a;
    mirror[b] = bm;
    mirror[bm] = b;
}

// ---------------- degrees, imbalance, sources, targets, short-edge CSR ----------------
__global__ void __launch_bounds__(TB) count_degrees(const u32* __restrict__ edge_from, const u32* __restrict__ unitig_w, u64 E, u32 k,
                                                    u32* __restrict__ out_deg, u32* __restrict__ deg_s, u32* __restrict__ short_flag) {
    u64 e = (u64)blockIdx.x * TB + threadIdx.x;
    if (e >= E) return;
    u32 f = edge_from[e];
    atomicAdd(&out_deg[f], 1u);
    bool is_short = unitig_w[e >> 1] <= k - 1;
    short_flag[e] = is_short ? 1u : 0u;
    if (is_short) atomicAdd(&deg_s[f], 1u);
}

// compute_eulerian_superfluous_out_biedges + the scan of greedytigs/mod.rs:229-245.
// in_degree(v) == out_degree(mirror(v)) by the mirror property of the bigraph.
__global__ void __launch_bounds__(TB)
    classify_nodes(const u32* __restrict__ out_deg, const u32* __restrict__ mirror, u64 N, bool self_mirror_zero, i32* __restrict__ imbalance,
                   u32* __restrict__ src_flag, u32* __restrict__ target_bits, unsigned long long* __restrict__ counters) {
    u64 v = (u64)blockIdx.x * TB + threadIdx.x;
    bool is_src = false, is_tgt = false, self_unb = false;
    i32 diff = 0;
    if (v < N) {
        u32 m = mirror[v];
        if (m == (u32)v) {
            diff = self_mirror_zero ? 0 : (i32)(out_deg[v] & 1u);  // assumption P2
            is_src = is_tgt = self_unb = diff != 0;
        } else {
            diff = (i32)out_deg[v] - (i32)out_deg[m];
            is_src = diff < 0;
            is_tgt = diff > 0;
        }
        imbalance[v] = diff;
        src_flag[v] = is_src ? 1u : 0u;
    }
    unsigned tb = __ballot_sync(0xffffffffu, is_tgt);
    unsigned sb = __ballot_sync(0xffffffffu, self_unb);
    if ((threadIdx.x & 31) == 0 && (v & ~31ull) < N) {
        target_bits[v >> 5] = tb;
        if (tb) atomicAdd(&counters[0], (unsigned long long)__popc(tb));
        if (sb) atomicAdd(&counters[1], (unsigned long long)__popc(sb));
    }
    // total target multiplicity: an upper bound for the number of matches (sizes the triple buffers without a round trip)
    const unsigned tm = __reduce_add_sync(0xffffffffu, is_tgt ? (unsigned)diff : 0u);
    if ((threadIdx.x & 31) == 0 && tm) atomicAdd(&counters[2], (unsigned long long)tm);
}

__global__ void __launch_bounds__(TB) compact_flagged(const u32* __restrict__ flag, const u32* __restrict__ pos, u64 n, u32* __restrict__ out) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = (u32)i;
}

__global__ void __launch_bounds__(TB) gather_short_edges(const u32* __restrict__ flag, const u32* __restrict__ pos,
                                                         const u32* __restrict__ edge_from, u64 E, u32* __restrict__ from_s,
                                                         u32* __restrict__ eid_s) {
    u64 e = (u64)blockIdx.x * TB + threadIdx.x;
    if (e < E && flag[e]) {
        from_s[pos[e]] = edge_from[e];
        eid_s[pos[e]] = (u32)e;
    }
}

__global__ void __launch_bounds__(TB) fill_short_csr(const u32* __restrict__ eid_sorted, const u32* __restrict__ edge_to,
                                                     const u32* __restrict__ unitig_w, u64 Es, u32* __restrict__ col_s, u8* __restrict__ w_s) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= Es) return;
    u32 e = eid_sorted[i];
    col_s[i] = edge_to[e];
    w_s[i] = (u8)unitig_w[e >> 1];
}

// ---------------- links path ----------------
__device__ __forceinline__ u32 slot_fwd_in(u32 u) { return u * 4; }       // src/clib.rs:104-122
__device__ __forceinline__ u32 slot_fwd_out(u32 u) { return u * 4 + 2; }
__device__ __forceinline__ u32 slot_bwd_in(u32 u) { return u * 4 + 3; }
__device__ __forceinline__ u32 slot_bwd_out(u32 u) { return u * 4 + 1; }

// union op 2i / 2i+1 of link i, in the order matchtigs_merge_nodes issues them (src/clib.rs:168-169)
__device__ __forceinline__ void link_op(const u64* a, const u8* sa, const u64* b, const u8* sb, u64 op, u32* x, u32* y) {
    u64 i = op >> 1;
    u32 ua = (u32)a[i], ub = (u32)b[i];
    bool fa = sa[i] != 0, fb = sb[i] != 0;
    if (!(op & 1)) {
        *x = fa ? slot_fwd_out(ua) : slot_bwd_out(ua);  // out_a
        *y = fb ? slot_fwd_in(ub) : slot_bwd_in(ub);    // in_b
    } else {
        *x = fa ? slot_bwd_in(ua) : slot_fwd_in(ua);    // mirror_in_a
        *y = fb ? slot_bwd_out(ub) : slot_fwd_out(ub);  // mirror_out_b
    }
}

__global__ void __launch_bounds__(TB) iota_u32(u32* __restrict__ p, u64 n) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < n) p[i] = (u32)i;
}

__device__ __forceinline__ u32 cc_find(volatile u32* cc, u32 x) {
    u32 p = cc[x];
    while (p != x) {
        u32 g = cc[p];
        cc[x] = g;  // path halving; racy but monotone (labels only decrease towards the root)
        x = p;
        p = g;
    }
    return x;
}

__global__ void __launch_bounds__(TB) cc_hook(const u64* __restrict__ a, const u8* __restrict__ sa, const u64* __restrict__ b,
                                              const u8* __restrict__ sb, u64 nops, u64 U, u32* cc, int* __restrict__ err) {
    u64 op = (u64)blockIdx.x * TB + threadIdx.x;
    if (op >= nops) return;
    if (a[op >> 1] >= U || b[op >> 1] >= U) {
        atomicExch(err, 4);
        return;
    }
    u32 x, y;
    link_op(a, sa, b, sb, op, &x, &y);
    for (;;) {
        x = cc_find(cc, x);
        y = cc_find(cc, y);
        if (x == y) break;
        if (x < y) {
            u32 t = x;
            x = y;
            y = t;
        }
        if (atomicCAS(&cc[x], x, y) == x) break;  // hook the larger root below the smaller
    }
}

__global__ void __launch_bounds__(TB) cc_flatten(u32* cc, u64 n) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i < n) {
        u32 r = cc_find(cc, (u32)i);
        cc[i] = r;
    }
}

__global__ void __launch_bounds__(TB)
    op_component_keys(const u64* __restrict__ a, const u8* __restrict__ sa, const u64* __restrict__ b, const u8* __restrict__ sb,
                      u64 nops, const u32* __restrict__ cc, u32* __restrict__ key, u32* __restrict__ val) {
    u64 op = (u64)blockIdx.x * TB + threadIdx.x;
    if (op >= nops) return;
    u32 x, y;
    link_op(a, sa, b, sb, op, &x, &y);
    key[op] = cc[x];
    val[op] = (u32)op;
}

// One thread per component: replay the component's unions in call order with the reference's
// union-by-rank rule (disjoint-sets 0.4.2, assumption P7): equal ranks attach the first root below the second.
__global__ void __launch_bounds__(TB)
    replay_unions(const u32* __restrict__ key_sorted, const u32* __restrict__ op_sorted, u64 nops, const u64* __restrict__ a,
                  const u8* __restrict__ sa, const u64* __restrict__ b, const u8* __restrict__ sb, bool first_root_wins, u32* parent, u8* rank) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= nops) return;
    u32 comp = key_sorted[i];
    if (i > 0 && key_sorted[i - 1] == comp) return;  // not the first op of its component
    for (u64 j = i; j < nops && key_sorted[j] == comp; j++) {
        u32 x, y;
        link_op(a, sa, b, sb, op_sorted[j], &x, &y);
        while (parent[x] != x) {
            parent[x] = parent[parent[x]];
            x = parent[x];
        }
        while (parent[y] != y) {
            parent[y] = parent[parent[y]];
            y = parent[y];
        }
        if (x == y) continue;
        u8 rx = rank[x], ry = rank[y];
        if (rx > ry) parent[y] = x;
        else if (ry > rx) parent[x] = y;
        else if (first_root_wins) {  // assumption P7 flipped
            parent[y] = x;
            rank[x] = rx + 1;
        } else {
            parent[x] = y;
            rank[y] = ry + 1;
        }
    }
}

__global__ void __launch_bounds__(TB) find_representatives(const u32* __restrict__ parent, u64 n, u32* __restrict__ rep, u32* __restrict__ is_rep) {
    u64 i = (u64)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    u32 x = (u32)i;
    while (parent[x] != x) x = parent[x];
    rep[i] = x;
    is_rep[i] = (x == (u32)i) ? 1u : 0u;
}

// matchtigs_build_graph, src/clib.rs:202-249
__global__ void __launch_bounds__(TB)
    make_edges_from_slots(const u32* __restrict__ rep, const u32* __restrict__ node_of_rep, u64 U, u32* __restrict__ edge_from,
                          u32* __restrict__ edge_to, u32* __restrict__ mirror) {
    u64 u = (u64)blockIdx.x * TB + threadIdx.x;
    if (u >= U) return;
    u32 n1 = node_of_rep[rep[slot_fwd_in((u32)u)]], n2 = node_of_rep[rep[slot_fwd_out((u32)u)]];
    u32 mn2 = node_of_rep[rep[slot_bwd_in((u32)u)]], mn1 = node_of_rep[rep[slot_bwd_out((u32)u)]];
    edge_from[2 * u] = n1;
    edge_to[2 * u] = n2;
    edge_from[2 * u + 1] = mn2;
    edge_to[2 * u + 1] = mn1;
    mirror[n1] = mn1;
    mirror[mn1] = n1;
    mirror[n2] = mn2;
    mirror[mn2] = n2;
}

// verify_node_pairing (src/clib.rs:251): every unitig must see a consistent involution.
__global__ void __launch_bounds__(TB) verify_pairing(const u32* __restrict__ edge_from, const u32* __restrict__ edge_to,
                                                     const u32* __restrict__ mirror, u64 U, int* __restrict__ err) {
    u64 u = (u64)blockIdx.x * TB + threadIdx.x;
    if (u >= U) return;
    u32 n1 = edge_from[2 * u], n2 = edge_to[2 * u], mn2 = edge_from[2 * u + 1], mn1 = edge_to[2 * u + 1];
    if (mirror[n1] != mn1 || mirror[mn1] != n1 || mirror[n2] != mn2 || mirror[mn2] != n2) atomicExch(err, 5);
}

__global__ void __launch_bounds__(TB) weights_to_u32(const u64* __restrict__ w, u64 U, u32* __restrict__ out, int* __restrict__ err) {
    u64 u = (u64)blockIdx.x * TB + threadIdx.x;
    if (u >= U) return;
    if (w[u] == 0 || w[u] > 0x7FFFFFFFull) atomicExch(err, 3);
    out[u] = (u32)w[u];
}

int bits_for(u64 n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) b++;
    return b;
}

void throw_on_err_flag(int h) {
    switch (h) {
        case 0: return;
        case 1: throw Error{MTG_ERR_INPUT, "sequence contains a character other than A, C, G, T"};
        case 2: throw Error{MTG_ERR_INPUT, "sequence shorter than k"};
        case 3: throw Error{MTG_ERR_INPUT, "unitig weight out of range"};
        case 4: throw Error{MTG_ERR_INPUT, "link references unknown unitig"};
        case 5: throw Error{MTG_ERR_INPUT, "node pairing violated (inconsistent links)"};
        default: throw Error{MTG_ERR_INTERNAL, "device error flag " + std::to_string(h)};
    }
}

void check_err_flag(mtg_ctx* ctx, int* d_err) {
    int h = 0;
    MTG_CUDA(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    throw_on_err_flag(h);
}

// Packs the sequences into ctx->seq_words / seq_off.  Returns device pointers to the ASCII (for nothing else).
void upload_and_pack(mtg_ctx* ctx, const char* seq, const u64* offsets, u64 U, bool on_device, u64 total_known, int* d_err) {
    cudaStream_t s = ctx->stream;
    u64 total = total_known;
    if (on_device) {
        if (total == UNKNOWN_TOTAL) {  // device-resident offsets from outside: the last one has to come back
            MTG_CUDA(cudaMemcpyAsync(&total, offsets + U, sizeof(u64), cudaMemcpyDeviceToHost, s));
            MTG_CUDA(cudaStreamSynchronize(s));
        }
        ctx->seq_off.resize(U + 1, s);
        MTG_CUDA(cudaMemcpyAsync(ctx->seq_off.p, offsets, (U + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    } else {
        total = offsets[U];
        ctx->seq_off.upload(offsets, U + 1, s);
    }
    ctx->total_bases = total;
    u64 nwords = (total + 31) / 32;
    ctx->seq_words.resize(nwords + 2, s);
    MTG_CUDA(cudaMemsetAsync(ctx->seq_words.p + nwords, 0, 2 * sizeof(u64), s));
    const char* d_seq = seq;
    DBuf<char> staged;
    if (!on_device && total) {
        staged.resize(total + 32, s);
        MTG_CUDA(cudaMemcpyAsync(staged.p, seq, total, cudaMemcpyHostToDevice, s));
        d_seq = staged.p;
    }
    if (nwords) MTG_LAUNCH(ctx, pack_sequences, grid_for(nwords, TB), TB, 0, d_seq, ctx->seq_words.p, total, nwords, d_err);
    ctx->have_seqs = true;
}

// Shared tail of both construction paths: degrees, imbalance, sources, target bitmap, short-edge CSR.
void finish_graph(mtg_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    const u64 N = ctx->N, E = ctx->E;
    const u32 k = ctx->k;
    ctx->out_deg.resize(N, s);
    ctx->out_deg.zero(s);
    DBuf<u32> deg_s, short_flag, pos, pos_e, from_a, from_b, eid_a, eid_b, src_flag;
    deg_s.resize(N, s);
    deg_s.zero(s);
    short_flag.resize(E, s);
    if (E) MTG_LAUNCH(ctx, count_degrees, grid_for(E, TB), TB, 0, ctx->edge_from.p, ctx->unitig_w.p, E, k, ctx->out_deg.p, deg_s.p, short_flag.p);
    // imbalance / sources / targets
    ctx->imbalance.resize(N, s);
    ctx->target_bits.resize((N + 31) / 32 + 1, s);
    ctx->target_bits.zero(s);
    src_flag.resize(N, s);
    DBuf<unsigned long long> counters_buf;  // owning buffers: released on every exit path, including the throwing ones
    counters_buf.resize(4, s);
    counters_buf.zero(s);
    unsigned long long* d_counters = counters_buf.p;
    if (N) {
        u64 padded = (N + 31) / 32 * 32;
        MTG_LAUNCH(ctx, classify_nodes, grid_for(padded, TB), TB, 0, ctx->out_deg.p, ctx->mirror.p, N, ctx->opt.p2_self_mirror_zero != 0, ctx->imbalance.p, src_flag.p,
                   ctx->target_bits.p, d_counters);
    }
    // source positions, short-edge rows and short-edge positions: three scans, then ONE round trip for all totals
    pos.resize(N + 1, s);
    pos_e.resize(E + 1, s);
    DBuf<u32> tot_buf;
    tot_buf.resize(2, s);
    tot_buf.zero(s);
    u32* d_tot = tot_buf.p;
    exclusive_sum_u32(ctx, src_flag.p, pos.p, N, d_tot);
    ctx->row_s.resize(N + 1, s);
    exclusive_sum_u32(ctx, deg_s.p, ctx->row_s.p, N, ctx->row_s.p + N);
    if (N == 0) MTG_CUDA(cudaMemsetAsync(ctx->row_s.p, 0, sizeof(u32), s));
    exclusive_sum_u32(ctx, short_flag.p, pos_e.p, E, d_tot + 1);
    u32 h_tot[2] = {0, 0};
    unsigned long long h_counters[4];
    MTG_CUDA(cudaMemcpyAsync(h_tot, d_tot, sizeof(h_tot), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaMemcpyAsync(h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->S = h_tot[0];
    ctx->Es = h_tot[1];
    ctx->T = h_counters[0];
    ctx->self_mirror_unbalanced = h_counters[1];
    ctx->target_mult_total = h_counters[2];
    ctx->sources.resize(ctx->S, s);
    if (N) MTG_LAUNCH(ctx, compact_flagged, grid_for(N, TB), TB, 0, src_flag.p, pos.p, N, ctx->sources.p);
    const u64 Es = ctx->Es;
    ctx->col_s.resize(Es, s);
    ctx->w_s.resize(Es, s);
    if (Es) {
        from_a.resize(Es, s);
        from_b.resize(Es, s);
        eid_a.resize(Es, s);
        eid_b.resize(Es, s);
        MTG_LAUNCH(ctx, gather_short_edges, grid_for(E, TB), TB, 0, short_flag.p, pos_e.p, ctx->edge_from.p, E, from_a.p, eid_a.p);
        int which = radix_sort_pairs_u32(ctx, from_a.p, from_b.p, eid_a.p, eid_b.p, Es, bits_for(N));
        MTG_LAUNCH(ctx, fill_short_csr, grid_for(Es, TB), TB, 0, which ? eid_b.p : eid_a.p, ctx->edge_to.p, ctx->unitig_w.p, Es,
                   ctx->col_s.p, ctx->w_s.p);
    }
    for (DBuf<u32>* b : {&deg_s, &short_flag, &pos, &pos_e, &from_a, &from_b, &eid_a, &eid_b, &src_flag}) b->release(s);
    MTG_CUDA(cudaEventRecord(ctx->ev_build[2], s));
    ctx->build_timed = true;
    ctx->have_graph = true;
    ctx->have_cand = ctx->have_triples = ctx->have_walks = false;
    // small graphs prepare the sequential tail on the host: send it its inputs now, behind the build
    ctx->tail_inputs_staged = false;
    if (N < TAIL_HOST_PREP_MAX_NODES && !getenv("MTG_TAIL_HOST")) stage_tail_inputs(ctx);
}

}  // namespace

void build_graph_from_sequences(mtg_ctx* ctx, const char* seq, const u64* offsets, u64 U, u32 k, bool on_device, u64 total_bases,
                                bool prepacked) {
    MTG_REQUIRE(k >= 2 && k <= 64, MTG_ERR_INVALID, "k must be in [2, 64]");
    MTG_REQUIRE(U < (1ull << 30), MTG_ERR_UNSUPPORTED, "more than 2^30 unitigs");
    MTG_REQUIRE(U == 0 || prepacked || (seq && offsets), MTG_ERR_INVALID, "null sequence input");
    cudaStream_t s = ctx->stream;
    ctx->have_graph = false;
    ctx->build_timed = false;
    if (!ctx->in_text_build) {  // no parsing in front of this build
        MTG_CUDA(cudaEventRecord(ctx->ev_build[0], s));
        MTG_CUDA(cudaEventRecord(ctx->ev_build[1], s));
    }
    ctx->k = k;
    ctx->U = U;
    ctx->E = 2 * U;
    DBuf<int> err_buf;
    err_buf.resize(1, s);
    err_buf.zero(s);
    int* d_err = err_buf.p;
    u64 zero_off = 0;
    if (prepacked) {
        ctx->have_seqs = true;
    } else if (U == 0) {
        ctx->seq_off.upload(&zero_off, 1, s);
        ctx->seq_words.resize(2, s);
        ctx->seq_words.zero(s);
        ctx->total_bases = 0;
        ctx->have_seqs = true;
    } else {
        upload_and_pack(ctx, seq, offsets, U, on_device, total_bases, d_err);
    }
    const u64 n = 2 * U;
    const u32 L = k - 1;
    const int nwords = L > 32 ? 2 : 1;
    DBuf<u64> klo_a, klo_b, khi_a, khi_b;
    DBuf<u32> val_a, val_b, head_idx, created, base, node_ep, mirror_ep;
    klo_a.resize(n, s);
    klo_b.resize(n, s);
    if (nwords == 2) {
        khi_a.resize(n, s);
        khi_b.resize(n, s);
    }
    val_a.resize(n, s);
    val_b.resize(n, s);
    ctx->unitig_w.resize(U, s);
    if (n) MTG_LAUNCH(ctx, extract_end_keys, grid_for(n, TB), TB, 0, ctx->seq_words.p, ctx->seq_off.p, U, k, klo_a.p, khi_a.p, val_a.p,
                      ctx->unitig_w.p, d_err);
    // the input error flag is read together with the node count below (sorting garbage keys is harmless)
    int which = radix_sort_pairs(ctx, klo_a.p, klo_b.p, khi_a.p, khi_b.p, val_a.p, val_b.p, n, nwords, 2 * (int)L);
    const u64* klo = which ? klo_b.p : klo_a.p;
    const u64* khi = nwords == 2 ? (which ? khi_b.p : khi_a.p) : nullptr;
    const u32* val = which ? val_b.p : val_a.p;
    head_idx.resize(n, s);
    created.resize(n, s);
    created.zero(s);
    base.resize(n, s);
    node_ep.resize(n, s);
    mirror_ep.resize(n, s);
    DBuf<u32> tot_buf;
    tot_buf.resize(1, s);
    tot_buf.zero(s);
    u32* d_tot = tot_buf.p;
    if (n) {
        MTG_LAUNCH(ctx, mark_groups, grid_for(n, TB), TB, 0, klo, khi, val, n, L, head_idx.p, created.p);
        inclusive_max_u32(ctx, head_idx.p, head_idx.p, n);
        exclusive_sum_u32(ctx, created.p, base.p, n, d_tot);
        MTG_LAUNCH(ctx, assign_nodes, grid_for(n, TB), TB, 0, val, head_idx.p, created.p, base.p, n, node_ep.p, mirror_ep.p);
    }
    u32 h_n = 0;
    int h_err = 0;
    MTG_CUDA(cudaMemcpyAsync(&h_n, d_tot, sizeof(u32), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    throw_on_err_flag(h_err);
    ctx->N = h_n;
    ctx->edge_from.resize(n, s);
    ctx->edge_to.resize(n, s);
    ctx->mirror.resize(ctx->N, s);
    if (U) MTG_LAUNCH(ctx, make_edges, grid_for(U, TB), TB, 0, node_ep.p, mirror_ep.p, U, ctx->edge_from.p, ctx->edge_to.p, ctx->mirror.p);
    for (DBuf<u64>* b : {&klo_a, &klo_b, &khi_a, &khi_b}) b->release(s);
    for (DBuf<u32>* b : {&val_a, &val_b, &head_idx, &created, &base, &node_ep, &mirror_ep}) b->release(s);
    finish_graph(ctx);
}

// k == 0: only the part that does not depend on k (union-find numbering, edges, mirror table, pairing check -- what
// matchtigs_build_graph does, src/clib.rs:180-259); finish_deferred_graph(ctx, k) completes the graph once k is known.
void build_graph_from_links(mtg_ctx* ctx, u64 U, const u64* weights, u64 n_links, const u64* a, const u8* sa, const u64* b,
                            const u8* sb, u32 k, const char* seq, const u64* offsets, bool on_device, u64 total_bases, bool prepacked) {
    MTG_REQUIRE(k == 0 || (k >= 2 && k <= 64), MTG_ERR_INVALID, "k must be in [2, 64]");
    MTG_REQUIRE(U < (1ull << 30), MTG_ERR_UNSUPPORTED, "more than 2^30 unitigs");
    MTG_REQUIRE(U == 0 || weights, MTG_ERR_INVALID, "null weights");
    MTG_REQUIRE(n_links == 0 || (a && sa && b && sb), MTG_ERR_INVALID, "null links");
    MTG_REQUIRE(2 * n_links < 0xFFFFFFFFull, MTG_ERR_UNSUPPORTED, "too many links");
    if (ctx->opt.p6_bcalm_kmer_numbering && ((seq && offsets) || prepacked))  // assumption P6 flipped: the links are ignored, nodes come from the k-mer join
        return build_graph_from_sequences(ctx, seq, offsets, U, k, on_device, total_bases, prepacked);
    cudaStream_t s = ctx->stream;
    ctx->have_graph = false;
    ctx->build_timed = false;
    if (!ctx->in_text_build) {  // no parsing in front of this build
        MTG_CUDA(cudaEventRecord(ctx->ev_build[0], s));
        MTG_CUDA(cudaEventRecord(ctx->ev_build[1], s));
    }
    ctx->have_seqs = prepacked;
    ctx->k = k;
    ctx->U = U;
    ctx->E = 2 * U;
    DBuf<int> err_buf;
    err_buf.resize(1, s);
    err_buf.zero(s);
    int* d_err = err_buf.p;
    if (!prepacked && seq && offsets && U) upload_and_pack(ctx, seq, offsets, U, on_device, total_bases, d_err);
    DBuf<u64> d_w, d_a, d_b;
    DBuf<u8> d_sa, d_sb, rank;
    DBuf<u32> cc, key_a, key_b, op_a, op_b, parent, rep, is_rep, node_of_rep;
    if (on_device) {  // borrowed device arrays (device-side parser): no copies
        d_w.borrow(const_cast<u64*>(weights), U);
        d_a.borrow(const_cast<u64*>(a), n_links);
        d_b.borrow(const_cast<u64*>(b), n_links);
        d_sa.borrow(const_cast<u8*>(sa), n_links);
        d_sb.borrow(const_cast<u8*>(sb), n_links);
    } else {
        d_w.upload(weights, U, s);
        d_a.upload(a, n_links, s);
        d_b.upload(b, n_links, s);
        d_sa.upload(sa, n_links, s);
        d_sb.upload(sb, n_links, s);
    }
    ctx->unitig_w.resize(U, s);
    if (U) MTG_LAUNCH(ctx, weights_to_u32, grid_for(U, TB), TB, 0, d_w.p, U, ctx->unitig_w.p, d_err);
    const u64 nslots = 4 * U, nops = 2 * n_links;
    cc.resize(nslots, s);
    parent.resize(nslots, s);
    rank.resize(nslots, s);
    rank.zero(s);
    if (nslots) {
        MTG_LAUNCH(ctx, iota_u32, grid_for(nslots, TB), TB, 0, cc.p, nslots);
        MTG_LAUNCH(ctx, iota_u32, grid_for(nslots, TB), TB, 0, parent.p, nslots);
    }
    if (nops) {
        MTG_LAUNCH(ctx, cc_hook, grid_for(nops, TB), TB, 0, d_a.p, d_sa.p, d_b.p, d_sb.p, nops, U, cc.p, d_err);
        check_err_flag(ctx, d_err);
        MTG_LAUNCH(ctx, cc_flatten, grid_for(nslots, TB), TB, 0, cc.p, nslots);
        key_a.resize(nops, s);
        key_b.resize(nops, s);
        op_a.resize(nops, s);
        op_b.resize(nops, s);
        MTG_LAUNCH(ctx, op_component_keys, grid_for(nops, TB), TB, 0, d_a.p, d_sa.p, d_b.p, d_sb.p, nops, cc.p, key_a.p, op_a.p);
        int which = radix_sort_pairs_u32(ctx, key_a.p, key_b.p, op_a.p, op_b.p, nops, bits_for(nslots));
        MTG_LAUNCH(ctx, replay_unions, grid_for(nops, TB), TB, 0, which ? key_b.p : key_a.p, which ? op_b.p : op_a.p, nops, d_a.p,
                   d_sa.p, d_b.p, d_sb.p, ctx->opt.p7_first_root_wins != 0, parent.p, rank.p);
    }
    rep.resize(nslots, s);
    is_rep.resize(nslots, s);
    node_of_rep.resize(nslots, s);
    DBuf<u32> tot_buf;
    tot_buf.resize(1, s);
    tot_buf.zero(s);
    u32* d_tot = tot_buf.p;
    if (nslots) {
        MTG_LAUNCH(ctx, find_representatives, grid_for(nslots, TB), TB, 0, parent.p, nslots, rep.p, is_rep.p);
        exclusive_sum_u32(ctx, is_rep.p, node_of_rep.p, nslots, d_tot);
    }
    u32 h_n = 0;
    MTG_CUDA(cudaMemcpyAsync(&h_n, d_tot, sizeof(u32), cudaMemcpyDeviceToHost, s));
    MTG_CUDA(cudaStreamSynchronize(s));
    ctx->N = h_n;
    ctx->edge_from.resize(2 * U, s);
    ctx->edge_to.resize(2 * U, s);
    ctx->mirror.resize(ctx->N, s);
    if (U) {
        MTG_LAUNCH(ctx, make_edges_from_slots, grid_for(U, TB), TB, 0, rep.p, node_of_rep.p, U, ctx->edge_from.p, ctx->edge_to.p, ctx->mirror.p);
        MTG_LAUNCH(ctx, verify_pairing, grid_for(U, TB), TB, 0, ctx->edge_from.p, ctx->edge_to.p, ctx->mirror.p, U, d_err);
    }
    check_err_flag(ctx, d_err);
    for (DBuf<u64>* x : {&d_w, &d_a, &d_b}) x->release(s);
    for (DBuf<u8>* x : {&d_sa, &d_sb, &rank}) x->release(s);
    for (DBuf<u32>* x : {&cc, &key_a, &key_b, &op_a, &op_b, &parent, &rep, &is_rep, &node_of_rep}) x->release(s);
    ctx->graph_needs_k = k == 0;
    if (k) finish_graph(ctx);
}

void finish_deferred_graph(mtg_ctx* ctx, u32 k) {
    MTG_REQUIRE(ctx->graph_needs_k, MTG_ERR_INVALID, "no graph is waiting for its k");
    MTG_REQUIRE(k >= 2 && k <= 64, MTG_ERR_INVALID, "k must be in [2, 64]");
    ctx->k = k;
    ctx->graph_needs_k = false;
    finish_graph(ctx);
}

}  // namespace mtg
