// api.cpp -- the C ABI declared in include/matchtigs_b200.h: exception firewall around the
// step functions plus the reference's own C API (src/clib.rs) on top of them.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include <memory>

#include "mtg_internal.cuh"

using namespace mtg;

namespace {

template <class F>
int guarded(mtg_ctx* ctx, F&& f) {
    if (!ctx) return MTG_ERR_INVALID;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) throw Error{MTG_ERR_CUDA, std::string("cudaSetDevice failed: ") + cudaGetErrorString(e)};
        f();
        ctx->err.clear();
        return MTG_OK;
    } catch (const Error& e) {
        ctx->err = e.msg;
        // leave the device in a defined state for the next call
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
        return e.code;
    } catch (const std::bad_alloc&) {
        ctx->err = "out of host memory";
        return MTG_ERR_INTERNAL;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return MTG_ERR_INTERNAL;
    } catch (...) {
        ctx->err = "unknown error";
        return MTG_ERR_INTERNAL;
    }
}

}  // namespace

static int* option_slot(mtg_ctx* ctx, const char* name) {
    static const struct {
        const char* name;
        int mtg_options::*field;
    } table[] = {{"p1_tie_desc", &mtg_options::p1_tie_desc},
                 {"p1_exclusive_bound", &mtg_options::p1_exclusive_bound},
                 {"p2_self_mirror_zero", &mtg_options::p2_self_mirror_zero},
                 {"p3_oldest_first", &mtg_options::p3_oldest_first},
                 {"p6_bcalm_kmer_numbering", &mtg_options::p6_bcalm_kmer_numbering},
                 {"p7_first_root_wins", &mtg_options::p7_first_root_wins}};
    for (const auto& t : table)
        if (!strcmp(name, t.name)) return &(ctx->opt.*(t.field));
    return nullptr;
}

// ---- device block cache (declared in mtg_internal.cuh) ----
namespace mtg {
namespace {
struct BlockCache {
    std::mutex m;
    std::map<std::pair<cudaStream_t, size_t>, std::vector<void*>> free_blocks;  // (stream, class) -> blocks
    size_t free_bytes = 0;  // bytes sitting in the free lists
    size_t limit = 0;       // beyond this, released blocks go straight back to the driver (0 = not initialised yet)
};
BlockCache& block_cache() {
    static BlockCache* c = new BlockCache();  // never destroyed: contexts may outlive static destruction order
    return *c;
}
}  // namespace

size_t block_class_bytes(size_t bytes) {
    if (bytes <= 512) return 512;
    int top = 63 - __builtin_clzll((unsigned long long)bytes);  // eight classes per power of two: at most 12.5 % over
    const size_t step = size_t(1) << (top - 3);
    return (bytes + step - 1) / step * step;
}

void* block_alloc(size_t class_bytes, cudaStream_t s) {
    BlockCache& c = block_cache();
    {
        std::lock_guard<std::mutex> g(c.m);
        auto it = c.free_blocks.find({s, class_bytes});
        if (it != c.free_blocks.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            c.free_bytes -= class_bytes;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, class_bytes, s);
    if (e == cudaErrorMemoryAllocation) {  // hand the cached blocks of this stream back and try once more
        cudaGetLastError();
        block_cache_trim(s);
        e = cudaMallocAsync(&p, class_bytes, s);
    }
    if (e != cudaSuccess) throw Error{MTG_ERR_CUDA, std::string("cudaMallocAsync failed: ") + cudaGetErrorString(e)};
    return p;
}

void block_free(void* p, size_t class_bytes, cudaStream_t s) {
    BlockCache& c = block_cache();
    std::lock_guard<std::mutex> g(c.m);
    if (c.limit == 0) {
        // idle blocks may hold up to a quarter of the device (MTG_BLOCK_CACHE_MAX_MB overrides): enough for the scratch
        // of any job that fits the device next to its resident graph, bounded for a process that runs many different jobs
        size_t free_b = 0, total_b = 0;
        if (const char* e = getenv("MTG_BLOCK_CACHE_MAX_MB")) c.limit = (size_t)std::max(1L, atol(e)) << 20;
        else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c.limit = std::max<size_t>(total_b / 4, size_t(1) << 30);
        else c.limit = size_t(16) << 30;
        cudaGetLastError();
    }
    if (c.free_bytes + class_bytes > c.limit) {
        cudaFreeAsync(p, s);
        return;
    }
    c.free_blocks[{s, class_bytes}].push_back(p);
    c.free_bytes += class_bytes;
}

void block_cache_trim(cudaStream_t s) {
    BlockCache& c = block_cache();
    std::lock_guard<std::mutex> g(c.m);
    for (auto it = c.free_blocks.begin(); it != c.free_blocks.end();) {
        if (it->first.first == s) {
            for (void* p : it->second) cudaFreeAsync(p, s);
            c.free_bytes -= it->first.second * it->second.size();
            it = c.free_blocks.erase(it);
        } else {
            ++it;
        }
    }
}
}  // namespace mtg

extern "C" {

int mtg_ctx_create(mtg_ctx** out, int device) {
    if (!out) return MTG_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return MTG_ERR_CUDA;  // no CPU fallback: without a CUDA device there is no context
    }
    mtg_ctx* ctx = new (std::nothrow) mtg_ctx();
    if (!ctx) return MTG_ERR_INTERNAL;
    ctx->device = device;
    for (const char* name : {"p1_tie_desc", "p1_exclusive_bound", "p2_self_mirror_zero", "p3_oldest_first", "p6_bcalm_kmer_numbering",
                             "p7_first_root_wins"}) {  // MTG_ASSUME_P1_TIE_DESC=1 ... in the environment
        std::string env = std::string("MTG_ASSUME_") + name;
        for (auto& c : env) c = (char)toupper((unsigned char)c);
        if (const char* v = getenv(env.c_str())) *option_slot(ctx, name) = atoi(v);
    }
    int rc = guarded(ctx, [&] {
        MTG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        MTG_CUDA(cudaEventCreate(&ctx->ev0));
        MTG_CUDA(cudaEventCreate(&ctx->ev1));
        MTG_CUDA(cudaEventCreate(&ctx->ev2));
        MTG_CUDA(cudaEventCreate(&ctx->ev3));
        for (auto& e : ctx->ev_build) MTG_CUDA(cudaEventCreate(&e));
        for (auto& e : ctx->tail_events) MTG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaDeviceProp prop{};
        MTG_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->num_sms = prop.multiProcessorCount;
        MTG_REQUIRE(prop.cooperativeLaunch, MTG_ERR_CUDA, "device lacks cooperative launch");
        cudaMemPool_t pool;
        MTG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;  // keep freed blocks cached: repeated runs reuse them
        MTG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    });
    if (rc != MTG_OK) {
        fprintf(stderr, "matchtigs_b200: context creation failed: %s\n", ctx->err.c_str());
        mtg_ctx_destroy(ctx);  // also releases the stream and the events created so far
        return rc;
    }
    *out = ctx;
    return MTG_OK;
}

void mtg_ctx_destroy(mtg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    mtg_comm_destroy(ctx);
    cudaStream_t s = ctx->stream;
    ctx->seq_words.release(s);
    ctx->seq_off.release(s);
    ctx->unitig_w.release(s);
    ctx->edge_from.release(s);
    ctx->edge_to.release(s);
    ctx->mirror.release(s);
    ctx->out_deg.release(s);
    ctx->imbalance.release(s);
    ctx->target_bits.release(s);
    ctx->sources.release(s);
    ctx->row_s.release(s);
    ctx->col_s.release(s);
    ctx->w_s.release(s);
    ctx->cand.release(s);
    ctx->cand_meta.release(s);
    ctx->dstats.release(s);
    ctx->triples.release(s);
    ctx->final_mult.release(s);
    ctx->d_walk_edges.release(s);
    ctx->d_walk_limits.release(s);
    ctx->d_dummy_w.release(s);
    ctx->scratch_a.release(s);
    ctx->scratch_b.release(s);
    for (mtg::DBuf<mtg::u32>* b : {&ctx->parse_ws.key, &ctx->parse_ws.n_rec, &ctx->parse_ws.n_seq, &ctx->parse_ws.n_link, &ctx->parse_ws.rbase,
                                   &ctx->parse_ws.sbase, &ctx->parse_ws.lbase, &ctx->parse_ws.totals})
        b->release(s);
    ctx->parse_ws.text.release(s);
    for (auto& b : ctx->text_stage) b.release();
    for (auto& b : ctx->tail_stage) b.release();
    ctx->h_triples.release();
    ctx->walk_edges.release();
    ctx->gathered_rec.release(s);
    ctx->gathered_meta.release(s);
    mtg::block_cache_trim(s);  // everything this context's stream ever cached goes back to the driver
    if (s) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (auto& e : ctx->ev_build)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->tail_events)
        if (e) cudaEventDestroy(e);
    delete ctx;
}

int mtg_ctx_set_option(mtg_ctx* ctx, const char* name, int value) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(name, MTG_ERR_INVALID, "null option name");
        int* slot = option_slot(ctx, name);
        MTG_REQUIRE(slot, MTG_ERR_INVALID, std::string("unknown option ") + name);
        *slot = value;
        ctx->have_graph = ctx->have_cand = ctx->have_triples = ctx->have_walks = false;  // results of other assumptions are void
    });
}

const char* mtg_last_error(const mtg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* mtg_ctx_stream(mtg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t mtg_ctx_kernel_launches(const mtg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mtg_build_graph_from_sequences(mtg_ctx* ctx, const char* seq_ascii, const uint64_t* offsets, uint64_t unitigs, uint32_t k,
                                   int seq_on_device) {
    return guarded(ctx, [&] { build_graph_from_sequences(ctx, seq_ascii, offsets, unitigs, k, seq_on_device != 0); });
}

int mtg_build_graph_from_links(mtg_ctx* ctx, uint64_t unitigs, const uint64_t* weights, uint64_t n_links, const uint64_t* link_a,
                               const uint8_t* strand_a, const uint64_t* link_b, const uint8_t* strand_b, uint32_t k,
                               const char* seq_ascii, const uint64_t* offsets) {
    return guarded(ctx, [&] {
        build_graph_from_links(ctx, unitigs, weights, n_links, link_a, strand_a, link_b, strand_b, k, seq_ascii, offsets, false);
    });
}

int mtg_build_graph_from_text(mtg_ctx* ctx, const char* text, uint64_t len, int bcalm, uint32_t k, int text_on_device) {
    return guarded(ctx, [&] { build_graph_from_text(ctx, text, len, bcalm != 0, k, text_on_device != 0); });
}

int mtg_graph_get_info(mtg_ctx* ctx, mtg_graph_info* info) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(info && ctx->have_graph, MTG_ERR_INVALID, "no graph resident");
        info->unitigs = ctx->U;
        info->nodes = ctx->N;
        info->edges = ctx->E;
        info->short_edges = ctx->Es;
        info->sources = ctx->S;
        info->targets = ctx->T;
        info->self_mirrors_unbalanced = ctx->self_mirror_unbalanced;
        info->k = ctx->k;
    });
}

int mtg_graph_export(mtg_ctx* ctx, uint32_t* edge_from, uint32_t* edge_to, uint32_t* mirror, int32_t* imbalance, uint32_t* sources) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_graph, MTG_ERR_INVALID, "no graph resident");
        cudaStream_t s = ctx->stream;
        if (edge_from) ctx->edge_from.download(edge_from, s);
        if (edge_to) ctx->edge_to.download(edge_to, s);
        if (mirror) ctx->mirror.download(mirror, s);
        if (imbalance) ctx->imbalance.download(imbalance, s);
        if (sources) ctx->sources.download(sources, s);
        MTG_CUDA(cudaStreamSynchronize(s));
    });
}

int mtg_dijkstra_candidates(mtg_ctx* ctx, uint32_t cap, uint32_t shard_rank, uint32_t shard_count) {
    return guarded(ctx, [&] { dijkstra_candidates(ctx, cap, shard_rank, shard_count); });
}

int mtg_candidates_local(mtg_ctx* ctx, void** d_records, void** d_meta, uint64_t* sources_local, uint32_t* cap) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_cand, MTG_ERR_INVALID, "mtg_dijkstra_candidates has not run");
        if (d_records) *d_records = ctx->cand.p;
        if (d_meta) *d_meta = ctx->cand_meta.p;
        if (sources_local) *sources_local = ctx->S_local;
        if (cap) *cap = ctx->cap;
    });
}

int mtg_candidates_export(mtg_ctx* ctx, uint32_t* nodes, uint32_t* dists, uint32_t* meta) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_cand, MTG_ERR_INVALID, "mtg_dijkstra_candidates has not run");
        cudaStream_t s = ctx->stream;
        const size_t n = (size_t)ctx->S_local * ctx->cap;
        std::vector<u64> rec(n);
        if (n) MTG_CUDA(cudaMemcpyAsync(rec.data(), ctx->cand.p, n * sizeof(u64), cudaMemcpyDeviceToHost, s));
        if (meta && ctx->S_local)
            MTG_CUDA(cudaMemcpyAsync(meta, ctx->cand_meta.p, ctx->S_local * sizeof(u32), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < n; i++) {
            if (nodes) nodes[i] = (u32)rec[i];
            if (dists) dists[i] = (u32)(rec[i] >> 32);
        }
    });
}

int mtg_greedy_match(mtg_ctx* ctx, const void* d_records_all, const void* d_meta_all, uint32_t shard_count, uint64_t* n_triples) {
    return guarded(ctx, [&] {
        greedy_match(ctx, (const u64*)d_records_all, (const u32*)d_meta_all, shard_count);
        if (n_triples) *n_triples = ctx->n_triples;
    });
}

int mtg_triples_export(mtg_ctx* ctx, uint32_t* triples) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_triples, MTG_ERR_INVALID, "mtg_greedy_match has not run");
        if (triples && ctx->n_triples) memcpy(triples, ctx->h_triples.data(), ctx->h_triples.size() * sizeof(u32));
    });
}

int mtg_finish_walks(mtg_ctx* ctx, uint64_t* n_walks, uint64_t* n_walk_edges) {
    return guarded(ctx, [&] {
        finish_walks(ctx);
        if (n_walks) *n_walks = ctx->walk_limits.size();
        if (n_walk_edges) *n_walk_edges = ctx->walk_edges.size();
    });
}

int mtg_walks_export(mtg_ctx* ctx, uint32_t* walk_edges, uint64_t* walk_limits) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_walks, MTG_ERR_INVALID, "mtg_finish_walks has not run");
        if (walk_edges && !ctx->walk_edges.empty()) memcpy(walk_edges, ctx->walk_edges.data(), ctx->walk_edges.size() * sizeof(u32));
        if (walk_limits && !ctx->walk_limits.empty())
            memcpy(walk_limits, ctx->walk_limits.data(), ctx->walk_limits.size() * sizeof(u64));
    });
}

int mtg_walks_export_capi(mtg_ctx* ctx, ptrdiff_t* tigs_edge_out, size_t* tigs_insert_out, size_t* tigs_out_limits) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_walks, MTG_ERR_INVALID, "mtg_finish_walks has not run");
        MTG_REQUIRE(tigs_edge_out && tigs_insert_out && tigs_out_limits, MTG_ERR_INVALID, "null output array");
        // src/clib.rs:393-407: +-unitig id (dummies carry the default handle 0), dummy weight or 0, end offsets
        for (size_t j = 0; j < ctx->walk_edges.size(); j++) {
            u32 e = ctx->walk_edges[j];
            bool fwd = !(e & 1);
            if (e >= ctx->E) {
                tigs_edge_out[j] = 0;
                tigs_insert_out[j] = ctx->h_dummy_w[e - ctx->E];
            } else {
                tigs_edge_out[j] = (ptrdiff_t)(e >> 1) * (fwd ? 1 : -1);
                tigs_insert_out[j] = 0;
            }
        }
        for (size_t i = 0; i < ctx->walk_limits.size(); i++) tigs_out_limits[i] = (size_t)ctx->walk_limits[i];
    });
}

int mtg_dup_bitvector(mtg_ctx* ctx, char* out, uint64_t cap, uint64_t* out_len) {
    return guarded(ctx, [&] {
        u64 n = dup_bitvector(ctx, out, cap, out == nullptr, nullptr);
        if (out_len) *out_len = n;
    });
}

int mtg_dup_bitvector_view(mtg_ctx* ctx, const char** out, uint64_t* out_len) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(out && out_len, MTG_ERR_INVALID, "null output");
        *out_len = dup_bitvector(ctx, nullptr, 0, false, out);
    });
}

int mtg_assemble_tigs_view(mtg_ctx* ctx, int format, const char** out, uint64_t* out_len) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(out && out_len, MTG_ERR_INVALID, "null output");
        *out_len = assemble_tigs(ctx, format, nullptr, 0, false, out);
    });
}

int mtg_dup_bitvector_range_view(mtg_ctx* ctx, uint64_t walk_lo, uint64_t walk_hi, const char** out, uint64_t* out_len, uint64_t* byte_offset,
                                 uint64_t* total_len) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(out && out_len, MTG_ERR_INVALID, "null output");
        u64 where[2];
        *out_len = dup_bitvector(ctx, nullptr, 0, false, out, walk_lo, walk_hi, where);
        if (byte_offset) *byte_offset = where[0];
        if (total_len) *total_len = where[1];
    });
}

int mtg_assemble_tigs_range_view(mtg_ctx* ctx, int format, uint64_t walk_lo, uint64_t walk_hi, const char** out, uint64_t* out_len,
                                 uint64_t* byte_offset, uint64_t* total_len) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(out && out_len, MTG_ERR_INVALID, "null output");
        u64 where[2];
        *out_len = assemble_tigs(ctx, format, nullptr, 0, false, out, walk_lo, walk_hi, where);
        if (byte_offset) *byte_offset = where[0];
        if (total_len) *total_len = where[1];
    });
}

int mtg_walk_count(mtg_ctx* ctx, uint64_t* n_walks) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(ctx->have_walks && n_walks, MTG_ERR_INVALID, "no walks resident");
        *n_walks = ctx->n_walks_dev;
    });
}

int mtg_assemble_tigs(mtg_ctx* ctx, int format, char* out, uint64_t cap, uint64_t* out_len) {
    return guarded(ctx, [&] {
        u64 n = assemble_tigs(ctx, format, out, cap, out == nullptr, nullptr);
        if (out_len) *out_len = n;
    });
}

int mtg_compute_greedytigs_from_sequences(mtg_ctx* ctx, const char* seq_ascii, const uint64_t* offsets, uint64_t unitigs, uint32_t k,
                                          uint32_t cap) {
    return guarded(ctx, [&] {
        build_graph_from_sequences(ctx, seq_ascii, offsets, unitigs, k, false);
        dijkstra_candidates(ctx, cap, 0, 1);
        greedy_match(ctx, nullptr, nullptr, 1);
        finish_walks(ctx);
    });
}

int mtg_get_search_stats(mtg_ctx* ctx, mtg_search_stats* stats) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(stats, MTG_ERR_INVALID, "null stats");
        *stats = ctx->stats;
    });
}

int mtg_get_diagnostics(mtg_ctx* ctx, double tail_ms[5], double build_ms[2]) {
    return guarded(ctx, [&] {
        if (tail_ms) memcpy(tail_ms, ctx->tail_ms, sizeof(ctx->tail_ms));
        if (build_ms) {
            build_ms[0] = build_ms[1] = 0;
            if (ctx->build_timed) {
                float a = 0, b = 0;
                MTG_CUDA(cudaEventSynchronize(ctx->ev_build[2]));
                MTG_CUDA(cudaEventElapsedTime(&a, ctx->ev_build[0], ctx->ev_build[1]));
                MTG_CUDA(cudaEventElapsedTime(&b, ctx->ev_build[1], ctx->ev_build[2]));
                build_ms[0] = a, build_ms[1] = b;
            }
        }
    });
}

// =====================================================================================
// The reference's C API (src/clib.rs).  Same names, argument meaning and error behaviour:
// the reference has no error channel and panics, so failures print a message and abort().
// =====================================================================================
struct MatchtigsData {
    size_t unitig_amount = 0;
    std::vector<u64> a, b;
    std::vector<u8> sa, sb;
    mtg_ctx* ctx = nullptr;  // owns the device graph between matchtigs_build_graph and matchtigs_compute_tigs
    bool built = false;
    ~MatchtigsData() {
        if (ctx) mtg_ctx_destroy(ctx);
    }
};

static bool g_initialised = false;

[[noreturn]] static void clib_panic(const char* what) {
    fprintf(stderr, "matchtigs (b200): %s\n", what);
    abort();
}

void matchtigs_initialise(void) {
    // src/clib.rs:90 -> initialise_logging(...).unwrap(): a second call panics (src/implementation/mod.rs:38-44)
    if (g_initialised) clib_panic("matchtigs_initialise called twice");
    g_initialised = true;
}

MatchtigsData* matchtigs_initialise_graph(size_t unitig_amount) {
    MatchtigsData* d = new MatchtigsData();
    d->unitig_amount = unitig_amount;
    return d;
}

void matchtigs_merge_nodes(MatchtigsData* d, size_t unitig_a, bool strand_a, size_t unitig_b, bool strand_b) {
    if (!d) clib_panic("matchtigs_merge_nodes: null handle");
    if (unitig_a >= d->unitig_amount || unitig_b >= d->unitig_amount) clib_panic("matchtigs_merge_nodes: unitig id out of range");
    d->a.push_back(unitig_a);
    d->sa.push_back(strand_a ? 1 : 0);
    d->b.push_back(unitig_b);
    d->sb.push_back(strand_b ? 1 : 0);
}

void matchtigs_build_graph(MatchtigsData* d, const size_t* unitig_weights) {
    if (!d) clib_panic("matchtigs_build_graph: null handle");
    if (!unitig_weights) clib_panic("matchtigs_build_graph: unitig_weights is null");  // assert! src/clib.rs:188
    if (d->built) clib_panic("matchtigs_build_graph called twice");
    // Like the reference this call builds the graph -- union-find numbering, edges, mirror table -- and asserts
    // verify_node_pairing / verify_edge_mirror_property (src/clib.rs:251-252): inconsistent links abort HERE, not one call
    // later.  Only what depends on k (weights against k-1, short-edge CSR) waits for matchtigs_compute_tigs.
    int dev = 0;
    if (const char* e = getenv("MTG_DEVICE")) dev = atoi(e);
    if (mtg_ctx_create(&d->ctx, dev) != MTG_OK) clib_panic("no usable CUDA device (there is no CPU fallback)");
    std::vector<u64> w(unitig_weights, unitig_weights + d->unitig_amount);
    const int rc = guarded(d->ctx, [&] {
        build_graph_from_links(d->ctx, d->unitig_amount, w.data(), d->a.size(), d->a.data(), d->sa.data(), d->b.data(), d->sb.data(), 0,
                               nullptr, nullptr, false);
    });
    if (rc != MTG_OK) clib_panic(mtg_last_error(d->ctx));
    d->built = true;
}

size_t matchtigs_compute_tigs(MatchtigsData* d, size_t tig_algorithm, size_t threads, size_t k, const char* matching_file_prefix,
                              const char* matcher_path, ptrdiff_t* tigs_edge_out, size_t* tigs_insert_out, size_t* tigs_out_limits) {
    (void)threads;
    if (!d) clib_panic("matchtigs_compute_tigs: null handle");
    std::unique_ptr<MatchtigsData> owned(d);  // the handle is consumed, like Box::from_raw (src/clib.rs:291)
    if (!d->built) clib_panic("matchtigs_compute_tigs: matchtigs_build_graph was not called");
    if (!matching_file_prefix) clib_panic("matching_file_prefix is null");  // assert! src/clib.rs:300
    if (!matcher_path) clib_panic("matcher_path is null");                  // assert! src/clib.rs:316
    if (!tigs_edge_out || !tigs_insert_out || !tigs_out_limits) clib_panic("output array is null");  // :333-345
    const size_t U = d->unitig_amount;
    if (tig_algorithm == 1) {  // unitigs: every forward edge on its own (src/clib.rs:351-361)
        for (size_t u = 0; u < U; u++) {
            tigs_edge_out[u] = (ptrdiff_t)u;
            tigs_insert_out[u] = 0;
            tigs_out_limits[u] = u + 1;
        }
        return U;
    }
    if (tig_algorithm != 5) {
        if (tig_algorithm >= 2 && tig_algorithm <= 4)
            clib_panic("tig algorithms 2 (pathtigs), 3 (eulertigs) and 4 (matchtigs) are reference-only; this library serves 1 and 5");
        clib_panic("Unknown tigs algorithm identifier");  // src/clib.rs:390
    }
    mtg_ctx* ctx = d->ctx;
    auto check = [&](int rc) {
        if (rc != MTG_OK) {
            fprintf(stderr, "matchtigs (b200): %s\n", mtg_last_error(ctx));
            abort();
        }
    };
    check(guarded(ctx, [&] { finish_deferred_graph(ctx, (u32)k); }));
    check(mtg_dijkstra_candidates(ctx, 8, 0, 1));
    uint64_t nt = 0, nw = 0, nwe = 0;
    check(mtg_greedy_match(ctx, nullptr, nullptr, 1, &nt));
    check(mtg_finish_walks(ctx, &nw, &nwe));
    check(mtg_walks_export_capi(ctx, tigs_edge_out, tigs_insert_out, tigs_out_limits));
    return (size_t)nw;  // `owned` releases the handle and its context
}

}  // extern "C"
