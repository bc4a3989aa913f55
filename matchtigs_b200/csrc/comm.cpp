// comm.cpp -- multi-GPU plumbing inside the library: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no distributed code; its only parallelism is the worker pool over Dijkstra sources
// (greedytigs/mod.rs:557-627).  Here the sources are dealt to the ranks (i % world == rank), every rank searches its share,
// and ONE exchange lines the candidate slices up for the (replicated, deterministic) matching:
//
//   mtg_comm_get_unique_id / mtg_comm_init   rendezvous: rank 0 creates an id, the host program ships its 128 bytes to the
//                                            other ranks any way it likes (file, pipe, MPI, torch.distributed ...)
//   mtg_allgather_candidates                 ncclAllGather of the rank's record slice and meta words on the context's stream
//   mtg_build_graph_from_text_slices         every rank copies 1/world of the file over its own PCIe link, the slices are
//                                            all-gathered over NVLink, then every rank parses and builds the replicated graph
//   mtg_broadcast_walks                      the walks of the rank that ran the sequential tail, for the sharded emission
//                                            (mtg_assemble_tigs_range_view / mtg_dup_bitvector_range_view, emit.cu): every
//                                            rank assembles and downloads its own share of the output bytes
//
// NCCL is bound at run time (dlopen), so that single-GPU users of the library need no NCCL at all and a host that already
// loaded one (torch bundles its own) shares it.
#include <dlfcn.h>

#include <cstring>

#include "mtg_internal.cuh"

using namespace mtg;

namespace {

// the handful of NCCL entry points used, with the types of nccl.h spelled out (no build-time dependency on the header)
struct NcclUniqueId {
    char internal[128];
};
typedef void* NcclComm;
typedef int NcclResult;                                      // ncclSuccess == 0
enum { NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_UINT64 = 5 };  // ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    NcclResult (*GetUniqueId)(NcclUniqueId*) = nullptr;
    NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    NcclResult (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    NcclResult (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(NcclResult) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* names[] = {getenv("MTG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return false;
        }
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(handle, name);
            if (!p) {
                ok = false;
                error = std::string("NCCL symbol missing: ") + name;
            }
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) handle = nullptr;
        return ok;
    }
};
NcclApi g_nccl;

#define MTG_NCCL(expr)                                                                                                 \
    do {                                                                                                               \
        NcclResult _r = (expr);                                                                                        \
        if (_r != 0) throw Error{MTG_ERR_CUDA, std::string(#expr) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?")}; \
    } while (0)

template <class F>
int guarded(mtg_ctx* ctx, F&& f) {
    if (!ctx) return MTG_ERR_INVALID;
    try {
        MTG_CUDA(cudaSetDevice(ctx->device));
        f();
        ctx->err.clear();
        return MTG_OK;
    } catch (const Error& e) {
        ctx->err = e.msg;
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return MTG_ERR_INTERNAL;
    }
}

void require_comm(mtg_ctx* ctx) { MTG_REQUIRE(ctx->comm, MTG_ERR_INVALID, "mtg_comm_init has not run on this context"); }

}  // namespace

extern "C" {

int mtg_comm_get_unique_id(void* id_out) {
    if (!id_out) return MTG_ERR_INVALID;
    if (!g_nccl.load()) {
        fprintf(stderr, "matchtigs_b200: %s\n", g_nccl.error.c_str());
        return MTG_ERR_UNSUPPORTED;
    }
    NcclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return MTG_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return MTG_OK;
}

int mtg_comm_init(mtg_ctx* ctx, const void* unique_id, int rank, int world) {
    return guarded(ctx, [&] {
        MTG_REQUIRE(unique_id && world >= 1 && rank >= 0 && rank < world, MTG_ERR_INVALID, "bad communicator arguments");
        MTG_REQUIRE(!ctx->comm, MTG_ERR_INVALID, "this context already owns a communicator");
        MTG_REQUIRE(g_nccl.load(), MTG_ERR_UNSUPPORTED, g_nccl.error);
        NcclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        NcclComm comm = nullptr;
        MTG_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
        ctx->comm = comm;
        ctx->comm_rank = (u32)rank;
        ctx->comm_world = (u32)world;
    });
}

int mtg_comm_destroy(mtg_ctx* ctx) {
    return guarded(ctx, [&] {
        if (ctx->comm) {
            MTG_CUDA(cudaStreamSynchronize(ctx->stream));
            g_nccl.CommDestroy(ctx->comm);
            ctx->comm = nullptr;
        }
        ctx->gathered_rec.release(ctx->stream);
        ctx->gathered_meta.release(ctx->stream);
    });
}

// Every rank has run mtg_dijkstra_candidates(ctx, cap, rank, world): all slices have the same padded size, so two plain
// all-gathers yield the layout [rank][local source][cap] that mtg_greedy_match indexes directly.
int mtg_allgather_candidates(mtg_ctx* ctx, void** d_records_all, void** d_meta_all) {
    return guarded(ctx, [&] {
        require_comm(ctx);
        MTG_REQUIRE(ctx->have_cand && ctx->shard_count == ctx->comm_world && ctx->shard_rank == ctx->comm_rank, MTG_ERR_INVALID,
                    "mtg_dijkstra_candidates(ctx, cap, rank, world) of this communicator has not run");
        cudaStream_t s = ctx->stream;
        const u64 padded = std::max<u64>((ctx->S + ctx->comm_world - 1) / ctx->comm_world, 1);
        ctx->gathered_rec.resize(padded * ctx->cap * ctx->comm_world, s);
        ctx->gathered_meta.resize(padded * ctx->comm_world, s);
        MTG_NCCL(g_nccl.AllGather(ctx->cand.p, ctx->gathered_rec.p, padded * ctx->cap, NCCL_UINT64, ctx->comm, s));
        MTG_NCCL(g_nccl.AllGather(ctx->cand_meta.p, ctx->gathered_meta.p, padded, NCCL_UINT32, ctx->comm, s));
        if (d_records_all) *d_records_all = ctx->gathered_rec.p;
        if (d_meta_all) *d_meta_all = ctx->gathered_meta.p;
    });
}

// `part` = bytes [rank * slice, min((rank + 1) * slice, total_len)) of the file with slice = mtg_text_slice_bytes(total_len,
// world), in (ideally page-locked) host memory.  H2D over this rank's own link, all-gather over NVLink, then the usual
// device-side parse + build of the replicated graph.
int mtg_build_graph_from_text_slices(mtg_ctx* ctx, const char* part, uint64_t part_len, uint64_t total_len, int bcalm, uint32_t k) {
    return guarded(ctx, [&] {
        require_comm(ctx);
        const u64 slice = mtg_text_slice_bytes(total_len, ctx->comm_world);
        const u64 lo = std::min<u64>((u64)ctx->comm_rank * slice, total_len), hi = std::min<u64>(lo + slice, total_len);
        MTG_REQUIRE(part_len == hi - lo && (part || part_len == 0), MTG_ERR_INVALID, "text slice does not match this rank's range");
        cudaStream_t s = ctx->stream;
        MTG_CUDA(cudaEventRecord(ctx->ev_build[0], s));
        ctx->parse_ws.text.resize(slice * ctx->comm_world + 16, s);
        char* d_text = ctx->parse_ws.text.p;
        if (part_len) MTG_CUDA(cudaMemcpyAsync(d_text + (u64)ctx->comm_rank * slice, part, part_len, cudaMemcpyHostToDevice, s));
        if (slice) MTG_NCCL(g_nccl.AllGather(d_text + (u64)ctx->comm_rank * slice, d_text, slice, NCCL_UINT8, ctx->comm, s));
        ctx->text_event_recorded = true;  // the parser keeps ev_build[0] as the start of the copy
        build_graph_from_text(ctx, d_text, total_len, bcalm != 0, k, true);
    });
}

uint64_t mtg_text_slice_bytes(uint64_t total_len, uint32_t world) {
    if (world == 0) return 0;
    const u64 s = (total_len + world - 1) / world;
    return (s + 15) / 16 * 16;  // keeps every slice 16-byte aligned for the vector loads of the parser
}

// The rank that ran mtg_finish_walks (root) sends its walks to everybody: sizes first, then the three arrays.
int mtg_broadcast_walks(mtg_ctx* ctx, int root) {
    return guarded(ctx, [&] {
        require_comm(ctx);
        MTG_REQUIRE(root >= 0 && (u32)root < ctx->comm_world, MTG_ERR_INVALID, "bad root");
        const bool is_root = (u32)root == ctx->comm_rank;
        MTG_REQUIRE(!is_root || ctx->have_walks, MTG_ERR_INVALID, "the root has no walks: call mtg_finish_walks there first");
        MTG_REQUIRE(ctx->have_graph, MTG_ERR_INVALID, "no graph resident");
        cudaStream_t s = ctx->stream;
        DBuf<u64> sizes;
        sizes.resize(3, s);
        u64 h_sizes[3] = {ctx->walk_edges.size(), ctx->walk_limits.size(), ctx->h_dummy_w.size()};
        if (is_root) MTG_CUDA(cudaMemcpyAsync(sizes.p, h_sizes, sizeof(h_sizes), cudaMemcpyHostToDevice, s));
        MTG_NCCL(g_nccl.Broadcast(sizes.p, sizes.p, 3, NCCL_UINT64, root, ctx->comm, s));
        MTG_CUDA(cudaMemcpyAsync(h_sizes, sizes.p, sizeof(h_sizes), cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        if (!is_root) {
            ctx->d_walk_edges.resize(h_sizes[0], s);
            ctx->d_walk_limits.resize(h_sizes[1], s);
            ctx->d_dummy_w.resize(h_sizes[2], s);
        }
        if (h_sizes[0]) MTG_NCCL(g_nccl.Broadcast(ctx->d_walk_edges.p, ctx->d_walk_edges.p, h_sizes[0], NCCL_UINT32, root, ctx->comm, s));
        if (h_sizes[1]) MTG_NCCL(g_nccl.Broadcast(ctx->d_walk_limits.p, ctx->d_walk_limits.p, h_sizes[1], NCCL_UINT64, root, ctx->comm, s));
        if (h_sizes[2]) MTG_NCCL(g_nccl.Broadcast(ctx->d_dummy_w.p, ctx->d_dummy_w.p, h_sizes[2], NCCL_UINT32, root, ctx->comm, s));
        if (!is_root) {
            // the host copies stay empty on the other ranks: they only assemble output bytes from the device arrays
            ctx->walk_edges.clear();
            ctx->walk_limits.clear();
            ctx->h_dummy_w.clear();
            ctx->n_walk_edges_dev = h_sizes[0];
            ctx->n_walks_dev = h_sizes[1];
            ctx->have_walks = true;
        }
    });
}

}  // extern "C"
