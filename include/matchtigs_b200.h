/*
 * matchtigs_b200.h -- C ABI of the B200-native greedy-matchtig hot path.
 *
 * Two groups of entry points:
 *
 *  (1) `matchtigs_*`: the reference's own C API, names and signatures unchanged
 *      (reference: src/clib.rs:90, :97, :135-141, :180-183, :280-290).  A program linked against
 *      the Rust `libmatchtigs` dylib can link against this library instead.  Only
 *      tig_algorithm 1 (unitigs) and 5 (greedy matchtigs -- see the id quirk at src/clib.rs:367-389)
 *      are served; 2/3/4 (pathtigs, eulertigs, Blossom-V matchtigs) stay reference-only and
 *      make the call fail loudly.
 *
 *  (2) `mtg_*`: the step API the (Rust) host calls underneath `TigAlgorithm::compute_tigs`
 *      (reference seam: src/implementation/mod.rs:50-59, greedy impl src/implementation/greedytigs/mod.rs:75-90).
 *      One entry per north-star step; each cites the reference region it replaces.
 *
 * Conventions: plain C types only; every `mtg_*` function returns 0 on success or a negative
 * `mtg_status`, never throws or aborts, and leaves a message retrievable with mtg_last_error().
 * A context is bound to one CUDA device and is not re-entrant.  There is no CPU fallback:
 * without a usable CUDA device mtg_ctx_create fails with MTG_ERR_CUDA.
 */
#ifndef MATCHTIGS_B200_H
#define MATCHTIGS_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mtg_status {
    MTG_OK = 0,
    MTG_ERR_INVALID = -1,  /* bad argument / call order */
    MTG_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    MTG_ERR_INPUT = -3,    /* malformed input data (non-ACGT, sequence shorter than k, inconsistent links) */
    MTG_ERR_INTERNAL = -4, /* invariant violated (the reference would panic) */
    MTG_ERR_UNSUPPORTED = -5
} mtg_status;

typedef struct mtg_ctx mtg_ctx;

/* Sizes of the resident graph. */
typedef struct mtg_graph_info {
    uint64_t unitigs;       /* U */
    uint64_t nodes;         /* N: (k-1)-mer nodes incl. mirrors */
    uint64_t edges;         /* 2U */
    uint64_t short_edges;   /* directed edges with weight <= k-1 (all the Dijkstra can traverse) */
    uint64_t sources;       /* S: out-imbalanced nodes (+ odd self-mirrors), ascending node id */
    uint64_t targets;       /* in-imbalanced nodes (+ odd self-mirrors) */
    uint64_t self_mirrors_unbalanced;
    uint32_t k;
} mtg_graph_info;

/* Counters of the last mtg_dijkstra_candidates / mtg_greedy_match calls. */
typedef struct mtg_search_stats {
    uint64_t sources_searched;  /* sources that had at least one short out-edge */
    uint64_t settled_nodes;     /* nodes popped with a final label (numerator of settled nodes/s) */
    uint64_t relaxed_edges;     /* short out-edges inspected */
    uint64_t candidates;        /* (dst, dist) records emitted */
    uint64_t truncated_sources; /* lists cut at `cap` */
    uint64_t overflow_sources;  /* sources that outgrew the shared-memory table and used the global-memory tier */
    uint64_t match_rounds;      /* blocked attempts (spins) of the dataflow matching kernel */
    uint64_t requery_phases;    /* extra search phases for sources whose capped list ran dry */
    uint64_t matched;           /* triples produced */
    float dijkstra_ms;          /* device time of the search kernels (CUDA events) */
    float match_ms;             /* device time of the matching step (set-up kernels, sort, dataflow kernel, copies) */
    float dijkstra_kernel_ms;   /* device time of the tier-0 (thread-per-source) search kernel of the last main search */
    float match_kernel_ms;      /* device time of the dataflow matching kernel (first phase) */
    /* What the reference's DijkstraPerformanceCounter reports (--dijkstra-performance-data-type Complete,
     * greedytigs/mod.rs:647-673), as far as it has a meaning here: a search labels nodes in a table ("distance array"); its
     * open labels are what a heap would hold.  There is no lazy deletion (decrease-key happens in place), so no label is ever
     * extracted twice: iterations == settled_nodes, unnecessary heap elements == 0. */
    uint64_t labelled_nodes;      /* labelled nodes summed over the searches */
    uint64_t max_labelled_nodes;  /* most labelled nodes in one search ("maximum maximum distance array size") */
    uint64_t max_open_nodes;      /* most open labels at once in one thread-tier search ("maximum maximum heap size") */
    uint64_t preextended_sources; /* truncated lists that mtg_greedy_match searched again with 8x cap before matching */
} mtg_search_stats;

/* ---- lifecycle ---- */
int mtg_ctx_create(mtg_ctx** out, int device);
void mtg_ctx_destroy(mtg_ctx* ctx);
const char* mtg_last_error(const mtg_ctx* ctx);
/* Parity assumptions about the reference's un-vendored dependencies as switches; 0 = as assumed.
 * Names: p1_tie_desc, p1_exclusive_bound (Dijkstra settle order / bound, traitgraph-algo), p2_self_mirror_zero (imbalance
 * of self-mirror nodes, bigraph), p3_oldest_first (adjacency iteration order, petgraph), p6_bcalm_kmer_numbering (bcalm2
 * reader numbering, genome-graph), p7_first_root_wins (union-find tie rule, disjoint-sets).  Same names as the oracle's
 * mto_set_option; MTG_ASSUME_<NAME>=1 in the environment sets the default.  Invalidates the resident graph. */
int mtg_ctx_set_option(mtg_ctx* ctx, const char* name, int value);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on; for event timing by the caller. */
void* mtg_ctx_stream(mtg_ctx* ctx);
/* Number of kernels this context has launched since creation (the caller's gpu_launches claim). */
uint64_t mtg_ctx_kernel_launches(const mtg_ctx* ctx);

/* ---- step 1: bidirected unitig overlap graph ----
 * Replaces read_bigraph_from_fasta_as_edge_centric (call site src/bin.rs:896-899) + compute_edge_weights
 * (src/bin.rs:359-379) + the imbalance scan (src/implementation/greedytigs/mod.rs:222-245).
 * `seq_ascii`: the U unitig sequences concatenated (ACGT only), `offsets[U+1]`: start of each unitig.
 * Unitigs are packed to 2 bit, the canonical (k-1)-mer keys of both ends are extracted, radix-sorted and
 * joined into node ids that equal the reference reader's first-seen numbering; CSR, mirror table,
 * imbalances, sources and the target bitmap stay resident on the device.
 * `seq_on_device` != 0: both pointers are device pointers (inputs already resident in HBM). */
int mtg_build_graph_from_sequences(mtg_ctx* ctx, const char* seq_ascii, const uint64_t* offsets, uint64_t unitigs,
                                   uint32_t k, int seq_on_device);

/* Replaces matchtigs_merge_nodes + matchtigs_build_graph (src/clib.rs:135-170, :180-259) and, for --bcalm-in,
 * read_bigraph_from_bcalm2_as_edge_centric (call site src/bin.rs:907-910).  Links are (a, strand_a, b, strand_b)
 * in call order; `weights[U]` = k-mers per unitig.  Optional sequences (may be NULL) are only packed for output. */
int mtg_build_graph_from_links(mtg_ctx* ctx, uint64_t unitigs, const uint64_t* weights, uint64_t n_links,
                               const uint64_t* link_a, const uint8_t* strand_a, const uint64_t* link_b,
                               const uint8_t* strand_b, uint32_t k, const char* seq_ascii, const uint64_t* offsets);

/* Same two builders fed by the device-side record parser: `text` is the raw FASTA (bcalm == 0, --fa-in semantics) or
 * bcalm2 FASTA (bcalm != 0, --bcalm-in semantics: ids must equal positions, `L:` fields become links) file content,
 * < 32 GiB and < 2^32 bases, on the host or (text_on_device != 0) already in HBM.  Line splitting, header/sequence separation,
 * multi-line records, id validation and link extraction all run as scans on the GPU
 * (replaces the record parsing of genome-graph's readers, call sites src/bin.rs:896-899, :907-910). */
int mtg_build_graph_from_text(mtg_ctx* ctx, const char* text, uint64_t len, int bcalm, uint32_t k, int text_on_device);

int mtg_graph_get_info(mtg_ctx* ctx, mtg_graph_info* info);
/* Copies the graph to host arrays (any pointer may be NULL): edge_from/edge_to [2U] (edge 2u = unitig u forward,
 * 2u+1 = its mirror), mirror [N], imbalance [N], sources [S]. */
int mtg_graph_export(mtg_ctx* ctx, uint32_t* edge_from, uint32_t* edge_to, uint32_t* mirror, int32_t* imbalance,
                     uint32_t* sources);

/* ---- step 2: many-source bounded Dijkstra ----
 * Replaces the Dijkstra::shortest_path_lens calls and their work distribution
 * (src/implementation/greedytigs/mod.rs:301-335, :557-627).  For every source with index i, i % shard_count ==
 * shard_rank, computes the first `cap` initially-open in-nodes within distance k-1 in settle order
 * (dist, node id) plus a `truncated` flag.  Results stay on the device. */
int mtg_dijkstra_candidates(mtg_ctx* ctx, uint32_t cap, uint32_t shard_rank, uint32_t shard_count);
/* Device pointers of this rank's candidate slice for the NVLink exchange: records are uint64
 * (node | dist << 32), `cap` per source; meta is uint32 per source (count | truncated << 31).
 * Local source l corresponds to global source l * shard_count + shard_rank. */
int mtg_candidates_local(mtg_ctx* ctx, void** d_records, void** d_meta, uint64_t* sources_local, uint32_t* cap);
/* Copies local candidate lists to the host (tests): nodes/dists [sources_local * cap], meta [sources_local]. */
int mtg_candidates_export(mtg_ctx* ctx, uint32_t* nodes, uint32_t* dists, uint32_t* meta);

/* ---- step 3: source-ordered greedy matching ----
 * Replaces the matching body of compute_dijkstras (src/implementation/greedytigs/mod.rs:350-502) under the
 * --threads 1 semantics.  `d_records_all` / `d_meta_all`: gathered candidate slices of all ranks, laid out
 * [shard][local source] (pass NULL, NULL, 1 to use this context's own full-range result).
 * Output: triples (out_node, in_node, dist) in the order of the reference's `results` vector. */
int mtg_greedy_match(mtg_ctx* ctx, const void* d_records_all, const void* d_meta_all, uint32_t shard_count,
                     uint64_t* n_triples);
int mtg_triples_export(mtg_ctx* ctx, uint32_t* triples /* 3 * n_triples */);

/* ---- host-sequential tail kept for byte parity (runs on the host inside the library) ----
 * Dummy-edge insertion (greedytigs/mod.rs:678-689), make_graph_eulerian_with_breaking_edges
 * (src/implementation/mod.rs:392-649), Euler decomposition (greedytigs/mod.rs:722) and cycle breaking
 * (greedytigs/mod.rs:726-789).  Produces walks over edge ids: original edge e < 2U, dummy edges >= 2U. */
int mtg_finish_walks(mtg_ctx* ctx, uint64_t* n_walks, uint64_t* n_walk_edges);
/* walk_edges [n_walk_edges], walk_limits [n_walks] (end offsets).  Edge ids >= 2 * unitigs are dummy edges in insertion
 * order; their weights are what mtg_walks_export_capi reports in tigs_insert_out. */
int mtg_walks_export(mtg_ctx* ctx, uint32_t* walk_edges, uint64_t* walk_limits);
/* C-API encoding of the walks (src/clib.rs:393-407). */
int mtg_walks_export_capi(mtg_ctx* ctx, ptrdiff_t* tigs_edge_out, size_t* tigs_insert_out, size_t* tigs_out_limits);

/* The same tail on caller-supplied host arrays, no GPU involved (what mtg_finish_walks runs after copying
 * the graph and the triples back).  edge_from/edge_to [2*unitigs], unitig_w [unitigs], mirror [nodes],
 * triples [3*n_triples].  Outputs are malloc'd (release with mtg_host_free): walk_edges, walk_limits (end
 * offsets), dummy_w (weight of dummy edge e at [e - 2*unitigs]).  phase_ms[5] (optional): degrees, eulerise,
 * adjacency, Euler walk, breaking. */
int mtg_host_tail(uint32_t k, uint64_t nodes, uint64_t unitigs, const uint32_t* edge_from, const uint32_t* edge_to,
                  const uint32_t* unitig_w, const uint32_t* mirror, const uint32_t* triples, uint64_t n_triples,
                  uint32_t** walk_edges, uint64_t** walk_limits, uint32_t** dummy_w, uint64_t* n_walks,
                  uint64_t* n_walk_edges, uint64_t* n_dummy_edges, double* phase_ms, char* errbuf, size_t errcap);
void mtg_host_free(void* p);

/* ---- step 3b: outputs ----
 * Duplicate-k-mer bitvector (src/implementation/mod.rs:671-702) and tig strings
 * (write_walks_gfa src/bin.rs:667-818, write_walks_fasta :466-606), assembled on the device from the
 * 2-bit store.  Call with out == NULL to obtain the required size. */
int mtg_dup_bitvector(mtg_ctx* ctx, char* out, uint64_t cap, uint64_t* out_len);
typedef enum mtg_text_format { MTG_FORMAT_GFA = 0, MTG_FORMAT_FASTA = 1 } mtg_text_format;
int mtg_assemble_tigs(mtg_ctx* ctx, int format, char* out, uint64_t cap, uint64_t* out_len);
/* Zero-copy variants: *out points into page-locked memory owned by the context (filled by DMA at full PCIe
 * speed) and stays valid until the next call that produces the same kind of text or mtg_ctx_destroy. */
int mtg_dup_bitvector_view(mtg_ctx* ctx, const char** out, uint64_t* out_len);
int mtg_assemble_tigs_view(mtg_ctx* ctx, int format, const char** out, uint64_t* out_len);

/* Shares of the two texts: only the bytes that belong to the walks [walk_lo, walk_hi) (clamped to the walk count; the GFA
 * header line belongs to the share that starts at walk 0).  *byte_offset = where the share starts inside the whole text,
 * *total_len = length of the whole text.  With the walks on every rank (mtg_broadcast_walks) each rank assembles its
 * share on its own GPU and brings it to the host over its own PCIe link. */
int mtg_dup_bitvector_range_view(mtg_ctx* ctx, uint64_t walk_lo, uint64_t walk_hi, const char** out, uint64_t* out_len,
                                 uint64_t* byte_offset, uint64_t* total_len);
int mtg_assemble_tigs_range_view(mtg_ctx* ctx, int format, uint64_t walk_lo, uint64_t walk_hi, const char** out,
                                 uint64_t* out_len, uint64_t* byte_offset, uint64_t* total_len);
/* Number of walks resident on the device (after mtg_finish_walks, or mtg_broadcast_walks on the receiving ranks). */
int mtg_walk_count(mtg_ctx* ctx, uint64_t* n_walks);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch, bound at run time ----
 * The reference has no distributed code (its parallelism is the worker pool of greedytigs/mod.rs:557-627); here the
 * Dijkstra sources are dealt to the ranks (source i belongs to rank i % world), the CSR graph is replicated, and ONE
 * exchange -- an all-gather of the equally sized candidate slices -- feeds the replicated, deterministic matching.
 *   rank 0: mtg_comm_get_unique_id(id); ship the MTG_UNIQUE_ID_BYTES bytes to the other ranks (file, pipe, MPI, ...)
 *   all:    mtg_comm_init(ctx, id, rank, world)
 *           mtg_build_graph_from_text_slices(...)      (or any mtg_build_graph_* with the whole input on every rank)
 *           mtg_dijkstra_candidates(ctx, cap, rank, world); mtg_allgather_candidates(ctx, &rec, &meta);
 *           mtg_greedy_match(ctx, rec, meta, world, &n)
 *   rank 0: mtg_finish_walks(ctx, ...)                 (the host-sequential tail)
 *   all:    mtg_broadcast_walks(ctx, 0); mtg_*_range_view(ctx, ..., T * rank / world, T * (rank + 1) / world, ...) */
#define MTG_UNIQUE_ID_BYTES 128
int mtg_comm_get_unique_id(void* id_out /* MTG_UNIQUE_ID_BYTES */);
int mtg_comm_init(mtg_ctx* ctx, const void* unique_id, int rank, int world);
int mtg_comm_destroy(mtg_ctx* ctx);
/* Collective.  Returns device pointers (owned by the context) to pass to mtg_greedy_match together with `world`. */
int mtg_allgather_candidates(mtg_ctx* ctx, void** d_records_all, void** d_meta_all);
/* Collective.  Every rank supplies bytes [rank * slice, min((rank + 1) * slice, total_len)) of the input file, slice =
 * mtg_text_slice_bytes(total_len, world): 1/world of the H2D traffic per PCIe link, the rest travels over NVLink. */
uint64_t mtg_text_slice_bytes(uint64_t total_len, uint32_t world);
int mtg_build_graph_from_text_slices(mtg_ctx* ctx, const char* part, uint64_t part_len, uint64_t total_len, int bcalm,
                                     uint32_t k);
/* Collective.  The walks of `root` (which ran mtg_finish_walks) become resident on every rank's device. */
int mtg_broadcast_walks(mtg_ctx* ctx, int root);

/* Everything above in order (single GPU): build -> search -> match -> walks. */
int mtg_compute_greedytigs_from_sequences(mtg_ctx* ctx, const char* seq_ascii, const uint64_t* offsets,
                                          uint64_t unitigs, uint32_t k, uint32_t cap);
int mtg_get_search_stats(mtg_ctx* ctx, mtg_search_stats* stats);
/* Diagnostics of the last run (either pointer may be NULL): host-tail phases in ms (preparation = degree counting on the
 * host-prepared path / record kernels on the device-prepared path, eulerise, walk records = built on the host / DMA + copy,
 * Euler walk, breaking) and device time of the last graph build in ms: build_ms[0] = H2D copy + record parsing
 * (mtg_build_graph_from_text only, else 0), build_ms[1] = graph construction proper (pack .. CSR). */
int mtg_get_diagnostics(mtg_ctx* ctx, double tail_ms[5], double build_ms[2]);

/* ---- host-side record reader ----
 * Splits FASTA / bcalm2 text into the arrays the step API takes; stands where genome-graph's readers stand
 * (call sites src/bin.rs:896-899, 907-910).  `bcalm` != 0: record ids must equal their position and
 * `L:<+/->:<id>:<+/->` header fields are returned as links in file order.  Character validation
 * (ACGT only) happens on the device while packing. */
typedef struct mtg_unitigs mtg_unitigs;
int mtg_unitigs_parse(const char* text, size_t len, int bcalm, mtg_unitigs** out, char* errbuf, size_t errcap);
void mtg_unitigs_free(mtg_unitigs* u);
int mtg_unitigs_view(const mtg_unitigs* u, const char** seq, const uint64_t** offsets, uint64_t* unitigs,
                     const uint64_t** link_a, const uint8_t** strand_a, const uint64_t** link_b,
                     const uint8_t** strand_b, uint64_t* n_links);

/* ---- the reference's C API (src/clib.rs) ---- */
typedef struct MatchtigsData MatchtigsData;
void matchtigs_initialise(void);                                          /* src/clib.rs:90 */
MatchtigsData* matchtigs_initialise_graph(size_t unitig_amount);          /* src/clib.rs:97 */
void matchtigs_merge_nodes(MatchtigsData* matchtigs_data, size_t unitig_a, bool strand_a, size_t unitig_b,
                           bool strand_b);                                /* src/clib.rs:135-141 */
void matchtigs_build_graph(MatchtigsData* matchtigs_data, const size_t* unitig_weights); /* src/clib.rs:180-183 */
size_t matchtigs_compute_tigs(MatchtigsData* matchtigs_data, size_t tig_algorithm, size_t threads, size_t k,
                              const char* matching_file_prefix, const char* matcher_path, ptrdiff_t* tigs_edge_out,
                              size_t* tigs_insert_out, size_t* tigs_out_limits); /* src/clib.rs:280-290 */

#ifdef __cplusplus
}
#endif
#endif /* MATCHTIGS_B200_H */
