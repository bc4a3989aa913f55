//! `src/implementation/greedytigs/b200.rs` -- FFI declarations and the thin safe wrapper a maintainer of
//! algbio/matchtigs would add to put `libmatchtigs_b200` underneath `GreedytigAlgorithm::compute_tigs`
//! (reference seam: `src/implementation/mod.rs:50-59`, greedy impl `src/implementation/greedytigs/mod.rs:75-90`).
//!
//! THIS FILE IS A SHIM SOURCE, NOT BUILT HERE: the image that builds the CUDA library has no Rust toolchain
//!.  It mirrors `include/matchtigs_b200.h` one to one; `tests/test_host_logic.py` checks
//! that every `fn` declared in the `extern "C"` block below is exported by the built library with the same name.
//!
//! build.rs of the host crate:
//! ```text
//! println!("cargo:rustc-link-search=native={}", env::var("MATCHTIGS_B200_LIB_DIR").unwrap());
//! println!("cargo:rustc-link-lib=static=matchtigs_b200");   // libmatchtigs_b200.a  (or dylib=matchtigs_b200 for the .so)
//! println!("cargo:rustc-link-lib=dylib=cudart");
//! println!("cargo:rustc-link-lib=dylib=gomp");
//! println!("cargo:rustc-link-lib=dylib=stdc++");
//! ```
#![allow(non_camel_case_types, dead_code)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct mtg_ctx {
    _private: [u8; 0],
}

/// `mtg_status` of include/matchtigs_b200.h
pub const MTG_OK: c_int = 0;
pub const MTG_ERR_INVALID: c_int = -1;
pub const MTG_ERR_CUDA: c_int = -2;
pub const MTG_ERR_INPUT: c_int = -3;
pub const MTG_ERR_INTERNAL: c_int = -4;
pub const MTG_ERR_UNSUPPORTED: c_int = -5;
pub const MTG_FORMAT_GFA: c_int = 0;
pub const MTG_FORMAT_FASTA: c_int = 1;
pub const MTG_UNIQUE_ID_BYTES: usize = 128;

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct mtg_graph_info {
    pub unitigs: u64,
    pub nodes: u64,
    pub edges: u64,
    pub short_edges: u64,
    pub sources: u64,
    pub targets: u64,
    pub self_mirrors_unbalanced: u64,
    pub k: u32,
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct mtg_search_stats {
    pub sources_searched: u64,
    pub settled_nodes: u64,
    pub relaxed_edges: u64,
    pub candidates: u64,
    pub truncated_sources: u64,
    pub overflow_sources: u64,
    pub match_rounds: u64,
    pub requery_phases: u64,
    pub matched: u64,
    pub dijkstra_ms: f32,
    pub match_ms: f32,
    pub dijkstra_kernel_ms: f32,
    pub match_kernel_ms: f32,
    pub labelled_nodes: u64,
    pub max_labelled_nodes: u64,
    pub max_open_nodes: u64,
    pub preextended_sources: u64,
}

#[link(name = "matchtigs_b200")]
extern "C" {
    // ---- lifecycle ----
    pub fn mtg_ctx_create(out: *mut *mut mtg_ctx, device: c_int) -> c_int;
    pub fn mtg_ctx_destroy(ctx: *mut mtg_ctx);
    pub fn mtg_last_error(ctx: *const mtg_ctx) -> *const c_char;
    pub fn mtg_ctx_set_option(ctx: *mut mtg_ctx, name: *const c_char, value: c_int) -> c_int;
    pub fn mtg_ctx_stream(ctx: *mut mtg_ctx) -> *mut c_void;
    pub fn mtg_ctx_kernel_launches(ctx: *const mtg_ctx) -> u64;
    // ---- step 1: graph (replaces the genome-graph readers' construction, bin.rs:896-899 / :907-910, clib.rs:135-259) ----
    pub fn mtg_build_graph_from_sequences(ctx: *mut mtg_ctx, seq_ascii: *const c_char, offsets: *const u64, unitigs: u64, k: u32,
                                          seq_on_device: c_int) -> c_int;
    pub fn mtg_build_graph_from_links(ctx: *mut mtg_ctx, unitigs: u64, weights: *const u64, n_links: u64, link_a: *const u64,
                                      strand_a: *const u8, link_b: *const u64, strand_b: *const u8, k: u32,
                                      seq_ascii: *const c_char, offsets: *const u64) -> c_int;
    pub fn mtg_build_graph_from_text(ctx: *mut mtg_ctx, text: *const c_char, len: u64, bcalm: c_int, k: u32, text_on_device: c_int) -> c_int;
    pub fn mtg_graph_get_info(ctx: *mut mtg_ctx, info: *mut mtg_graph_info) -> c_int;
    pub fn mtg_graph_export(ctx: *mut mtg_ctx, edge_from: *mut u32, edge_to: *mut u32, mirror: *mut u32, imbalance: *mut i32,
                            sources: *mut u32) -> c_int;
    // ---- step 2: many-source bounded Dijkstra (greedytigs/mod.rs:301-335, 557-627) ----
    pub fn mtg_dijkstra_candidates(ctx: *mut mtg_ctx, cap: u32, shard_rank: u32, shard_count: u32) -> c_int;
    pub fn mtg_candidates_local(ctx: *mut mtg_ctx, d_records: *mut *mut c_void, d_meta: *mut *mut c_void, sources_local: *mut u64,
                                cap: *mut u32) -> c_int;
    pub fn mtg_candidates_export(ctx: *mut mtg_ctx, nodes: *mut u32, dists: *mut u32, meta: *mut u32) -> c_int;
    // ---- step 3: matching (greedytigs/mod.rs:350-502) ----
    pub fn mtg_greedy_match(ctx: *mut mtg_ctx, d_records_all: *const c_void, d_meta_all: *const c_void, shard_count: u32,
                            n_triples: *mut u64) -> c_int;
    pub fn mtg_triples_export(ctx: *mut mtg_ctx, triples: *mut u32) -> c_int;
    // ---- tail (greedytigs/mod.rs:678-789, implementation/mod.rs:392-649) ----
    pub fn mtg_finish_walks(ctx: *mut mtg_ctx, n_walks: *mut u64, n_walk_edges: *mut u64) -> c_int;
    pub fn mtg_walks_export(ctx: *mut mtg_ctx, walk_edges: *mut u32, walk_limits: *mut u64) -> c_int;
    pub fn mtg_walks_export_capi(ctx: *mut mtg_ctx, tigs_edge_out: *mut isize, tigs_insert_out: *mut usize, tigs_out_limits: *mut usize) -> c_int;
    pub fn mtg_walk_count(ctx: *mut mtg_ctx, n_walks: *mut u64) -> c_int;
    // ---- outputs (bin.rs:466-818, implementation/mod.rs:671-702) ----
    pub fn mtg_dup_bitvector(ctx: *mut mtg_ctx, out: *mut c_char, cap: u64, out_len: *mut u64) -> c_int;
    pub fn mtg_assemble_tigs(ctx: *mut mtg_ctx, format: c_int, out: *mut c_char, cap: u64, out_len: *mut u64) -> c_int;
    pub fn mtg_dup_bitvector_view(ctx: *mut mtg_ctx, out: *mut *const c_char, out_len: *mut u64) -> c_int;
    pub fn mtg_assemble_tigs_view(ctx: *mut mtg_ctx, format: c_int, out: *mut *const c_char, out_len: *mut u64) -> c_int;
    pub fn mtg_dup_bitvector_range_view(ctx: *mut mtg_ctx, walk_lo: u64, walk_hi: u64, out: *mut *const c_char, out_len: *mut u64,
                                        byte_offset: *mut u64, total_len: *mut u64) -> c_int;
    pub fn mtg_assemble_tigs_range_view(ctx: *mut mtg_ctx, format: c_int, walk_lo: u64, walk_hi: u64, out: *mut *const c_char,
                                        out_len: *mut u64, byte_offset: *mut u64, total_len: *mut u64) -> c_int;
    // ---- multi-GPU: one process per GPU, NCCL bound at run time ----
    pub fn mtg_comm_get_unique_id(id_out: *mut c_void) -> c_int;
    pub fn mtg_comm_init(ctx: *mut mtg_ctx, unique_id: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn mtg_comm_destroy(ctx: *mut mtg_ctx) -> c_int;
    pub fn mtg_allgather_candidates(ctx: *mut mtg_ctx, d_records_all: *mut *mut c_void, d_meta_all: *mut *mut c_void) -> c_int;
    pub fn mtg_text_slice_bytes(total_len: u64, world: u32) -> u64;
    pub fn mtg_build_graph_from_text_slices(ctx: *mut mtg_ctx, part: *const c_char, part_len: u64, total_len: u64, bcalm: c_int, k: u32) -> c_int;
    pub fn mtg_broadcast_walks(ctx: *mut mtg_ctx, root: c_int) -> c_int;
    // ---- diagnostics ----
    pub fn mtg_get_search_stats(ctx: *mut mtg_ctx, stats: *mut mtg_search_stats) -> c_int;
    pub fn mtg_get_diagnostics(ctx: *mut mtg_ctx, tail_ms: *mut f64, build_ms: *mut f64) -> c_int;
}

/// Error of a step: the status code and the library's message.
#[derive(Debug)]
pub struct B200Error {
    pub code: c_int,
    pub message: String,
}

/// Owning handle of one GPU context.
pub struct B200Context {
    ctx: *mut mtg_ctx,
}

impl B200Context {
    pub fn new(device: i32) -> Result<Self, B200Error> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { mtg_ctx_create(&mut ctx, device) };
        if rc != MTG_OK {
            return Err(B200Error { code: rc, message: "no usable CUDA device (there is no CPU fallback)".into() });
        }
        Ok(Self { ctx })
    }

    fn check(&self, rc: c_int) -> Result<(), B200Error> {
        if rc == MTG_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(mtg_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(B200Error { code: rc, message })
    }

    /// What `GreedytigAlgorithm::compute_tigs` does on the GPU path.  `seq`/`offsets`: the unitig characters the
    /// sequence store holds (bin.rs:871), in edge order.  Returns the walks as edge ids with the numbering of the
    /// reference: `2u` = unitig `u` forward, `2u + 1` = its mirror, ids `>= 2U` = dummy edges in insertion order
    /// (greedytigs/mod.rs:683-688), plus the end offset of every walk.
    pub fn greedytigs(&mut self, seq: &[u8], offsets: &[u64], k: u32, cap: u32) -> Result<(Vec<u32>, Vec<u64>), B200Error> {
        let unitigs = (offsets.len() - 1) as u64;
        unsafe {
            self.check(mtg_build_graph_from_sequences(self.ctx, seq.as_ptr() as *const c_char, offsets.as_ptr(), unitigs, k, 0))?;
            self.check(mtg_dijkstra_candidates(self.ctx, cap, 0, 1))?;
            let mut n_triples = 0u64;
            self.check(mtg_greedy_match(self.ctx, std::ptr::null(), std::ptr::null(), 1, &mut n_triples))?;
            let (mut n_walks, mut n_edges) = (0u64, 0u64);
            self.check(mtg_finish_walks(self.ctx, &mut n_walks, &mut n_edges))?;
            let mut edges = vec![0u32; n_edges as usize];
            let mut limits = vec![0u64; n_walks as usize];
            self.check(mtg_walks_export(self.ctx, edges.as_mut_ptr(), limits.as_mut_ptr()))?;
            Ok((edges, limits))
        }
    }

    /// GFA bytes assembled on the GPU (stands where `write_walks_gfa`, bin.rs:667-818, stands); valid until the next call.
    pub fn gfa(&mut self) -> Result<&[u8], B200Error> {
        let (mut p, mut n) = (std::ptr::null(), 0u64);
        self.check(unsafe { mtg_assemble_tigs_view(self.ctx, MTG_FORMAT_GFA, &mut p, &mut n) })?;
        Ok(unsafe { std::slice::from_raw_parts(p as *const u8, n as usize) })
    }

    /// Duplicate-k-mer bitvector (`write_duplication_bitvector`, implementation/mod.rs:671-702).
    pub fn duplication_bitvector(&mut self) -> Result<&[u8], B200Error> {
        let (mut p, mut n) = (std::ptr::null(), 0u64);
        self.check(unsafe { mtg_dup_bitvector_view(self.ctx, &mut p, &mut n) })?;
        Ok(unsafe { std::slice::from_raw_parts(p as *const u8, n as usize) })
    }
}

impl Drop for B200Context {
    fn drop(&mut self) {
        unsafe { mtg_ctx_destroy(self.ctx) }
    }
}
