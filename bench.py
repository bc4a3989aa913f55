#!/usr/bin/env python
"""bench.py -- greedy-matchtig throughput on synthetic unitigs (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--scale F]

One "step" = one pass of the whole hot path over one batch of unitigs:
graph build (2-bit pack, end keys, radix sort, join, CSR) -> many-source bounded Dijkstra ->
greedy matching -> host Euler tail -> duplicate-k-mer bitvector + GFA assembly.

* ``value``: unitigs/s with the unitig characters already resident in HBM when the timed region starts.
* ``e2e``:   the same through the public API with HOST buffers: the unitig FASTA text (page-locked) is copied to
             the device and parsed there, and the GFA + bitvector bytes come back to page-locked host memory,
             all inside the timed region -- the same span the reference arm times (parse -> outputs).
* ``roofline``: the Dijkstra tier-0 kernel (one thread per source) against measured HBM bandwidth.
* ``cpu_baseline``: the CPU oracle (a C++ restatement of matchtigs 2.1.9 greedy, NOT the Rust binary) on
  the same workload, 1 thread (the deterministic reference semantics).
* ``--impl reference``: the same oracle with all host threads it can use (the reference's worker scheme).

N > 1 (torchrun): the graph is replicated, Dijkstra sources are sharded i % N == rank, candidate slices
are all-gathered over NCCL, the matching is replicated, rank 0 finishes the walks and outputs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# torchrun exports OMP_NUM_THREADS=1 for every rank; the host-side helpers of this bench (the synthetic unitig builder on rank 0)
# are OpenMP code that should use the box's cores.  (The library itself sizes its host threads from the CPU affinity mask.)
if os.environ.get("OMP_NUM_THREADS") == "1":
    del os.environ["OMP_NUM_THREADS"]

import numpy as np  # noqa: E402

WORKLOADS = {
    # name -> recipe of tools.config_unitigs, default scale, reader the BASELINE config names, sample scale of the reference arm
    "ecoli": {"cfg": "ecoli", "scale": 1.0, "bcalm": False, "ref_scale": 1.0,
              "desc": "configs[1]: synthetic 4.6 Mbp E. coli-sized genome with injected repeats, k=31, --fa-in"},
    "pangenome": {"cfg": "pangenome", "scale": 1.0, "bcalm": False, "ref_scale": 0.25,
                  "desc": "configs[3]: synthetic pangenome, 100 E. coli-like strains with SNP/indel variation, k=31, --fa-in"},
    "chr1": {"cfg": "chr1", "scale": 1.0, "bcalm": True, "ref_scale": 0.1,
             "desc": "configs[2]: synthetic 250 Mbp chr1-sized genome with repeat families, k=31, --bcalm-in"},
    "human": {"cfg": "human", "scale": 0.2, "bcalm": True, "ref_scale": 0.02,
              "desc": "configs[4] (scaled: the synthetic unitig builder needs ~34 B of host RAM per distinct k-mer): "
                      "synthetic human-like genome, k=51, --bcalm-in"},
}
DEFAULT_WORKLOAD = "chr1"  # the largest BASELINE config that fits one GPU and whose input this box can build
CAP = 16
OUTPUTS = "GFA + duplicate-kmer bitvector"


def config_of(args, world: int) -> dict:
    """The `config` object: recipe-level keys only, so that both arms print the same one."""
    w = WORKLOADS[args.workload]
    scale = w["scale"] if args.scale is None else args.scale
    return {"workload": args.workload, "description": w["desc"], "k": 51 if w["cfg"] == "human" else 31, "scale": scale,
            "reader": "--bcalm-in (links -> union-find numbering)" if w["bcalm"] else "--fa-in (k-mer join)",
            "outputs": OUTPUTS, "candidate_cap": CAP,
            "l2": "flushed between timed iterations (256 MiB memset); the workload's text alone exceeds L2",
            "parallelism": f"sources sharded over {world} GPU(s), graph replicated"}


def load_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name: str, scale: float | None, rank: int = 0):
    """Unitigs of the named recipe (bcalm2-style FASTA text).  Built once per box and cached on disk (tools.cache_dir());
    under torchrun rank 0 builds and the other ranks wait for the file."""
    import tools
    w = WORKLOADS[name]
    scale = w["scale"] if scale is None else scale
    t0 = time.time()
    text, k, info = tools.cached_config_unitigs(w["cfg"], scale, wait_for_other=rank != 0)
    info["generate_s"] = round(time.time() - t0, 2)
    info["description"] = w["desc"]
    return text, k, info


# ----------------------------------------------------------------------------------------------
# reference arm: the CPU oracle with every host thread (C++ restatement, not the Rust binary)
# ----------------------------------------------------------------------------------------------
ORACLE_PHASES = ("parse", "build", "scan", "dijkstra", "insert", "eulerise", "euler", "break", "bitvector", "write")
ORACLE_COMPUTE = ("build", "scan", "dijkstra", "insert", "eulerise", "euler", "break", "bitvector")  # T_compute


def oracle_pass(text: bytes, k: int, bcalm: bool, threads: int):
    """One whole CPU pass producing exactly the GPU arm's outputs (GFA + bitvector, no FASTA, no C-API arrays).
    Returns (oracle, wall seconds, phase seconds)."""
    import oracle
    o = oracle.Oracle(euler_fast=True, outputs=oracle.Oracle.OUT_GFA | oracle.Oracle.OUT_BITVECTOR)
    t0 = time.perf_counter()
    (o.load_bcalm if bcalm else o.load_fasta)(text, k)
    o.run(threads=threads)
    wall = time.perf_counter() - t0
    return o, wall, {n: o.time(n) for n in ORACLE_PHASES}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    w = WORKLOADS[args.workload]
    scale = w["scale"] if args.scale is None else args.scale
    # bounded sample: the same recipe at a reduced genome length, so that warmup + steps whole passes end within minutes
    # (a chr1-size pass of the restatement takes ~17 s on 8-16 cores: 25 of them would take 7 minutes)
    sample_scale = min(scale, w["ref_scale"]) if args.ref_scale is None else args.ref_scale
    text, k, info = make_workload(args.workload, sample_scale)
    cores = os.cpu_count() or 1
    U = info["unitigs"]
    times, phases, settled, dj = [], {}, 0, 0.0
    for it in range(args.warmup + args.steps):
        o, dt, ph = oracle_pass(text, k, w["bcalm"], cores)
        if it >= args.warmup:
            times.append(dt)
            for n, v in ph.items():
                phases[n] = phases.get(n, 0.0) + v / args.steps
            settled, dj = o.num("settled"), o.time("dijkstra")
        del o
    ms = 1e3 * float(np.mean(times))
    value = U / (ms / 1e3)
    t_compute = sum(phases[n] for n in ORACLE_COMPUTE)
    sample = (f"each step = one whole pass (parse -> {OUTPUTS}) over the {args.workload} recipe at scale {sample_scale:g} "
              f"({U} unitigs, {info['input_bp']} bp); C++ restatement of matchtigs 2.1.9 greedy with the reference's worker scheme "
              f"on {cores} threads (not the Rust binary, which cannot be built in this image)")
    line = {
        "impl": "reference", "metric": "greedy_matchtig_unitigs_per_sec", "value": value, "unit": "unitigs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_of(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "unitigs/s", "cores": cores, "kind": "port", "sample": sample,
                         "sample_scale": sample_scale, "sample_unitigs": U},
        "e2e": {"value": value, "unit": "unitigs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "t_wall_ms": ms, "t_compute_ms": 1e3 * t_compute,
        "unitigs_per_sec_compute": U / t_compute if t_compute > 0 else None,
        "phases_ms": {n: 1e3 * v for n, v in phases.items()},
        "settled_nodes": {"reference_semantics": settled},
        "settled_nodes_per_sec": settled / dj if dj > 0 else None,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import matchtigs_b200 as mt
    from matchtigs_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; matchtigs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=40))

    w = WORKLOADS[args.workload]
    bcalm = w["bcalm"]
    text, k, info = make_workload(args.workload, args.scale, rank)
    text_host = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()  # the unitig file content, page-locked
    text_dev = text_host.cuda()                                                   # ... and resident in HBM
    ctx = mt.Context(local_rank)
    if world > 1:
        # rendezvous of the library's own NCCL communicator: rank 0 creates the id, torch.distributed only carries its bytes
        uid = [mt.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        lo_b, hi_b = sharding.text_slice(text_host.numel(), rank, world)
        text_part = text_host[lo_b:hi_b].clone().pin_memory()  # this rank's share of the file, page-locked
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def link_gbps(to_device: bool) -> float:
        """Pinned-memory copy bandwidth of this box's host link (context for the e2e number; boxes differ a lot)."""
        h = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
        d = flush[: h.numel()]
        best = 0.0
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            (d.copy_(h, non_blocking=True) if to_device else h.copy_(d, non_blocking=True))
            b.record()
            b.synchronize()
            best = max(best, h.numel() / (a.elapsed_time(b) * 1e-3) / 1e9)
        return best

    link = {"h2d_GBps": link_gbps(True), "d2h_GBps": link_gbps(False)} if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    phase_s = {}  # host wall time per library call, summed over the timed steps of the current leg (every call blocks)

    def call(name, fn, *a, **kw):
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        phase_s[name] = phase_s.get(name, 0.0) + time.perf_counter() - t0
        return r

    def step(resident: bool):
        """One pass of the hot path.  Returns (gfa, bitvector) on rank 0."""
        if resident:
            # the file content already in HBM: records parsed on the device -> graph
            call("parse+graph_build", ctx.build_graph_from_text, None, k, bcalm=bcalm, device_ptr=text_dev.data_ptr(),
                 length=text_dev.numel())
        elif world == 1:
            # end to end: raw file bytes in page-locked host memory -> H2D -> parsed on the device -> graph
            call("parse+graph_build", ctx.build_graph_from_text, text_host.numpy(), k, bcalm=bcalm)
        else:
            # every rank uploads 1/world of the file over its own PCIe link; the slices are all-gathered over NVLink
            call("parse+graph_build", ctx.build_graph_from_text_slices, text_part.numpy(), text_host.numel(), k, bcalm)
        dg = ctx.diagnostics()
        phase_s["(parse, device)"] = phase_s.get("(parse, device)", 0.0) + dg["build_ms"]["parse"] * 1e-3
        if world == 1:
            call("dijkstra", ctx.dijkstra_candidates, CAP, 0, 1)
            call("match", ctx.greedy_match, export=False)
        else:
            call("dijkstra", ctx.dijkstra_candidates, CAP, rank, world)
            rec_all, meta_all = call("all_gather", ctx.allgather_candidates)  # ncclAllGather inside the library, over NVLink
            call("match", ctx.greedy_match, rec_all, meta_all, world, export=False)
        if rank == 0:
            call("tail", ctx.finish_walks)
        if world == 1:
            # results land in page-locked host memory owned by the context (zero-copy views)
            bv = call("emit_bitvector", ctx.dup_bitvector_view)
            return call("emit_gfa", ctx.assemble_tigs_view, "gfa"), bv
        # sharded emission: the walks go to every rank, each assembles and downloads its share of the two texts
        call("broadcast_walks", ctx.broadcast_walks, 0)
        w_lo, w_hi = sharding.walk_share(ctx.walk_count(), rank, world)
        bv = call("emit_bitvector", ctx.dup_bitvector_range_view, w_lo, w_hi)
        return call("emit_gfa", ctx.assemble_tigs_range_view, "gfa", w_lo, w_hi), bv

    def timed(resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(resident)
        barrier()
        total_ms, dj_ms, mt_ms, djk_ms, mtk_ms, stats, out = 0.0, 0.0, 0.0, 0.0, 0.0, None, None
        per_step, tail, per_step_phases = [], {}, []
        launches0 = ctx.kernel_launches
        phase_s.clear()
        for _ in range(steps):
            before = dict(phase_s)
            with torch.cuda.stream(ext):
                flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(ext)
            out = step(resident)
            e1.record(ext)
            barrier()
            per_step.append(e0.elapsed_time(e1))
            per_step_phases.append({n: round(1e3 * (v - before.get(n, 0.0)), 2) for n, v in phase_s.items()})
            total_ms += per_step[-1]
            stats = ctx.search_stats()
            dj_ms += stats["dijkstra_ms"]
            mt_ms += stats["match_ms"]
            djk_ms += stats["dijkstra_kernel_ms"]
            mtk_ms += stats["match_kernel_ms"]
            if rank == 0:
                for n, v in ctx.diagnostics()["tail_ms"].items():
                    tail[n] = tail.get(n, 0.0) + v / steps
        launches = ctx.kernel_launches - launches0
        t = torch.tensor([total_ms, dj_ms, mt_ms, djk_ms, mtk_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, dj_ms, mt_ms, djk_ms, mtk_ms = (float(x) for x in t.cpu())
        slowest = max(range(steps), key=lambda i: per_step[i])
        slow_info = {"step": slowest, "ms": per_step[slowest], "phases_ms": per_step_phases[slowest]}
        per_step.sort()
        stats = dict(stats, dijkstra_kernel_ms=djk_ms / steps, match_kernel_ms=mtk_ms / steps,
                     spread={"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1], "slowest": slow_info},
                     phases_ms={n: 1e3 * v / steps for n, v in phase_s.items()}, tail_ms=tail)
        return total_ms / steps, dj_ms / steps, mt_ms / steps, stats, launches, out

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("MTG_BENCH_NOSAMPLER"):  # (experiments only: does the polling disturb the run?)
        sampler.start()
    ms_res, dj_ms, match_ms, stats, launches, _ = timed(True, args.steps, args.warmup)
    if args.profile:
        print(json.dumps({"profile_run": True, "ms_per_step": ms_res, "phases_ms": stats["phases_ms"], "tail_ms": stats["tail_ms"]}), flush=True)
        ctx.close()
        return
    ms_e2e, _, _, stats_e2e, _, out = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # settled / relaxed / candidates summed over ranks for the roofline numerator
    cnt = torch.tensor([stats["settled_nodes"], stats["relaxed_edges"], stats["candidates"], stats["sources_searched"]],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    settled, relaxed, cands, searched = (float(x) for x in cnt.cpu())

    parts = None
    if world > 1:
        # (offset, length, sha256) of every rank's share of the two texts, for the byte comparison with the oracle on rank 0
        import hashlib
        mine = [(int(off), int(v.size), hashlib.sha256(v.tobytes()).hexdigest(), int(tot)) for (v, off, tot) in out]
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank == 0:
        gi = ctx.graph_info()
        U = gi["unitigs"]
        if world == 1:
            gfa, bv = (x.tobytes() for x in out)
            out_bytes = len(gfa) + len(bv)
        else:
            out_bytes = sum(p[0][1] + p[1][1] for p in parts)
        peak, peak_src = load_peaks()
        # algorithmic bytes of the Dijkstra kernel: 12 B per settled node (row_ptr pair + target probe),
        # 5 B per relaxed short edge (col + weight), 8 B per emitted candidate
        alg_bytes = 12.0 * settled + 5.0 * relaxed + 8.0 * cands
        # duration: CUDA events around the main search kernel on the library's stream, averaged over the timed steps
        djk_ms = stats["dijkstra_kernel_ms"]
        achieved = alg_bytes / world / (djk_ms * 1e-3) / 1e9 if djk_ms > 0 else 0.0
        key = (args.workload, float(info["scale"]))
        traffic = NCU_DRAM_BYTES.get(key) if world == 1 else None
        sectors = NCU_L2_READ_SECTORS.get(key) if world == 1 else None
        sector_gbps = sectors * 32 / (djk_ms * 1e-3) / 1e9 if sectors and djk_ms > 0 else None
        # cpu baseline: the oracle at 1 thread (deterministic reference semantics), whole workload, once; its outputs are
        # also what the GPU arm's bytes are compared with
        o, cpu_s, cpu_ph = oracle_pass(text, k, bcalm, 1)
        if world == 1:
            identical = (gfa == o.text("gfa")) and (bv == o.text("bitvector"))
        else:
            import hashlib
            ref = (o.text("gfa"), o.text("bitvector"))
            identical = True
            for which in (0, 1):
                covered = 0
                for p in sorted(parts, key=lambda p: p[which][0]):
                    off, n, digest, tot = p[which]
                    identical &= off == covered and tot == len(ref[which]) and hashlib.sha256(ref[which][off:off + n]).hexdigest() == digest
                    covered += n
                identical &= covered == len(ref[which])
        ref_settled, ref_dj_s = o.num("settled"), o.time("dijkstra")
        del o
        # T_compute: graph build from parsed records .. bitvector; T_wall adds parse, H2D/D2H and the GFA
        ph = stats["phases_ms"]
        t_compute = ms_res - ph.get("(parse, device)", 0.0) - ph.get("emit_gfa", 0.0)
        line = {
            "metric": "greedy_matchtig_unitigs_per_sec", "value": U / (ms_res * 1e-3), "unit": "unitigs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_of(args, world),
            "workload_sizes": {"unitigs": U, "nodes": gi["nodes"], "sources": gi["sources"], "short_edges": gi["short_edges"],
                               "distinct_kmers": info["distinct_kmers"], "input_bp": info["input_bp"], "text_bytes": len(text),
                               "generate_s": info["generate_s"], "cache": info.get("cache")},
            "t_compute_ms": t_compute, "t_wall_ms": ms_e2e,
            "unitigs_per_sec_compute": U / (t_compute * 1e-3), "unitigs_per_sec_wall": U / (ms_e2e * 1e-3),
            "ms_per_step_spread_rank0": stats["spread"],
            "phases_ms_rank0": stats["phases_ms"],  # host wall time per library call (each one blocks), mean per step
            "tail_ms_rank0": stats["tail_ms"],
            "e2e": {"value": U / (ms_e2e * 1e-3), "unit": "unitigs/s", "ms_per_step": ms_e2e, "ms_per_step_spread_rank0": stats_e2e["spread"],
                    "phases_ms_rank0": stats_e2e["phases_ms"], "tail_ms_rank0": stats_e2e["tail_ms"],
                    "h2d_bytes_per_step": int(text_host.numel()), "input": "unitig file content (parsed on the device)",
                    "host_link": link,
                    "d2h_bytes_per_step": int(out_bytes),
                    "sharded": None if world == 1 else "every rank uploads 1/N of the file (all-gather over NVLink) and downloads "
                                                       "the output bytes of 1/N of the walks"},
            "gpu_launches": int(launches),
            "settled_nodes": {"gpu_kernel": settled, "reference_semantics": ref_settled,
                              "note": "the GPU searches every source to the candidate cap up front; the reference stops each "
                                      "search after m+1 targets and skips satisfied sources (SURVEY.md 8d)"},
            "settled_nodes_per_sec": settled / (dj_ms * 1e-3) if dj_ms > 0 else None,
            "settled_nodes_per_sec_cpu_1thread": ref_settled / ref_dj_s if ref_dj_s > 0 else None,
            "dijkstra": {"ms_per_step": dj_ms, "settled_nodes": settled, "relaxed_edges": relaxed, "candidates": cands,
                         "sources_searched": searched, "kernel_ms_per_step": djk_ms, "match_ms_per_step": match_ms,
                         "match_kernel_ms_per_step": stats["match_kernel_ms"], "match_blocked_retries": stats["match_rounds"],
                         "requery_phases": stats["requery_phases"], "overflow_sources": stats["overflow_sources"],
                         "truncated_sources": stats["truncated_sources"], "preextended_sources": stats["preextended_sources"]},
            "roofline": {"kernel": SEARCH_KERNEL, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / world,
                         "sector_granular": {"l2_read_sectors_per_launch": sectors, "achieved_GBps": sector_gbps,
                                             "ceilings_GBps": GATHER_CEILING_GBPS,
                                             "frac_of_l2_resident_ceiling": sector_gbps / GATHER_CEILING_GBPS["l2_resident"]
                                             if sector_gbps else None,
                                             "source": "ncu lts__t_sectors_srcunit_tex_op_read.sum of the committed capture x 32 B over "
                                                       "the live kernel time; ceilings from scripts/micro/gather_ceiling.cu"},
                         "note": "random 32-B-sector gathers along dependent chains (see DESIGN.md section 4)"},
            "cpu_baseline": {"value": U / cpu_s, "unit": "unitigs/s", "cores": 1, "kind": "port",
                             "sample": f"whole workload, one pass (parse -> {OUTPUTS}); C++ restatement of matchtigs 2.1.9 greedy at "
                                       "--threads 1 (not the Rust binary)", "seconds": cpu_s,
                             "t_compute_ms": 1e3 * sum(cpu_ph[n] for n in ORACLE_COMPUTE), "phases_ms": {n: 1e3 * v for n, v in cpu_ph.items()}},
            "byte_identical_to_oracle": bool(identical),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


SEARCH_KERNEL = "dijkstra_thread_kernel"
# (workload, scale) -> DRAM bytes (read + write) of one main launch of dijkstra_thread_kernel, from ncu --set full
NCU_DRAM_BYTES = {("chr1", 1.0): 99224576 + 119722496}  # profiles/round2_ncu_full_dijkstra_thread_kernel_chr1.txt
# same captures: L2 sectors the kernel read (lts__t_sectors_srcunit_tex_op_read.sum), i.e. the sector-granular traffic
NCU_L2_READ_SECTORS = {("chr1", 1.0): 109156110}
# random 32-byte-sector gather ceilings measured on a B200 of this pool with scripts/micro/gather_ceiling.cu
# (profiles/round1_gather_ceiling.jsonl): footprint 8 GiB (HBM) and 32 MiB (L2-resident), independent gathers
GATHER_CEILING_GBPS = {"hbm_random": 1174.8, "l2_resident": 6663.9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=None)
    ap.add_argument("--ref-scale", type=float, default=None, help="--impl reference: recipe scale of the timed sample")
    ap.add_argument("--profile", action="store_true", help="profiling runs (ncu): one leg, no CPU pass, no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
