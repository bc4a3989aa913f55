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

import numpy as np  # noqa: E402

WORKLOADS = {
    # name -> (tools.config_unitigs name, default scale, description)
    "ecoli": ("ecoli", 1.0, "config[1]: synthetic 4.6 Mbp E. coli-sized genome with injected repeats, k=31, "
                            "unitigs -> greedy matchtigs GFA + duplicate-kmer bitvector"),
    "pangenome": ("pangenome", 0.25, "config[3] (scaled): synthetic pangenome, 100 strains with SNP/indel variation, k=31"),
    "chr1": ("chr1", 0.1, "config[2] (scaled): synthetic chr1-like genome with repeat families, k=31"),
    "human": ("human", 0.01, "config[4] (scaled): synthetic human-like genome, k=51"),
}
CAP = 16


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


def make_workload(name: str, scale: float | None):
    import tools
    cfg, dscale, desc = WORKLOADS[name]
    scale = dscale if scale is None else scale
    t0 = time.time()
    text, k, info = tools.config_unitigs(cfg, scale)
    info["generate_s"] = round(time.time() - t0, 2)
    info["description"] = desc
    return text, k, info


# ----------------------------------------------------------------------------------------------
# reference arm: the CPU oracle with every host thread (C++ restatement, not the Rust binary)
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    import oracle
    text, k, info = make_workload(args.workload, args.scale)
    cores = os.cpu_count() or 1
    U = info["unitigs"]
    times, settled = [], 0
    for it in range(args.warmup + args.steps):
        o = oracle.Oracle(euler_fast=True)
        t0 = time.perf_counter()
        o.load_fasta(text, k)
        o.run(threads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            settled = o.num("settled")
            dj = o.time("dijkstra")
    ms = 1e3 * float(np.mean(times))
    value = U / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "greedy_matchtig_unitigs_per_sec", "value": value, "unit": "unitigs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "k": k, **{k_: info[k_] for k_ in ("scale", "unitigs", "distinct_kmers", "input_bp")}},
        "cpu_baseline": {"value": value, "unit": "unitigs/s", "cores": cores, "kind": "port",
                         "sample": "whole workload per step; C++ restatement of matchtigs 2.1.9 greedy with the reference's "
                                   "worker scheme (not the Rust binary, which cannot be built in this image)"},
        "e2e": {"value": value, "unit": "unitigs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    text, k, info = make_workload(args.workload, args.scale)
    units = mt.Unitigs(text, bcalm=False)
    U = units.count
    seq_host = torch.from_numpy(units.seq.copy()).pin_memory()
    off_host = torch.from_numpy(units.offsets.astype(np.int64)).pin_memory()
    seq_dev, off_dev = seq_host.cuda(), off_host.cuda()
    text_host = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()  # the unitig FASTA file content
    ctx = mt.Context(local_rank)
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
            call("graph_build", ctx.build_graph_from_device_sequences, seq_dev.data_ptr(), off_dev.data_ptr(), U, k)
        else:
            # end to end: raw FASTA bytes in page-locked host memory -> H2D -> records parsed on the device -> graph
            call("parse+graph_build", ctx.build_graph_from_text, text_host.numpy(), k, bcalm=False)
        if world == 1:
            call("dijkstra", ctx.dijkstra_candidates, CAP, 0, 1)
            call("match", ctx.greedy_match)
        else:
            call("dijkstra", ctx.dijkstra_candidates, CAP, rank, world)
            rec_all, meta_all = call("all_gather", sharding.all_gather_candidates, ctx, dist, torch, world)  # NCCL over NVLink
            call("match", ctx.greedy_match, rec_all.data_ptr(), meta_all.data_ptr(), world)
        if rank == 0:
            call("tail", ctx.finish_walks)
            # results land in page-locked host memory owned by the context (zero-copy views)
            return call("emit_gfa", ctx.assemble_tigs_view, "gfa"), call("emit_bitvector", ctx.dup_bitvector_view)
        return None, None

    def timed(resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(resident)
        barrier()
        total_ms, dj_ms, mt_ms, djk_ms, mtk_ms, stats, out = 0.0, 0.0, 0.0, 0.0, 0.0, None, None
        per_step = []
        launches0 = ctx.kernel_launches
        phase_s.clear()
        for _ in range(steps):
            with torch.cuda.stream(ext):
                flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(ext)
            out = step(resident)
            e1.record(ext)
            barrier()
            per_step.append(e0.elapsed_time(e1))
            total_ms += per_step[-1]
            stats = ctx.search_stats()
            dj_ms += stats["dijkstra_ms"]
            mt_ms += stats["match_ms"]
            djk_ms += stats["dijkstra_kernel_ms"]
            mtk_ms += stats["match_kernel_ms"]
        launches = ctx.kernel_launches - launches0
        t = torch.tensor([total_ms, dj_ms, mt_ms, djk_ms, mtk_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, dj_ms, mt_ms, djk_ms, mtk_ms = (float(x) for x in t.cpu())
        per_step.sort()
        stats = dict(stats, dijkstra_kernel_ms=djk_ms / steps, match_kernel_ms=mtk_ms / steps,
                     spread={"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
                     phases_ms={n: 1e3 * v / steps for n, v in phase_s.items()})
        return total_ms / steps, dj_ms / steps, mt_ms / steps, stats, launches, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, dj_ms, match_ms, stats, launches, _ = timed(True, args.steps, args.warmup)
    ms_e2e, _, _, stats_e2e, _, out = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # settled / relaxed / candidates summed over ranks for the roofline numerator
    cnt = torch.tensor([stats["settled_nodes"], stats["relaxed_edges"], stats["candidates"], stats["sources_searched"]],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    settled, relaxed, cands, searched = (float(x) for x in cnt.cpu())

    if rank == 0:
        gi = ctx.graph_info()
        gfa, bv = (x.tobytes() for x in out)
        peak, peak_src = load_peaks()
        # algorithmic bytes of the Dijkstra kernel: 12 B per settled node (row_ptr pair + target probe),
        # 5 B per relaxed short edge (col + weight), 8 B per emitted candidate
        alg_bytes = 12.0 * settled + 5.0 * relaxed + 8.0 * cands
        # duration: CUDA events around the tier-0 search kernel on the library's stream, averaged over the timed steps
        djk_ms = stats["dijkstra_kernel_ms"]
        achieved = alg_bytes / world / (djk_ms * 1e-3) / 1e9 if djk_ms > 0 else 0.0
        # dram__bytes_read.sum + dram__bytes_write.sum of the kernel's main launch from the committed `ncu --set full`
        # captures (profiles/round1_ncu_full_dijkstra_thread_main_*.txt); only known for the captured workloads at N=1
        traffic = NCU_DRAM_BYTES.get((args.workload, float(info["scale"]))) if world == 1 else None
        sectors = NCU_L2_READ_SECTORS.get((args.workload, float(info["scale"]))) if world == 1 else None
        sector_gbps = sectors * 32 / (djk_ms * 1e-3) / 1e9 if sectors and djk_ms > 0 else None
        # cpu baseline: the oracle at 1 thread (deterministic reference semantics), whole workload, once
        import oracle
        o = oracle.Oracle(euler_fast=True)
        t0 = time.perf_counter()
        o.load_fasta(text, k)
        o.run(1)
        cpu_s = time.perf_counter() - t0
        identical = (gfa == o.text("gfa")) and (bv == o.text("bitvector"))
        line = {
            "metric": "greedy_matchtig_unitigs_per_sec", "value": U / (ms_res * 1e-3), "unit": "unitigs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload, "description": info["description"], "k": k, "scale": info["scale"],
                       "unitigs": U, "nodes": gi["nodes"], "sources": gi["sources"], "short_edges": gi["short_edges"],
                       "distinct_kmers": info["distinct_kmers"], "input_bp": info["input_bp"], "candidate_cap": CAP,
                       "l2": "flushed between timed iterations (256 MiB memset)", "reader": "fa-in semantics (k-mer join)",
                       "parallelism": f"sources sharded over {world} GPU(s), graph replicated"},
            "ms_per_step_spread_rank0": stats["spread"],
            "phases_ms_rank0": stats["phases_ms"],  # host wall time per library call (each one blocks), mean per step
            "e2e": {"value": U / (ms_e2e * 1e-3), "unit": "unitigs/s", "ms_per_step": ms_e2e, "ms_per_step_spread_rank0": stats_e2e["spread"],
                    "phases_ms_rank0": stats_e2e["phases_ms"],
                    "h2d_bytes_per_step": int(text_host.numel()), "input": "unitig FASTA text (parsed on the device)",
                    "host_link": link,
                    "d2h_bytes_per_step": int(len(gfa) + len(bv))},
            "gpu_launches": int(launches),
            "settled_nodes_per_sec": settled / (dj_ms * 1e-3) if dj_ms > 0 else None,
            "dijkstra": {"ms_per_step": dj_ms, "settled_nodes": settled, "relaxed_edges": relaxed, "candidates": cands,
                         "sources_searched": searched, "kernel_ms_per_step": djk_ms, "match_ms_per_step": match_ms,
                         "match_kernel_ms_per_step": stats["match_kernel_ms"], "match_blocked_retries": stats["match_rounds"],
                         "requery_phases": stats["requery_phases"], "overflow_sources": stats["overflow_sources"]},
            "roofline": {"kernel": "dijkstra_thread_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / world,
                         "sector_granular": {"l2_read_sectors_per_launch": sectors, "achieved_GBps": sector_gbps,
                                             "ceilings_GBps": GATHER_CEILING_GBPS,
                                             "frac_of_l2_resident_ceiling": sector_gbps / GATHER_CEILING_GBPS["l2_resident"]
                                             if sector_gbps else None,
                                             "source": "ncu lts__t_sectors_srcunit_tex_op_read.sum of the committed capture x 32 B over "
                                                       "the live kernel time; ceilings from scripts/micro/gather_ceiling.cu"},
                         "note": "random 32-B-sector gathers along dependent chains; the CSR of this workload fits in L2, so the "
                                 "kernel is bound by L2 latency x chain depth, not by HBM bandwidth (see DESIGN.md section 4)"},
            "cpu_baseline": {"value": U / cpu_s, "unit": "unitigs/s", "cores": 1, "kind": "port",
                             "sample": "whole workload, one run; C++ restatement of matchtigs 2.1.9 greedy at --threads 1 "
                                       "(not the Rust binary)", "seconds": cpu_s},
            "byte_identical_to_oracle": bool(identical),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# (workload, scale) -> DRAM bytes (read + write) of one main launch of dijkstra_thread_kernel, from ncu --set full
NCU_DRAM_BYTES = {("ecoli", 1.0): 208896 + 0, ("chr1", 0.3): 27804160 + 17134592}
# same captures: L2 sectors the kernel read (lts__t_sectors_srcunit_tex_op_read.sum), i.e. the sector-granular traffic
NCU_L2_READ_SECTORS = {("ecoli", 1.0): 144437, ("chr1", 0.3): 33546649}
# random 32-byte-sector gather ceilings measured on a B200 of this pool with scripts/micro/gather_ceiling.cu
# (profiles/round1_gather_ceiling.jsonl): footprint 8 GiB (HBM) and 32 MiB (L2-resident), independent gathers
GATHER_CEILING_GBPS = {"hbm_random": 1174.8, "l2_resident": 6663.9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ecoli", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
