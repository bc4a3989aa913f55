"""A/B timings of the host tail on one workload: the graph, the searches and the matching run once, mtg_finish_walks is
repeated under different environment switches.  usage: tail_ab.py <workload> <scale> [reps]"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import matchtigs_b200 as mt  # noqa: E402

name, scale = sys.argv[1], float(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
text, k, info = bench.make_workload(name, scale)
ctx = mt.Context(0)
ctx.build_graph_from_text(text, k, bcalm=bench.WORKLOADS[name]["bcalm"])
ctx.dijkstra_candidates(bench.CAP, 0, 1)
ctx.greedy_match()
print(ctx.graph_info())
ctx.finish_walks()
print('THP:', open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip(), '| defrag:', open('/sys/kernel/mm/transparent_hugepage/defrag').read().strip(), '|', [l.strip() for l in open('/proc/self/smaps_rollup') if 'AnonHuge' in l or l.startswith('Rss')])
VARIANTS = (("default", {}), ("memcpy", {"MTG_TAIL_COPY": "memcpy"}), ("spin0", {"MTG_WALK_SPIN": "0"}), ("spin10", {"MTG_WALK_SPIN": "10"}),
            ("spin20", {"MTG_WALK_SPIN": "20"}), ("spin40", {"MTG_WALK_SPIN": "40"}), ("spin80", {"MTG_WALK_SPIN": "80"}), ("sources1", {"MTG_WALK_SOURCES": "1"}), ("fast0", {"MTG_WALK_FAST": "0"}), ("store", {"MTG_WALK_NTSTORE": "0"}),
            ("t0", {"MTG_WALK_PREFETCH": "t0"}), ("nohint", {"MTG_TAIL_NOHINT": "1"}), ("default", {}), ("probe", {"MTG_WALK_PROBE": "1"}),
            ("carried", {"MTG_WALK_CHAIN": "carried"}), ("lean2", {}), ("carried2", {"MTG_WALK_CHAIN": "carried"}), ("lean3", {}),
            # perfect-lookahead replay of the finished walk: mode (1: address depends on the record in hand, 2: and on the far used
            # word, 4: no used-bit work, 8: no queue appends) : depth
            ("replay", {"MTG_WALK_REPLAY": os.environ.get("TAIL_AB_REPLAY", ",".join(f"{m}:{d}" for m in (0, 1, 3) for d in (3, 4, 5, 6, 8, 12)) +
                                                          ",4:4,4:8,5:4,5:8,8:4,8:8,9:4,9:8,12:4,12:8,13:4,13:8")}))
only = os.environ.get("TAIL_AB_ONLY")  # comma-separated labels
for label, env in VARIANTS:
    if only and label not in only.split(","):
        continue
    for k_, v in env.items():
        os.environ[k_] = v
    rows = []
    for _ in range(reps):
        t0 = time.perf_counter()
        nw, ne = ctx.finish_walks()
        dt = 1e3 * (time.perf_counter() - t0)
        rows.append((dt, ctx.diagnostics()["tail_ms"]))
    rows.sort(key=lambda r: r[0])
    dt, ms = rows[len(rows) // 2]
    steps = ne + nw
    print(f"{label:14s} tail {dt:7.1f} ms  {ms}  walks {nw}  ~{1e6 * ms['euler_walk'] / max(steps, 1):.1f} ns/step", flush=True)
    for k_ in env:
        del os.environ[k_]
