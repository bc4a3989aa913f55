"""Wall-clock split of one hot-path step into its phases (diagnostic, not a bench number)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import matchtigs_b200 as mt
import tools

name = sys.argv[1] if len(sys.argv) > 1 else "pangenome"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
text, k, info = tools.config_unitigs(name, scale)
u = mt.Unitigs(text, bcalm=False)
ctx = mt.Context(0)
print(info)
for it in range(4):
    t = [time.perf_counter()]
    def mark():
        torch.cuda.synchronize(); t.append(time.perf_counter())
    ctx.build_graph_from_sequences(u.seq, u.offsets, k); mark()
    ctx.dijkstra_candidates(int(sys.argv[3]) if len(sys.argv) > 3 else 8); mark()
    ctx.greedy_match(); mark()
    ctx.finish_walks(); mark()
    gfa = ctx.assemble_tigs("gfa"); mark()
    bv = ctx.dup_bitvector(); mark()
    names = ["build", "dijkstra", "match", "finish_walks", "gfa", "bitvector"]
    d = np.diff(t) * 1e3
    print("iter", it, " ".join(f"{n}={x:.2f}ms" for n, x in zip(names, d)), f"total={d.sum():.2f}ms", ctx.search_stats())
print(ctx.graph_info(), len(gfa), len(bv))
print(ctx.diagnostics())
print([l.strip() for l in open('/proc/self/smaps_rollup') if 'AnonHuge' in l or 'Rss' in l][:3])
