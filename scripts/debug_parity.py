"""Stage-by-stage comparison GPU vs oracle on a named workload; prints the first divergence."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import matchtigs_b200 as mt, oracle, tools

name = sys.argv[1]; scale = float(sys.argv[2]); cap = int(sys.argv[3]) if len(sys.argv) > 3 else 16
text, k, info = tools.config_unitigs(name, scale)
print(info)
o = oracle.Oracle(euler_fast=True); o.load_fasta(text, k); o.run()
ctx = mt.Context(0)
g = mt.read_bigraph_from_fasta_as_edge_centric(text, k, ctx)
U = o.num("unitigs")
ex = ctx.graph_export()
for nm, ref in (("edge_from", o.array("edge_from")[:2*U]), ("edge_to", o.array("edge_to")[:2*U]), ("mirror", o.array("mirror")),
                ("sources", o.array("out_nodes"))):
    print(nm, "equal" if np.array_equal(ex[nm], ref) else "DIFFERENT")
print("imbalance", "equal" if np.array_equal(ex["imbalance"].astype(np.int64), o.array("mult0")) else "DIFFERENT")
ctx.dijkstra_candidates(cap)
nodes, dists, meta = ctx.candidates_export()
on, od, ol = o.candidates(cap)
cnt = (meta & 0xFFFFFF).astype(np.int64)
bad = np.flatnonzero(cnt != np.minimum(ol, cap))
print("candidate count mismatches:", len(bad), bad[:10])
mask = np.arange(cap)[None, :] < cnt[:, None]
nb = np.flatnonzero(((nodes != on) & mask).any(axis=1) | ((dists != od) & mask).any(axis=1))
print("candidate content mismatches:", len(nb), nb[:10])
for i in nb[:3]:
    print(" src idx", i, "node", ex["sources"][i], "gpu", list(zip(nodes[i][:cnt[i]], dists[i][:cnt[i]])), "oracle", list(zip(on[i][:min(ol[i],cap)], od[i][:min(ol[i],cap)])), "meta", hex(meta[i]))
trunc = (meta & 0x80000000) != 0
print("missing trunc flags:", int(np.sum(~trunc & (ol > cap))))
tr = ctx.greedy_match().reshape(-1)
ot = o.array("triples")
print("triples gpu", len(tr)//3, "oracle", len(ot)//3, ctx.search_stats())
m = min(len(tr), len(ot))
d = np.flatnonzero(tr[:m] != ot[:m])
if len(d) or len(tr) != len(ot):
    j = (d[0] // 3) if len(d) else m // 3
    print("first differing triple index", j, "gpu", tr[3*j-3:3*j+6], "oracle", ot[3*j-3:3*j+6])
    src = ot[3*j]; si = int(np.searchsorted(ex["sources"], src))
    print(" oracle source idx", si, "list gpu", list(zip(nodes[si][:cnt[si]], dists[si][:cnt[si]])), "meta", hex(meta[si]), "full len", ol[si])
else:
    print("triples equal")
