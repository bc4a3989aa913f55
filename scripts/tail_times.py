"""Times the host-sequential tail (no GPU) on a named workload, feeding it the oracle's graph + triples."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import matchtigs_b200 as mt, oracle, tools
name = sys.argv[1] if len(sys.argv) > 1 else "pangenome"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
text, k, info = tools.cached_config_unitigs(name, scale)
o = oracle.Oracle(euler_fast=True); o.load_fasta(text, k); o.run()
U = o.num("unitigs")
args = (o.array("edge_from")[:2*U], o.array("edge_to")[:2*U], o.array("edge_weight")[:2*U:2].astype(np.uint32), o.array("mirror"), o.array("triples"))
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    t = time.perf_counter(); walks, dw, ms = mt.api.host_tail(k, *args); dt = time.perf_counter() - t
    print(f"U={U} total={dt*1e3:.1f}ms phases(degrees,eulerise,csr,walk,break)={[round(x,2) for x in ms]} oracle_euler={o.time('euler')*1e3:.1f}ms")
ow = o.walks(); assert len(ow) == len(walks) and all(np.array_equal(a, b) for a, b in zip(walks, ow))
