import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from graphtap_b200 import engine as E
from oracle import oracle as O
E.Env.init(); E.Env.quiet = True
g = np.load(os.path.join(ROOT, "tests/golden/fixture.npz"))
tu = np.fromfile(os.path.join(ROOT, "tests/golden/rmat10_1024.bin"), dtype="<u4").reshape(-1, 2)
def loader(G, **fl):
    ct = fl.pop("compression_type"); G.load_triples(tu, 1024, compression_type=ct, **fl)
fl = dict(O.APP_FLAGS["pr"]); fl.pop("weighted")
og = O.OracleGraph(tu, 1024, 1, weighted=0, **fl); cls = og.classify(0)
ref = g["pr_np1_V"]
for comp in (E._TCSC_, E._TCSC_CF_):
    for layout in (1, 0):
        for iters in (1, 2, 20):
            G, P = E.run_pr(loader, iters, compression=comp, pr_layout=layout)
            V = P.V; cs = P.checksum(quiet=True)
            r, _ = og.pagerank(iters)
            err = np.abs(V["rank"][:1025] - r["rank"][:1025]) / r["rank"][:1025]
            bad = np.nonzero(err > 1e-9)[0]
            print(f"comp {comp} layout {layout} iters {iters}: cs {cs} maxerr {err.max():.3e} nbad {len(bad)} by class {[int((cls[bad]==c).sum()) for c in range(4)]} degbad {(V['degree'][:1025]!=r['degree'][:1025]).sum()}")
            if len(bad): print("   first bad", bad[:8], V["rank"][bad[:8]], r["rank"][bad[:8]])
            P.free(); G.free()
G = E.Graph(); loader(G, directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True, compression_type=E._TCSC_CF_)
t = G.tile(0); base = "pr_np1_tile_r0.t0"
print("IA equal", (t["IA"] == g[base + "_IA"]).all(), "JA", (t["JA"] == g[base + "_JA"]).all())
G.free()
