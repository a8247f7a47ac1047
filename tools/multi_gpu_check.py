"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank builds its tiles from the same global edge list, runs PR / BFS / CC / SSSP through the C ABI
with NCCL exchanges along the reference's row / column groups, and compares its owned segment with the
CPU oracle simulating the same p (bit-exact for the integer apps, 1e-6 relative for PageRank)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtap_b200 import engine as E  # noqa: E402
from graphtap_b200.rmat import rmat_edges  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    E.Env.init()
    rank, p = E.Env.rank, E.Env.nranks
    ok = True
    cases = [("fixture", np.fromfile(os.path.join(ROOT, "tests/golden/rmat10_1024_w.bin"), dtype="<u4").reshape(-1, 3), 1024)]
    for scale in (13, 16):
        cases.append((f"rmat{scale}", rmat_edges(scale, seed=scale + 1, weighted=True), 1 << scale))
    for name, tw, n in cases:
        tu = tw[:, :2].copy()

        def loader(tri):
            def load(G, **fl):
                ct = fl.pop("compression_type")
                G.load_triples(tri, n, compression_type=ct, **fl)
            return load

        for app, layouts in (("pr", (1, 0)), ("bfs", (None,)), ("cc", (None,)), ("sssp", (None,))):
            tri = tw if app == "sssp" else tu
            ref, rit = O.run_app(app, tri, n, p, 20 if app == "pr" else (None if app == "cc" else 0))
            for layout in layouts:
                if app == "pr":
                    G, V = E.run_pr(loader(tri), 20, pr_layout=layout)
                elif app == "bfs":
                    G, V = E.run_bfs(loader(tri), 0)
                elif app == "cc":
                    G, V = E.run_cc(loader(tri))
                else:
                    G, V = E.run_sssp(loader(tri), 0)
                mine = V.V
                lay = G.info().layout
                th, seg = lay.tile_height, lay.owned_segment
                mref = ref[seg * th:(seg + 1) * th]
                cs = V.checksum(quiet=True)
                good = V.iteration == rit
                if app == "pr":
                    rel = np.abs(mine["rank"] - mref["rank"]) / np.abs(mref["rank"])
                    good &= bool(rel.max() <= 1e-6) and bool((mine["degree"] == mref["degree"]).all())
                    detail = f"max rel {rel.max():.2e}"
                else:
                    for f in mine.dtype.names:
                        r = mref[f] if mref.dtype.names else mref
                        good &= bool((mine[f] == r).all())
                    detail = "bit-exact" if good else "MISMATCH"
                good &= cs == O.checksum(app, ref, n + 1)
                V.free(); G.free()
                print(f"[rank {rank}/{p}] {name} {app}{'' if layout is None else ' layout ' + str(layout)}: iters {V.iteration}/{rit} {detail} checksum {cs} -> {'OK' if good else 'FAIL'}", flush=True)
                ok &= good
    # ---- windows across calls: execute() twice on one program, and the three phases driven one by one ----
    tw = rmat_edges(13, seed=14, weighted=True)
    n = 1 << 13

    def load_w(G, **fl):
        ct = fl.pop("compression_type")
        G.load_triples(tw, n, compression_type=ct, **fl)

    def load_u(G, **fl):
        ct = fl.pop("compression_type")
        G.load_triples(tw[:, :2].copy(), n, compression_type=ct, **fl)

    ref, rit = O.run_app("sssp", tw, n, p, 0)
    G = E.Graph(weighted=True)
    load_w(G, directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    V = E.SSSP_Program(G, False, True, False, E._ROW_)
    V.root = 0
    V.execute(2)                      # two fixed iterations ...
    V.execute(4)                      # ... two more (iteration counts are absolute, as in the reference) ...
    it = V.execute()                  # ... then until convergence, all on the same windows
    lay = G.info().layout
    mine = V.V["distance"]
    good = it == rit and bool((mine == ref[lay.owned_segment * lay.tile_height:(lay.owned_segment + 1) * lay.tile_height]).all())
    print(f"[rank {rank}/{p}] sssp execute(2)+execute(4)+execute(): iters {it}/{rit} -> {'OK' if good else 'FAIL'}", flush=True)
    ok &= good
    V.free(); G.free()
    G = E.Graph(weighted=False)
    load_u(G, directed=False, transpose=False, self_loops=True, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    A = E.CC_Program(G, False, True, False, E._ROW_)
    A.execute(3)
    B = E.CC_Program(G, False, True, False, E._ROW_)
    for _ in range(3):                # CC's applicator does not depend on the iteration number: 3 rounds of phases == execute(3)
        for ph in (0, 1, 2):
            B.run_phase(ph)
    good = bool((A.V["label"] == B.V["label"]).all())
    B.execute()                       # and the program is still usable afterwards
    refc, _ = O.run_app("cc", tw[:, :2].copy(), n, p, None)
    good &= bool((B.V["label"] == refc[lay.owned_segment * lay.tile_height:(lay.owned_segment + 1) * lay.tile_height]).all())
    print(f"[rank {rank}/{p}] cc run_phase x3 == execute(3), then execute(): {'OK' if good else 'FAIL'}", flush=True)
    ok &= good
    A.free(); B.free(); G.free()
    E.Env.barrier()
    print(f"[rank {rank}] MULTI_GPU_CHECK {'PASS' if ok else 'FAIL'}", flush=True)
    E.Env.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
