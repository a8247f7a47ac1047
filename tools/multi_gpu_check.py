"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank builds its tiles from the same global edge list (and, in the last block, from its 1/p share of it:
partitioned ingest must give the bit-identical graph), runs PR / BFS / CC / SSSP through the C ABI
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
    # ---- partitioned ingest (gt_graph_build_partitioned): every rank passes 1/p of the records; the graph must come out
    # bit-identical to the build where every rank scans the whole list (tiles, maps, CF lists, hot-order-dependent results) ----
    def graph_fingerprint(G):
        out, tiles = [], [G.tile(k) for k in range(G.info().ntiles_local)]
        for t in tiles:
            out += [t[f] for f in ("JA", "IA", "JC", "IR")] + ([t["A"]] if t["A"] is not None else [])
        for s_ in sorted({t["row_slot"] for t in tiles}):
            out += list(G.rowgrp_maps(s_)[:2])
        for s_ in sorted({t["col_slot"] for t in tiles}):
            out += list(G.colgrp_maps(s_)[:2])
        return out

    import tempfile
    import zlib
    for name, tri, n, weighted, fl in (
            ("fixture directed+transpose (pr)", np.fromfile(os.path.join(ROOT, "tests/golden/rmat10_1024.bin"), dtype="<u4").reshape(-1, 2), 1024, False,
             dict(directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True, compression_type=E._TCSC_CF_)),
            ("rmat14 undirected dedup (cc)", rmat_edges(14, seed=15, weighted=False), 1 << 14, False,
             dict(directed=False, transpose=False, self_loops=True, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)),
            ("rmat14 weighted transpose dedup (sssp)", rmat_edges(14, seed=15, weighted=True), 1 << 14, True,
             dict(directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_))):
        Gg = E.Graph(weighted=weighted).load_triples(tri, n, **fl)
        # (a) an uneven in-memory split, (b) the reference's split of a binary file, read by byte range
        cut = [0] + sorted(int(x) for x in np.random.default_rng(7).integers(0, tri.shape[0] + 1, p - 1)) + [tri.shape[0]]
        Gp = E.Graph(weighted=weighted).load_triples(tri[cut[rank]:cut[rank + 1]], n, partitioned=True, **fl)
        path = os.path.join(tempfile.gettempdir(), f"gt_check_{os.environ.get('MASTER_PORT', '0')}_{zlib.crc32(name.encode())}.bin")
        if rank == 0:
            tri.astype("<u4").tofile(path)
        E.Env.barrier()
        Gf = E.Graph(weighted=weighted).load(path, n, n, **fl)           # nranks > 1: partitioned by default
        a, b, c = graph_fingerprint(Gg), graph_fingerprint(Gp), graph_fingerprint(Gf)
        good = len(a) == len(b) == len(c) and all(x.shape == y.shape == z.shape and (x == y).all() and (x == z).all() for x, y, z in zip(a, b, c))
        good &= Gg.info().nnz_global == Gp.info().nnz_global == Gf.info().nnz_global and Gp.info().nedges_input == tri.shape[0] == Gf.info().nedges_input
        if fl["compression_type"] == E._TCSC_CF_:
            for k in range(Gg.info().ntiles_local):
                x, y = Gg.tile_cf(k), Gp.tile_cf(k)
                good &= all(x[f"{f}{i}"].shape == y[f"{f}{i}"].shape and (x[f"{f}{i}"] == y[f"{f}{i}"]).all() for f in ("JA", "JC") for i in range(4))
        E.Env.barrier()
        if rank == 0:
            os.unlink(path)
        print(f"[rank {rank}/{p}] partitioned ingest, {name}: {len(a)} arrays, nnz {Gp.info().nnz_local}/{Gp.info().nnz_global} -> {'OK' if good else 'FAIL'}", flush=True)
        ok &= good
        Gg.free(); Gp.free(); Gf.free()
    # generated graph, each rank generating 1/p of the records: same PageRank as the oracle
    def load_rp(G, **fl):
        G.load_rmat(13, seed=14, partitioned=True, **fl)
    G, V = E.run_pr(load_rp, 20)
    ref, _ = O.run_app("pr", rmat_edges(13, seed=14, weighted=False), 1 << 13, p, 20)
    lay = G.info().layout
    mref = ref[lay.owned_segment * lay.tile_height:(lay.owned_segment + 1) * lay.tile_height]
    rel = np.abs(V.V["rank"] - mref["rank"]) / np.abs(mref["rank"])
    good = bool(rel.max() <= 1e-6)
    print(f"[rank {rank}/{p}] partitioned rmat13 pr: max rel {rel.max():.2e} -> {'OK' if good else 'FAIL'}", flush=True)
    ok &= good
    V.free(); G.free()
    E.Env.barrier()
    print(f"[rank {rank}] MULTI_GPU_CHECK {'PASS' if ok else 'FAIL'}", flush=True)
    E.Env.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
