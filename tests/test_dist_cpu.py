"""The N > 1 exchange plan on CPU: world_size 2 and 4 over gloo.

Each process is one rank.  It takes the 2DT layout from the PRODUCT library (gt_layout_*: tile owners,
local segments, leaders, row/column group lists — the tables gt_engine.cu drives NCCL with), owns the
tiles that table assigns to it, and runs the engine's schedule with gloo in place of NCCL:
messenger on the owned segment -> broadcast of every local x segment from its leader along the column
group -> SpMV on the owned tiles -> reduce of every local y segment to its leader along the row group ->
applicator at the leader -> allreduce for convergence.  The per-tile arithmetic comes from the CPU oracle
(this is a test of the plan, not of the kernels).  The gathered result must equal the oracle's
single-process simulation of the same p and the unmodified reference's dump (tests/golden)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _table(capi, n, p, r, which):
    cnt = C.c_uint32()
    capi.check(capi.lib().gt_layout_table(n, p, r, which, None, 0, C.byref(cnt)))
    out = (C.c_int32 * max(1, cnt.value))()
    capi.check(capi.lib().gt_layout_table(n, p, r, which, out, cnt.value, C.byref(cnt)))
    return list(out[: cnt.value])


def _worker(rank, world, port, app, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from graphtap_b200 import capi
    from oracle import oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1024
    weighted = app == "sssp"
    tri = np.fromfile(os.path.join(GOLDEN, "rmat10_1024_w.bin" if weighted else "rmat10_1024.bin"), dtype="<u4").reshape(-1, 3 if weighted else 2)
    fl = dict(O.APP_FLAGS[app]); w = fl.pop("weighted")
    g = O.OracleGraph(tri, n, world, weighted=w, **fl)
    th = g.th

    lay = capi.Layout()
    capi.check(capi.lib().gt_layout_query(n, world, rank, C.byref(lay)))
    tile_rank = np.array(_table(capi, n, world, rank, capi.GT_LT_TILE_RANK)).reshape(world, world)
    leaders = _table(capi, n, world, rank, capi.GT_LT_LEADER_RANKS)
    row_segs = _table(capi, n, world, rank, capi.GT_LT_LOCAL_ROW_SEGMENTS)
    col_segs = _table(capi, n, world, rank, capi.GT_LT_LOCAL_COL_SEGMENTS)
    # every process must create every group, in the same order (as ncclCommSplit does collectively)
    groups = {}
    for r in range(world):
        for which in (capi.GT_LT_ALL_ROWGRP_RANKS, capi.GT_LT_ALL_COLGRP_RANKS):
            key = tuple(_table(capi, n, world, r, which))
            if key not in groups:
                groups[key] = dist.new_group(list(key))
    rowgrp = groups[tuple(_table(capi, n, world, rank, capi.GT_LT_ALL_ROWGRP_RANKS))]
    colgrp = groups[tuple(_table(capi, n, world, rank, capi.GT_LT_ALL_COLGRP_RANKS))]
    own = lay.owned_segment
    assert leaders[own] == rank and own in row_segs and own in col_segs

    base = own * th
    vid = np.arange(base, base + th, dtype=np.uint32)
    rows_own, cols_own = g.seg(False, own), g.seg(True, own)
    INF = O.INF
    if app == "pr":
        deg_all = g.degree(1)                                   # Deg pass (pr.cpp:40-43); exchange-free in the oracle
        deg = deg_all[base:base + th].copy()
        deg[rows_own["bits"] == 0] = 0                          # initialize(other) only where the row is non-empty
        rank_v = np.full(th, 0.15)
        ftype, red = torch.float64, dist.ReduceOp.SUM
    else:
        a = vid.copy() if app == "cc" else np.where(vid == 0, 0, INF).astype(np.uint32)
        hops = np.where(vid == 0, 0, INF).astype(np.uint32)
        parent = np.where(vid == 0, vid, 0).astype(np.uint32)
        Cflag = np.ones(th, dtype=bool) if app == "cc" else (vid == 0)
        ftype, red = torch.int64, dist.ReduceOp.MIN             # gloo has no uint32 MIN; int64 holds the u32 range
    X = {s: None for s in col_segs}
    Y = {s: (np.zeros(g.seg(False, s)["nnz"]) if app == "pr" else np.full(g.seg(False, s)["nnz"], INF, dtype=np.uint32)) for s in row_segs}
    it = 0
    while True:
        # scatter_gather: messenger on the owned segment, then one broadcast per local column segment
        ids = cols_own["ids"]
        if app == "pr":
            d = deg[ids].astype(np.float64)
            xo = np.divide(rank_v[ids], d, out=np.zeros(len(ids)), where=d > 0)
        else:
            src = (vid[ids] if app == "bfs" else a[ids])
            xo = np.where(Cflag[ids], src, INF).astype(np.uint32)
        for s in col_segs:
            nc = g.seg(True, s)["nnz"]
            buf = torch.from_numpy((xo if s == own else np.zeros(nc, dtype=xo.dtype)).astype(np.float64 if app == "pr" else np.int64))
            if nc:
                dist.broadcast(buf, src=leaders[s], group=colgrp)
            X[s] = buf.numpy().astype(np.float64 if app == "pr" else np.uint32)
        # combine: SpMV on the tiles this rank owns, then reduce each local row segment to its leader
        for s in row_segs:
            if app == "pr":
                Y[s][:] = 0.0
        for rg in row_segs:
            for cg in col_segs:
                if tile_rank[rg, cg] != rank or g.tile(rg, cg)["nnz"] == 0:
                    continue
                if app == "pr":
                    g.spmv_f64(rg, cg, X[cg], Y[rg], 0)
                else:
                    g.spmv_u32(rg, cg, X[cg], Y[rg])
        for s in row_segs:
            buf = torch.from_numpy(Y[s].astype(np.float64 if app == "pr" else np.int64))
            if len(buf):
                dist.reduce(buf, dst=leaders[s], op=red, group=rowgrp)
            if s == own:
                y = buf.numpy().astype(np.float64 if app == "pr" else np.uint32)
                if app != "pr":
                    Y[s][:] = y                              # the leader keeps the running minimum
        # apply on the owned segment
        rid = rows_own["ids"]
        if app == "pr":
            new = 0.15 + (1.0 - 0.15) * y
            rank_v[rid] = new
            active = 0
        else:
            if it == 0:
                Cflag[rows_own["bits"] == 0] = False
            if app == "bfs":
                ch = (hops[rid] == INF) & (y != INF)
                hops[rid] = np.where(ch, it + 1, hops[rid]); parent[rid] = np.where(ch, y, parent[rid])
            elif app == "cc":
                ch = y < a[rid]
                a[rid] = np.where(ch, y, a[rid])
            else:
                ch = y < a[rid]
                a[rid] = np.where(ch, y, a[rid])
            Cflag[rid] = ch
            active = int(ch.sum())
        it += 1
        t = torch.tensor([active], dtype=torch.int64)
        dist.all_reduce(t)
        if (app == "pr" and it >= 20) or (app != "pr" and t.item() == 0):
            break
    if app == "pr":
        out = np.zeros(th, dtype=O.PR_STATE); out["rank"], out["degree"] = rank_v, deg
    elif app == "bfs":
        out = np.zeros(th, dtype=O.BFS_STATE); out["parent"], out["hops"], out["vid"] = parent, hops, vid
    else:
        out = a
    np.save(os.path.join(out_dir, f"seg{own}.npy"), out)
    with open(os.path.join(out_dir, f"it{rank}.txt"), "w") as f:
        f.write(str(it))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("app", ["pr", "bfs", "cc", "sssp"])
def test_exchange_plan_over_gloo(world, app, tmp_path, golden_fixture):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    port = 29600 + world * 10 + ["pr", "bfs", "cc", "sssp"].index(app)
    mp.spawn(_worker, args=(world, port, app, str(tmp_path)), nprocs=world, join=True)
    th = (1024 + 1) // world + 1
    got = np.concatenate([np.load(tmp_path / f"seg{s}.npy") for s in range(world)])
    ref = golden_fixture[f"{app}_np{world}_V"]
    assert int(open(tmp_path / "it0.txt").read()) == golden_fixture[f"{app}_np{world}_meta"][0]
    n = 1025
    if app == "pr":
        assert (got["degree"][:n] == ref["degree"][:n]).all()
        np.testing.assert_allclose(got["rank"][:n], ref["rank"][:n], rtol=1e-12)
    elif app == "bfs":
        for f in ("parent", "hops", "vid"):
            assert (got[f][:n] == ref[f][:n]).all()
    else:
        assert (got[:n] == ref[:n]).all()
    assert len(got) == world * th
