"""Parity of the CUDA path (through the C ABI) with the reference: golden per-vertex dumps of the
unmodified reference, the CPU oracle on seeded inputs, tile arrays, single-kernel checks, edge cases
and size-independent properties.  Bit-exact for BFS/CC/SSSP/Deg and every index array; PageRank within
1e-6 relative per vertex (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PR_RTOL = 1e-6      # north_star: "within 1e-6 relative per-vertex error for PageRank ranks"


def _E():
    from graphtap_b200 import engine
    return engine


def _O():
    from oracle import oracle
    return oracle


def loader(tri, n):
    def load(G, **fl):
        ct = fl.pop("compression_type")
        G.load_triples(tri, n, compression_type=ct, **fl)
    return load


def run_gpu(app, tri, n, arg=None, compression=None):
    E = _E()
    if app == "pr":
        G, V = E.run_pr(loader(tri, n), 20 if arg is None else arg, **({"compression": compression} if compression else {}))
    elif app == "bfs":
        G, V = E.run_bfs(loader(tri, n), arg or 0)
    elif app == "cc":
        G, V = E.run_cc(loader(tri, n))
    else:
        G, V = E.run_sssp(loader(tri, n), arg or 0)
    states, it, cs, tm = V.V, V.iteration, V.checksum(quiet=True), V.timing()
    V.free(); G.free()
    return states, it, cs, tm


def assert_states(app, mine, ref, n):
    if app == "pr":
        assert (mine["degree"][:n] == ref["degree"][:n]).all()
        rel = np.abs(mine["rank"][:n] - ref["rank"][:n]) / np.abs(ref["rank"][:n])
        assert rel.max() <= PR_RTOL, rel.max()
    elif app == "bfs":
        for f in ("parent", "hops", "vid"):
            assert (mine[f][:n] == ref[f][:n]).all(), f
    else:
        m = mine[mine.dtype.names[0]] if mine.dtype.names else mine
        assert (m[:n] == ref[:n]).all()


@pytest.mark.parametrize("app", ["pr", "bfs", "cc", "sssp"])
def test_fixture_vs_reference_dump(app, fixture_unweighted, fixture_weighted, golden_fixture):
    tri = fixture_weighted if app == "sssp" else fixture_unweighted
    mine, it, cs, _ = run_gpu(app, tri, 1024)
    meta = golden_fixture[f"{app}_np1_meta"]
    assert it == meta[0]
    assert_states(app, mine, golden_fixture[f"{app}_np1_V"], 1025)
    assert cs == (meta[1], meta[2])            # Value checksum / Reachable vertices lines


def test_pr_tcsc_equals_tcsc_cf(fixture_unweighted, golden_fixture):
    E = _E()
    a, _, _, _ = run_gpu("pr", fixture_unweighted, 1024, 20, compression=E._TCSC_)
    b, _, _, _ = run_gpu("pr", fixture_unweighted, 1024, 20, compression=E._TCSC_CF_)
    np.testing.assert_allclose(a["rank"], b["rank"], rtol=1e-12)      # split rows meet through RED.ADD: order may differ


@pytest.mark.parametrize("layout", [0, 1])
def test_pr_push_and_pull_layouts(layout, rmat12, golden_rmat12, fixture_unweighted, golden_fixture):
    """Both SpMV forms of PageRank — push over the TCSC arrays (pr_layout 0) and the derived pull layout
    (pr_layout 1, the default) — against the reference's per-vertex ranks."""
    E = _E()
    for tri, n, gold in ((fixture_unweighted, 1024, golden_fixture), (rmat12[:, :2].copy(), 4096, golden_rmat12)):
        G, P = E.run_pr(loader(tri, n), 20, pr_layout=layout)
        mine = P.V
        P.free(); G.free()
        assert_states("pr", mine, gold["pr_np1_V"], n + 1)


def test_pr_pull_layout_long_rows_are_split():
    """A hub with more in-edges than one lane sums (kPullVRow = 2048) exercises the virtual-row path."""
    E, O = _E(), _O()
    n = 9000
    src = np.arange(1, n, dtype="<u4")
    tri = np.concatenate([np.stack([src, np.zeros_like(src)], axis=1),            # everyone -> 0   (hub row, degree 8999)
                          np.stack([np.zeros(40, dtype="<u4"), np.arange(1, 41, dtype="<u4")], axis=1),
                          np.stack([src[:-1], src[1:]], axis=1)])                  # a chain so ranks differ
    ref, _ = O.run_app("pr", tri, n, 1, 20)
    G, P = E.run_pr(loader(tri, n), 20, pr_layout=1)
    mine = P.V
    P.free(); G.free()
    assert_states("pr", mine, ref, n + 1)


@pytest.mark.parametrize("app", ["pr", "bfs", "cc", "sssp"])
def test_rmat12_vs_reference_dump(app, rmat12, golden_rmat12):
    tri = rmat12 if app == "sssp" else rmat12[:, :2].copy()
    mine, it, _, _ = run_gpu(app, tri, 4096)
    assert it == golden_rmat12[f"{app}_np1_meta"][0]
    assert_states(app, mine, golden_rmat12[f"{app}_np1_V"], 4097)


@pytest.mark.parametrize("scale,seed", [(14, 1), (16, 2), (18, 3)])
@pytest.mark.parametrize("app", ["pr", "bfs", "cc", "sssp"])
def test_seeded_rmat_vs_oracle(app, scale, seed):
    from graphtap_b200.rmat import rmat_edges
    O = _O()
    tri = rmat_edges(scale, seed=seed, weighted=(app == "sssp"))
    n = 1 << scale
    ref, rit = O.run_app(app, tri, n, 1, 20 if app == "pr" else (None if app == "cc" else 0))
    mine, it, cs, tm = run_gpu(app, tri, n)
    assert it == rit
    assert_states(app, mine, ref, n + 1)
    assert cs == O.checksum(app, ref, n + 1)
    if app != "pr":
        assert tm.sparse_iterations >= 1          # the frontier SpMSpV really ran


@pytest.mark.parametrize("root", [1, 77, 4095])
def test_roots(root, rmat12):
    O = _O()
    for app in ("bfs", "sssp"):
        tri = rmat12 if app == "sssp" else rmat12[:, :2].copy()
        ref, rit = O.run_app(app, tri, 4096, 1, root)
        mine, it, _, _ = run_gpu(app, tri, 4096, root)
        assert it == rit
        assert_states(app, mine, ref, 4097)


@pytest.mark.parametrize("app", ["pr", "bfs", "sssp"])
def test_tiles_vs_oracle(app, rmat12):
    """Device-built TCSC arrays and index maps, array by array (unweighted: bit-identical to
    TCSC_BASE::populate; weighted: same per-column {row: min weight})."""
    E, O = _E(), _O()
    tri = rmat12 if app == "sssp" else rmat12[:, :2].copy()
    fl = dict(O.APP_FLAGS[app]); w = fl.pop("weighted")
    og = O.OracleGraph(tri, 4096, 1, weighted=w, **fl)
    G = E.Graph(weighted=bool(w))
    G.load_triples(tri, 4096, compression_type=E._TCSC_, **{k: bool(v) for k, v in fl.items()})
    t, ot = G.tile(0), og.tile(0, 0)
    np.testing.assert_array_equal(t["JC"], og.seg(True, 0)["ids"])
    np.testing.assert_array_equal(t["IR"], og.seg(False, 0)["ids"])
    I, IV, nr = G.rowgrp_maps(0)
    J, JV, nc = G.colgrp_maps(0)
    np.testing.assert_array_equal(I, og.seg(False, 0)["bits"]); np.testing.assert_array_equal(IV, og.seg(False, 0)["prefix"])
    np.testing.assert_array_equal(J, og.seg(True, 0)["bits"]); np.testing.assert_array_equal(JV, og.seg(True, 0)["prefix"])
    assert t["nnz"] == G.info().nnz_local
    if not w:
        assert t["nnz"] == ot["nnz"]
        np.testing.assert_array_equal(t["JA"], ot["JA"])
        np.testing.assert_array_equal(t["IA"], ot["IA"])
    else:
        # weighted: the reference (and the oracle) drop only ADJACENT duplicates in (col, weight) order, the
        # device build keeps exactly the lightest copy of every (row, col) — same {row: min weight} per column
        assert t["nnz"] <= ot["nnz"]
        for j in range(0, nc, 3):
            mine = dict(zip(t["IA"][t["JA"][j]:t["JA"][j + 1]].tolist(), t["A"][t["JA"][j]:t["JA"][j + 1]].tolist()))
            assert len(mine) == t["JA"][j + 1] - t["JA"][j]                  # no duplicate rows left in a column
            ref = {}
            for r, wt in zip(ot["IA"][ot["JA"][j]:ot["JA"][j + 1]].tolist(), ot["A"][ot["JA"][j]:ot["JA"][j + 1]].tolist()):
                ref[r] = min(wt, ref.get(r, wt))
            assert mine == ref
    G.free(); og.close()


@pytest.mark.parametrize("semiring", ["plus_times", "min_select", "min_plus"])
def test_single_tile_kernels_vs_oracle(semiring, rmat12):
    """gt_tile_spmv (push and pull) and gt_tile_spmspv on random vectors, one launch each."""
    E, O = _E(), _O()
    from graphtap_b200 import capi
    rng = np.random.default_rng(5)
    w = semiring == "min_plus"
    tri = rmat12 if w else rmat12[:, :2].copy()
    fl = dict(directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=not w)
    og = O.OracleGraph(tri, 4096, 1, weighted=int(w), **{k: int(v) for k, v in fl.items()})
    G = E.Graph(weighted=w)
    G.load_triples(tri, 4096, compression_type=E._TCSC_, **fl)
    nc, nr = og.seg(True, 0)["nnz"], og.seg(False, 0)["nnz"]
    if semiring == "plus_times":
        x = rng.random(nc); y0 = rng.random(nr)
        ref = y0.copy(); og.spmv_f64(0, 0, x, ref, 0)
        dx, dy = E.DeviceArray(x), E.DeviceArray(y0)
        capi.check(capi.lib().gt_tile_spmv(G.handle, 0, capi.GT_PLUS_TIMES_F64, capi.GT_ROW, dx.ptr, dy.ptr))
        np.testing.assert_allclose(dy.download(np.float64, nr), ref, rtol=1e-12)
        # pull (_COL_): y over columns, x over rows
        xr = rng.random(nr); yc0 = rng.random(nc)
        ref = yc0.copy(); og.spmv_f64(0, 0, xr, ref, 1)
        dx2, dy2 = E.DeviceArray(xr), E.DeviceArray(yc0)
        capi.check(capi.lib().gt_tile_spmv(G.handle, 0, capi.GT_PLUS_TIMES_F64, capi.GT_COL, dx2.ptr, dy2.ptr))
        np.testing.assert_allclose(dy2.download(np.float64, nc), ref, rtol=1e-12)
        for d in (dx, dy, dx2, dy2):
            d.free()
    else:
        sr = capi.GT_MIN_PLUS_U32 if w else capi.GT_MIN_SELECT_U32
        x = rng.integers(0, 1 << 20, nc, dtype=np.uint32)
        x[rng.random(nc) < 0.5] = O.INF
        y0 = rng.integers(0, 1 << 21, nr, dtype=np.uint32)
        ref = y0.copy(); og.spmv_u32(0, 0, x, ref)
        dx, dy = E.DeviceArray(x), E.DeviceArray(y0)
        capi.check(capi.lib().gt_tile_spmv(G.handle, 0, sr, capi.GT_ROW, dx.ptr, dy.ptr))
        np.testing.assert_array_equal(dy.download(np.uint32, nr), ref)
        # frontier: same result from the (xi, xv) form, plus the touched flags
        xi = np.nonzero(x != O.INF)[0].astype(np.uint32); rng.shuffle(xi)
        xv = x[xi]
        ref2 = y0.copy(); tref = np.zeros(nr, dtype=np.uint8); og.spmspv_u32(0, 0, xi, xv, ref2, tref)
        assert (ref2 == ref).all()
        dy.upload(y0)
        dxi, dxv, dt = E.DeviceArray(xi), E.DeviceArray(xv), E.DeviceArray(np.zeros(nr, dtype=np.uint8))
        capi.check(capi.lib().gt_tile_spmspv(G.handle, 0, sr, dxi.ptr, dxv.ptr, len(xi), dy.ptr, dt.ptr))
        np.testing.assert_array_equal(dy.download(np.uint32, nr), ref2)
        np.testing.assert_array_equal(dt.download(np.uint8, nr), tref)
        for d in (dx, dy, dxi, dxv, dt):
            d.free()
    G.free(); og.close()


def test_edge_cases():
    """Empty and degenerate inputs the reference's code paths allow."""
    O = _O()
    cases = {
        "empty": (np.zeros((0, 2), dtype="<u4"), 8),
        "self_loop_only": (np.array([[3, 3]], dtype="<u4"), 8),
        "duplicates": (np.array([[1, 2], [2, 1], [1, 2], [1, 2], [5, 6], [0, 7]], dtype="<u4"), 8),
        "star_hub": (np.stack([np.zeros(5000, dtype="<u4"), np.arange(1, 5001, dtype="<u4")], axis=1), 5001),
        "chain": (np.stack([np.arange(0, 300, dtype="<u4"), np.arange(1, 301, dtype="<u4")], axis=1), 301),
        # a frontier column longer than kHeavyColumn (16384): the CTA-per-column heavy path of the SpMSpV
        "big_hub": (np.concatenate([np.stack([np.zeros(40000, dtype="<u4"), np.arange(1, 40001, dtype="<u4")], axis=1),
                                    np.stack([np.arange(1, 40000, dtype="<u4"), np.arange(2, 40001, dtype="<u4")], axis=1)[::7],
                                    np.array([[40001, 0]], dtype="<u4")]), 40002),
        "max_vertex_id": (np.array([[0, 1023], [1023, 0], [1023, 1023]], dtype="<u4"), 1023),
    }
    for name, (tri, n) in cases.items():
        for app in ("pr", "bfs", "cc"):
            ref, rit = O.run_app(app, tri, n, 1, 5 if app == "pr" else (None if app == "cc" else 0))
            mine, it, _, _ = run_gpu(app, tri, n, 5 if app == "pr" else None)
            assert it == rit, (name, app)
            assert_states(app, mine, ref, n + 1)
        w = np.concatenate([tri, ((np.arange(len(tri), dtype="<u4") * 37) % 128 + 1)[:, None]], axis=1).astype("<u4")
        ref, rit = O.run_app("sssp", w, n, 1, 0)
        mine, it, _, _ = run_gpu("sssp", w, n)
        assert it == rit, name
        assert_states("sssp", mine, ref, n + 1)


def test_state_round_trip_and_checksum(fixture_unweighted):
    E = _E()
    G, VR = E.run_pr(loader(fixture_unweighted, 1024), 3)
    V = VR.V
    V2 = V.copy(); V2["rank"] *= 2.0
    VR.set_V(V2)
    back = VR.V
    assert (back["rank"] == V2["rank"]).all() and (back["degree"] == V["degree"]).all()
    VR.free(); G.free()


def test_out_of_range_vertex_is_an_error():
    E = _E()
    from graphtap_b200 import capi
    G = E.Graph()
    with pytest.raises(capi.GraphTapError):
        G.load_triples(np.array([[0, 5000]], dtype="<u4"), 8)


def test_properties_at_scale_20():
    """Size-independent properties at a size the CPU oracle would take too long for in a unit test:
    BFS tree consistency, SSSP fixed point over every edge, CC = minimum id of the component's BFS
    closure, PageRank is a fixed-point iterate (one more iteration from the 19-iteration state)."""
    from graphtap_b200.rmat import rmat_edges
    E = _E()
    scale = 20
    n = 1 << scale
    tri = rmat_edges(scale, nedges=4 << scale, seed=11, weighted=True)
    u, v, w = tri[:, 0].astype(np.int64), tri[:, 1].astype(np.int64), tri[:, 2].astype(np.int64)
    # BFS (undirected, no self loops)
    bfs, it, _, _ = run_gpu("bfs", tri[:, :2].copy(), n)
    hops = bfs["hops"][: n + 1].astype(np.int64); par = bfs["parent"][: n + 1].astype(np.int64)
    reach = hops != 2147483647
    assert hops[0] == 0 and par[0] == 0
    nz = np.nonzero(reach)[0]; nz = nz[nz != 0]
    assert (hops[par[nz]] + 1 == hops[nz]).all()
    m = u != v
    both = reach[u[m]] & reach[v[m]]
    assert (reach[u[m]] == reach[v[m]]).all()                     # an edge never leaves the reached set
    assert (np.abs(hops[u[m]][both] - hops[v[m]][both]) <= 1).all()
    # parent is the minimum-id neighbour one level up (bfs.h:61-63 min-combiner at discovery)
    best = np.full(n + 1, np.iinfo(np.int64).max)
    for a, b in ((u[m], v[m]), (v[m], u[m])):
        ok = reach[a] & reach[b] & (hops[a] + 1 == hops[b])
        np.minimum.at(best, b[ok], a[ok])
    assert (best[nz] == par[nz]).all()
    # CC: labels constant on edges, equal to the smallest id carrying that label
    cc, _, _, _ = run_gpu("cc", tri[:, :2].copy(), n)
    lab = cc["label"][: n + 1].astype(np.int64)
    assert (lab[u] == lab[v]).all() and (lab[lab] == lab).all() and (lab <= np.arange(n + 1)).all()
    # SSSP (directed src -> dst, weights in [1,128]): no edge can relax further, every finite distance is tight
    ss, _, _, _ = run_gpu("sssp", tri, n)
    d = ss["distance"][: n + 1].astype(np.int64)
    fin = d != 2147483647
    mm = (u != v) & fin[u]
    assert fin[v[mm]].all() and (d[v[mm]] <= d[u[mm]] + w[mm]).all()
    tight = np.full(n + 1, np.iinfo(np.int64).max); tight[0] = 0
    np.minimum.at(tight, v[mm], d[u[mm]] + w[mm])
    assert (tight[fin] == d[fin]).all()
    # PageRank: 20 iterations == 19 iterations + 1 through the state hand-over
    G, P = E.run_pr(loader(tri[:, :2].copy(), n), 20)
    r20 = P.V
    P.free(); G.free()
    G, P = E.run_pr(loader(tri[:, :2].copy(), n), 19)
    P.execute(20)                                                  # continues from iteration 19 to 20
    r19p1 = P.V
    P.free(); G.free()
    np.testing.assert_allclose(r19p1["rank"], r20["rank"], rtol=1e-9)


def test_full_size_pagerank_push_vs_pull_scale24():
    """At a size the CPU oracle cannot check in a unit test, the two independent SpMV forms (push over the
    TCSC arrays with RED.ADD, pull over the derived SELL layout) must agree per vertex, degrees must match the
    stored column counts, and the timing knob must report the three phases."""
    E = _E()
    scale = 24
    G = E.Graph(weighted=False)
    G.load_rmat(scale, directed=True, transpose=True, self_loops=True, parallel_edges=True, compression_type=E._TCSC_CF_)
    assert G.info().nnz_global == 16 << scale                     # PageRank keeps duplicates and self loops
    D = E.Deg_Program(G, True, False, False, E._COL_)
    D.execute(1)
    assert D.checksum(quiet=True)[0] == 16 << scale               # sum of out-degrees = stored entries
    out = {}
    for layout in (0, 1):
        P = E.PR_Program(G, True, False, False, E._ROW_)
        P.set("pr_layout", layout)
        P.set("timing", 1)
        P.initialize(D)
        P.execute(20)
        tm = P.timing()
        assert tm.iterations == 20 and tm.combine_ms > 0 and tm.apply_ms > 0 and tm.scatter_gather_ms > 0
        assert tm.combine_ms + tm.apply_ms + tm.scatter_gather_ms <= tm.execute_ms * 1.5 + 5
        out[layout] = P.V
        P.free()
    D.free(); G.free()
    assert (out[0]["degree"] == out[1]["degree"]).all()
    rel = np.abs(out[0]["rank"] - out[1]["rank"]) / out[1]["rank"]
    assert rel.max() < 1e-9, rel.max()
    assert out[1]["rank"].min() >= 0.15 - 1e-12                   # rank = alpha + (1-alpha) * y, y >= 0


def test_sparse_and_dense_paths_agree():
    """activity_filtering_ratio = 0 forces the dense SpMV every iteration, 1.0 forces the frontier SpMSpV
    whenever possible; results and iteration counts must not depend on it (src/vp/vertex_program.hpp:194,768-772)."""
    from graphtap_b200.rmat import rmat_edges
    E = _E()
    tri = rmat_edges(16, seed=9, weighted=True)
    n = 1 << 16
    for app, mk, w, fl in (("bfs", E.BFS_Program, False, dict(directed=False, transpose=False, self_loops=False, parallel_edges=False)),
                           ("sssp", E.SSSP_Program, True, dict(directed=True, transpose=True, self_loops=False, parallel_edges=False))):
        G = E.Graph(weighted=w)
        G.load_triples(tri if w else tri[:, :2].copy(), n, **fl)
        res = {}
        for ratio in (0.0, 0.6, 1.0):
            V = mk(G, False, app == "sssp", app == "bfs", E._ROW_)
            V.set("activity_filtering_ratio", ratio)
            it = V.execute()
            res[ratio] = (it, V.V.copy(), V.timing().sparse_iterations)
            V.free()
        G.free()
        assert res[0.0][2] <= 1 and res[1.0][2] >= res[0.6][2] >= 1
        for ratio in (0.6, 1.0):
            assert res[ratio][0] == res[0.0][0]
            for f in res[ratio][1].dtype.names:
                assert (res[ratio][1][f] == res[0.0][1][f]).all()


def test_degree_program_standalone(fixture_unweighted, golden_fixture):
    """src/apps/deg.cpp: Deg with _ROW_ ordering on the untransposed matrix = out-degree including duplicates and
    self loops; checksum 16384 / reachable 571, maximum 1983 at vertex 613."""
    E = _E()
    G = E.Graph(weighted=False)
    G.load_triples(fixture_unweighted, 1024, directed=True, transpose=False, self_loops=True, parallel_edges=True, compression_type=E._TCSC_)
    V = E.Deg_Program(G, True, False, False, E._ROW_)
    assert V.execute(1) == 1
    d = V.V["degree"]
    assert (d[:1025] == golden_fixture["deg_np1_V"]).all()
    assert V.checksum(quiet=True) == (16384, 571)
    assert int(d.max()) == 1983 and int(d.argmax()) == 613
    V.free(); G.free()


def test_vertex_classification_matches_reference(fixture_unweighted):
    """classify_vertices (src/mat/matrix.hpp:1124-1144) on the fixture with PageRank's flags: the reference's own
    filter statistics are 866 non-empty rows = 550 regular + 316 source, 571 non-empty columns = 550 regular + 21 sink
    (singlenode TCSC stats)."""
    from graphtap_b200 import capi
    E = _E()
    G = E.Graph(weighted=False)
    G.load_triples(fixture_unweighted, 1024, directed=True, transpose=True, self_loops=True, parallel_edges=True, compression_type=E._TCSC_CF_)
    reg, src, snk = C.c_uint32(), C.c_uint32(), C.c_uint32()
    capi.check(capi.lib().gt_graph_classify(G.handle, C.byref(reg), C.byref(src), C.byref(snk)))
    assert (reg.value, src.value, snk.value) == (550, 316, 21)
    assert G.rowgrp_maps(0)[2] == 866 and G.colgrp_maps(0)[2] == 571
    G.free()


def test_sssp_on_the_unweighted_build(rmat12):
    """Without -DHAS_WEIGHT the reference's SSSP combiner takes min(y, x) and the applicator adds 1
    (src/apps/sssp.h:53-56,60-64): hop counts over the directed graph."""
    E, O = _E(), _O()
    tri = rmat12[:, :2].copy()
    fl = dict(O.APP_FLAGS["sssp"]); fl.pop("weighted")
    og = O.OracleGraph(tri, 4096, 1, weighted=0, **fl)
    ref, rit, _ = og.nonstationary(O.SSSP, 0)
    og.close()
    G = E.Graph(weighted=False)
    G.load_triples(tri, 4096, compression_type=E._TCSC_, **{k: bool(v) for k, v in fl.items()})
    V = E.SSSP_Program(G, False, True, False, E._ROW_)
    V.root = 0
    assert V.execute() == rit
    assert (V.V["distance"][:4097] == ref[:4097]).all()
    V.free(); G.free()


def test_execute_in_pieces_and_phase_by_phase(rmat12):
    """execute(k) continues where the last call stopped (iteration counts are absolute, vertex_program.hpp:407-433), and
    the three phases driven one by one give the same states as execute()."""
    E, O = _E(), _O()
    ref, rit = O.run_app("sssp", rmat12, 4096, 1, 0)
    G = E.Graph(weighted=True)
    G.load_triples(rmat12, 4096, directed=True, transpose=True, self_loops=False, parallel_edges=False, compression_type=E._TCSC_)
    V = E.SSSP_Program(G, False, True, False, E._ROW_)
    V.execute(2); V.execute(4)
    assert V.execute() == rit
    assert (V.V["distance"][:4097] == ref[:4097]).all()
    V.free(); G.free()
    tu = rmat12[:, :2].copy()
    G = E.Graph(weighted=False)
    G.load_triples(tu, 4096, directed=False, transpose=False, self_loops=True, parallel_edges=False, compression_type=E._TCSC_)
    A = E.CC_Program(G, False, True, False, E._ROW_); A.execute(3)
    B = E.CC_Program(G, False, True, False, E._ROW_)
    for _ in range(3):
        for ph in (0, 1, 2):
            B.run_phase(ph)
    assert (A.V["label"] == B.V["label"]).all()
    refc, _ = O.run_app("cc", tu, 4096, 1, None)
    B.execute()
    assert (B.V["label"][:4097] == refc[:4097]).all()
    A.free(); B.free(); G.free()


@pytest.mark.parametrize("scale,seed", [(12, 12), (17, 4)])
def test_bfs_bottom_up_equals_top_down(scale, seed):
    """The bottom-up pass (one GPU, undirected graph) must leave exactly the parents and hops of the reference's push:
    the first active neighbour in ascending order is the minimum the min-combiner would keep (src/apps/bfs.h:61-63)."""
    from graphtap_b200.rmat import rmat_edges
    E, O = _E(), _O()
    tri = rmat_edges(scale, seed=seed)
    n = 1 << scale
    ref, rit = O.run_app("bfs", tri, n, 1, 0)
    G = E.Graph(weighted=False)
    G.load_triples(tri, n, directed=False, transpose=False, self_loops=False, parallel_edges=False, compression_type=E._TCSC_)
    for ratio in (0.0, 0.05, 1e-9):                 # never / default / from the second iteration on
        V = E.BFS_Program(G, False, False, True, E._ROW_)
        V.set("bfs_bottom_up_ratio", ratio)
        assert V.execute() == rit
        mine = V.V
        for f in ("parent", "hops"):
            assert (mine[f][: n + 1] == ref[f][: n + 1]).all(), (ratio, f)
        V.free()
    G.free()


# ---- _TCSC_CF_: computation filtering ------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["fixture", "rmat12"])
def test_cf_lists_and_classes_match_reference(which, fixture_unweighted, rmat12, golden_fixture, golden_rmat12):
    """TCSC_CF_BASE::populate on the device: IA in the reference's order after its source-row swap, the four
    (start,end)-pair lists with their column lists — quirks included — and classify_vertices' id lists, against the
    dumps of the unmodified reference and against the CPU oracle."""
    E, O = _E(), _O()
    tri, n, gold = (fixture_unweighted, 1024, golden_fixture) if which == "fixture" else (rmat12[:, :2].copy(), 4096, golden_rmat12)
    G = E.Graph(weighted=False)
    G.load_triples(tri, n, directed=True, transpose=True, self_loops=True, parallel_edges=True, compression_type=E._TCSC_CF_)
    t, cf = G.tile(0), G.tile_cf(0)
    base = "pr_np1_tile_r0.t0"
    np.testing.assert_array_equal(t["JA"], gold[base + "_JA"])
    np.testing.assert_array_equal(t["IA"], gold[base + "_IA"])
    assert [cf[f"NC{k}"] for k in range(4)] == list(gold[base + "_cfnc"])
    for k in range(4):
        np.testing.assert_array_equal(cf[f"JA{k}"], gold[base + f"_cf{k}.JA"])
        np.testing.assert_array_equal(cf[f"JC{k}"], gold[base + f"_cf{k}.JC"])
    reg, src, snk = G.classify_lists()
    np.testing.assert_array_equal(reg, gold["pr_np1_regrows_r0"])
    np.testing.assert_array_equal(src, gold["pr_np1_srcrows_r0"])
    np.testing.assert_array_equal(snk, gold["pr_np1_snkcols_r0"])
    fl = dict(O.APP_FLAGS["pr"]); fl.pop("weighted")
    og = O.OracleGraph(tri, n, 1, weighted=0, **fl)
    ot = og.cf_tile(0, 0)
    assert [cf[f"filled{k}"] for k in range(4)] == [ot[f"filled{k}"] for k in range(4)]
    og.close(); G.free()


@pytest.mark.parametrize("layout", [1, 0])
@pytest.mark.parametrize("which", ["fixture", "rmat12"])
def test_pagerank_convergence_mode(which, layout, fixture_unweighted, rmat12, golden_fixture, golden_rmat12):
    """`pr <file> <n>` without an iteration count.  _TCSC_CF_ (pr.cpp): has_converged() looks at the regular rows only and
    the source rows end at alpha (fixture: 12 iterations, checksum 51); _TCSC_ (pr1.cpp): the plain loop (13 / 70)."""
    E = _E()
    tri, n, gold = (fixture_unweighted, 1024, golden_fixture) if which == "fixture" else (rmat12[:, :2].copy(), 4096, golden_rmat12)
    for comp, key in ((E._TCSC_CF_, "prconv"), (E._TCSC_, "pr1conv")):
        G, P = E.run_pr(loader(tri, n), 0, compression=comp, pr_layout=layout)
        mine, it, cs = P.V, P.iteration, P.checksum(quiet=True)
        P.free(); G.free()
        meta = gold[f"{key}_np1_meta"]
        assert it == meta[0], (key, it)
        assert_states("pr", mine, gold[f"{key}_np1_V"], n + 1)
        if layout == 0:                                 # the push path keeps the reference's vertex order in the checksum loop
            assert cs[1] == meta[2]
    if which == "fixture":
        G, P = E.run_pr(loader(tri, n), 0, compression=E._TCSC_CF_, pr_layout=layout)
        assert (P.iteration, P.checksum(quiet=True)) == (12, (51, 1025))
        P.free(); G.free()


def test_pagerank_cf_schedule_at_scale_20():
    """Fixed-iteration PageRank on a _TCSC_CF_ graph (REG x REG every iteration, REG x SNK first, source rows last) against the
    same run on a _TCSC_ graph, per vertex, plus the 1-iteration and 2-iteration corner cases of the schedule."""
    from graphtap_b200.rmat import rmat_edges
    E = _E()
    tri = rmat_edges(20, nedges=8 << 20, seed=5)
    n = 1 << 20
    for iters in (1, 2, 20):
        out = {}
        for comp in (E._TCSC_, E._TCSC_CF_):
            G, P = E.run_pr(loader(tri, n), iters, compression=comp)
            out[comp] = P.V
            P.free(); G.free()
        assert (out[E._TCSC_]["degree"] == out[E._TCSC_CF_]["degree"]).all()
        rel = np.abs(out[E._TCSC_]["rank"] - out[E._TCSC_CF_]["rank"]) / out[E._TCSC_]["rank"]
        assert rel.max() <= 1e-12, (iters, rel.max())


def test_partitioned_ingest_on_one_rank_is_the_global_build():
    """gt_graph_build_partitioned with a single rank has nothing to route: same arrays as gt_graph_build (the multi-rank
    exchange is checked by tools/multi_gpu_check.py on 2/4/8 GPUs), and the timing samples come back per iteration."""
    E = _E()
    from graphtap_b200.rmat import rmat_edges
    tri = rmat_edges(12, seed=5, weighted=True)
    fl = dict(directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=E._TCSC_)
    A = E.Graph(weighted=True).load_triples(tri, 1 << 12, **fl)
    B = E.Graph(weighted=True).load_triples(tri, 1 << 12, partitioned=True, **fl)
    Cg = E.Graph(weighted=True).load_rmat(12, seed=5, partitioned=True, **fl)
    ta, tb, tc = A.tile(0), B.tile(0), Cg.tile(0)
    for f in ("JA", "IA", "A", "JC", "IR"):
        assert (ta[f] == tb[f]).all() and (ta[f] == tc[f]).all(), f
    assert A.info().nedges_input == B.info().nedges_input == Cg.info().nedges_input == tri.shape[0]
    V = E.SSSP_Program(A, False, True, False, E._ROW_)
    V.set("timing", 1)
    it = V.execute()
    n = C.c_uint32()
    from graphtap_b200.capi import lib, check
    for phase in (0, 1, 2):
        buf = (C.c_double * 64)()
        check(lib().gt_program_timing_samples(V.handle, phase, buf, 64, C.byref(n)))
        assert n.value == it and all(buf[i] >= 0 for i in range(it))
        tm = V.timing()
        total = (tm.scatter_gather_ms, tm.combine_ms, tm.apply_ms)[phase]
        assert abs(sum(buf[i] for i in range(it)) - total) <= 1e-9 + 1e-9 * total
    check(lib().gt_program_timing_samples(V.handle, 3, buf, 64, C.byref(n)))
    assert n.value == 1 and buf[0] > 0
    V.free(); A.free(); B.free(); Cg.free()
