"""Parity at BASELINE-config scale against the UNMODIFIED reference, per vertex.

`oracle/_ref/ref_{bfs,pr,sssp}` (the reference's own code, compiled from /root/reference by
oracle/ref_build/Makefile; the binaries travel to the GPU box) run on the host cores of the GPU box on a
seeded RMAT scale-22 edge file; the CUDA path runs the same graph through the C ABI (the device generator
emits the same stream, tests/test_rmat.py + test_gpu_parity.py::test_device_rmat_matches_host).

  * BFS root 0, RMAT-22            = BASELINE.json configs[1] exactly: parents and hops bit-exact
  * PageRank 20 iterations, RMAT-22: every rank within 1e-6 relative (north_star), degrees exact
  * SSSP root 0, weighted RMAT-22  : distances bit-exact
  * Connected components, RMAT-22  : labels bit-exact (BASELINE.json configs[4] is the same program at RMAT-27 on 8 GPUs)

The reference runs at np = 16 (or the largest power of two the host offers), the GPU at p = 1: integer
results are p-independent, PageRank differs by f64 summation order only."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALE = 22
PR_RTOL = 1e-6


def _np_ranks():
    n, p = os.cpu_count() or 1, 1
    while p * 2 <= min(n, 16):
        p *= 2
    return p


@pytest.fixture(scope="module")
def edge_files():
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref is not built (needs the container that has /root/reference)")
    d = tempfile.mkdtemp(prefix="gtcfg_")
    pu, pw = os.path.join(d, "u.bin"), os.path.join(d, "w.bin")
    O.write_rmat(pu, SCALE, seed=SCALE, weighted=False)
    O.write_rmat(pw, SCALE, seed=SCALE, weighted=True)
    yield pu, pw
    for p in (pu, pw):
        os.remove(p)
    os.rmdir(d)


def _gpu(app, weighted, arg):
    from graphtap_b200 import engine as E

    def load(G, **fl):
        ct = fl.pop("compression_type")
        G.load_rmat(SCALE, seed=SCALE, compression_type=ct, **fl)

    if app == "pr":
        G, V = E.run_pr(load, arg)
    elif app == "bfs":
        G, V = E.run_bfs(load, arg)
    elif app == "cc":
        G, V = E.run_cc(load)
    else:
        G, V = E.run_sssp(load, arg)
    out = V.V, V.iteration, V.checksum(quiet=True), V.timing()
    V.free(); G.free()
    return out


def _ref(app, path, arg):
    from oracle import oracle as O
    V, it, stdout, _ = O.ref_run(app, path, 1 << SCALE, arg, np_ranks=_np_ranks())
    g = lambda k: int([l for l in stdout.splitlines() if l.startswith(k)][-1].split()[-1])
    return V, it, (g("Value checksum:"), g("Reachable vertices:"))


def test_bfs_rmat22_root0_vs_reference(edge_files):
    ref, rit, rcs = _ref("bfs", edge_files[0], 0)
    mine, it, cs, tm = _gpu("bfs", False, 0)
    n = (1 << SCALE) + 1
    assert it == rit
    for f in ("parent", "hops", "vid"):
        assert (mine[f][:n] == ref[f][:n]).all(), f
    assert cs == rcs
    assert (mine["hops"][:n] != 2147483647).sum() > n // 4          # root 0 reaches the giant component


def test_pagerank_rmat22_vs_reference(edge_files):
    ref, rit, rcs = _ref("pr", edge_files[0], 20)
    mine, it, cs, _ = _gpu("pr", False, 20)
    n = (1 << SCALE) + 1
    assert it == rit == 20
    assert (mine["degree"][:n] == ref["degree"][:n]).all()
    rel = np.abs(mine["rank"][:n] - ref["rank"][:n]) / np.abs(ref["rank"][:n])
    assert rel.max() <= PR_RTOL, rel.max()
    assert cs[1] == rcs[1]                       # the truncating value checksum depends on the summation order over ranks (K11)
    assert abs(mine["rank"][:n].sum() - ref["rank"][:n].sum()) <= 1e-9 * ref["rank"][:n].sum()


def test_sssp_rmat22_root0_vs_reference(edge_files):
    ref, rit, rcs = _ref("sssp", edge_files[1], 0)
    mine, it, cs, _ = _gpu("sssp", True, 0)
    n = (1 << SCALE) + 1
    assert it == rit
    m = mine["distance"]
    assert (m[:n] == ref[:n]).all()
    assert cs == rcs


def test_cc_rmat22_vs_reference(edge_files):
    ref, rit, rcs = _ref("cc", edge_files[0], None)
    mine, it, cs, _ = _gpu("cc", False, None)
    n = (1 << SCALE) + 1
    assert it == rit
    m = mine["label"]
    assert (m[:n] == ref[:n]).all()
    assert cs == rcs
