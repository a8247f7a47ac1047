"""ctypes/numpy face of the CPU oracle (``oracle/gt_oracle.c``) and of the reference-built binaries
(``oracle/_ref``).  TEST INFRASTRUCTURE: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs only — never by ``graphtap_b200``."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
INF = 2147483647
DEG, PR, BFS, CC, SSSP = range(5)

# reference Vertex_State layouts (src/apps/*.h)
PR_STATE = np.dtype([("degree", "<u4"), ("_pad", "<u4"), ("rank", "<f8")])
BFS_STATE = np.dtype([("parent", "<u4"), ("hops", "<u4"), ("vid", "<u4")])
U32_STATE = np.dtype("<u4")

# per-app graph flags of the reference drivers (src/apps/{pr,bfs,cc,sssp}.cpp)
APP_FLAGS = {
    "pr": dict(directed=1, transpose=1, self_loops=1, acyclic=0, parallel_edges=1, weighted=0),
    "deg": dict(directed=1, transpose=0, self_loops=1, acyclic=0, parallel_edges=1, weighted=0),
    "bfs": dict(directed=0, transpose=0, self_loops=0, acyclic=0, parallel_edges=0, weighted=0),
    "cc": dict(directed=0, transpose=0, self_loops=1, acyclic=0, parallel_edges=0, weighted=0),
    "sssp": dict(directed=1, transpose=1, self_loops=0, acyclic=0, parallel_edges=0, weighted=1),
}

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        l.gto_layout.argtypes = [u32, i32, vp, vp, vp, vp, vp]
        l.gto_build.restype = vp
        l.gto_build.argtypes = [vp, u64, i32, u32, i32, i32, i32, i32, i32, i32]
        l.gto_free.argtypes = [vp]
        l.gto_tile_height.restype = u32; l.gto_tile_height.argtypes = [vp]
        l.gto_nnz.restype = u64; l.gto_nnz.argtypes = [vp]
        l.gto_tile_nnz.restype = u64; l.gto_tile_nnz.argtypes = [vp, u32, u32]
        for f in ("gto_tile_JA", "gto_tile_IA", "gto_tile_A"):
            getattr(l, f).restype = vp; getattr(l, f).argtypes = [vp, u32, u32]
        l.gto_seg_nnz.restype = u32; l.gto_seg_nnz.argtypes = [vp, i32, u32]
        for f in ("gto_seg_bits", "gto_seg_prefix", "gto_seg_ids"):
            getattr(l, f).restype = vp; getattr(l, f).argtypes = [vp, i32, u32]
        l.gto_tile_spmv_f64.argtypes = [vp, u32, u32, i32, vp, vp]
        l.gto_tile_spmv_u32.argtypes = [vp, u32, u32, i32, vp, vp, vp]
        l.gto_tile_spmspv_u32.argtypes = [vp, u32, u32, i32, vp, vp, u32, vp, vp]
        l.gto_degree.argtypes = [vp, i32, vp]
        l.gto_pagerank.restype = u32; l.gto_pagerank.argtypes = [vp, u32, C.c_double, C.c_double, vp, vp]
        l.gto_pagerank_cf.restype = u32; l.gto_pagerank_cf.argtypes = [vp, u32, C.c_double, C.c_double, vp, vp]
        l.gto_classify.argtypes = [vp, u32, vp]
        l.gto_cf_build.restype = vp; l.gto_cf_build.argtypes = [vp, u32, u32]
        l.gto_cf_free.argtypes = [vp]
        for f in ("gto_cf_nc", "gto_cf_filled"):
            getattr(l, f).restype = u32; getattr(l, f).argtypes = [vp, i32]
        for f in ("gto_cf_ja", "gto_cf_jc"):
            getattr(l, f).restype = vp; getattr(l, f).argtypes = [vp, i32]
        for f in ("gto_cf_ia", "gto_cf_a"):
            getattr(l, f).restype = vp; getattr(l, f).argtypes = [vp]
        l.gto_nonstationary.restype = u32; l.gto_nonstationary.argtypes = [vp, i32, u32, C.c_double, vp, vp, vp]
        l.gto_checksum_f64.argtypes = [vp, u64, vp, vp]
        l.gto_checksum_u32.argtypes = [vp, u64, u32, vp, vp]
        l.gto_rmat.argtypes = [u32, u64, u64, u64, i32, vp]
        _lib = l
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _view(ptr, n, dtype):
    if not n:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def layout(nvertices: int, p: int):
    tr = np.zeros(p * p, dtype=np.int32); lead = np.zeros(p, dtype=np.int32)
    th, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().gto_layout(nvertices, p, _ptr(tr), _ptr(lead), C.byref(th), C.byref(a), C.byref(b))
    return dict(tile_rank=tr.reshape(p, p), leader_ranks=lead, tile_height=th.value, rowgrp_nranks=a.value, colgrp_nranks=b.value)


class OracleGraph:
    """p x p TCSC tiles built on the CPU from the global edge list."""

    def __init__(self, triples: np.ndarray, nvertices: int, p: int = 1, *, directed=1, transpose=0, self_loops=1,
                 acyclic=0, parallel_edges=1, weighted=0):
        triples = np.ascontiguousarray(triples, dtype="<u4")
        assert triples.ndim == 2 and triples.shape[1] == (3 if weighted else 2)
        self.p, self.weighted, self.nvertices = p, weighted, nvertices
        self.h = lib().gto_build(_ptr(triples), triples.shape[0], int(weighted), nvertices, p, int(directed), int(transpose),
                                 int(self_loops), int(acyclic), int(parallel_edges))
        self.th = lib().gto_tile_height(self.h)
        self.nnz = lib().gto_nnz(self.h)

    def close(self):
        if self.h:
            lib().gto_free(self.h); self.h = None

    def __del__(self):
        self.close()

    def seg(self, is_col: bool, s: int):
        n = lib().gto_seg_nnz(self.h, int(is_col), s)
        return dict(nnz=n, bits=_view(lib().gto_seg_bits(self.h, int(is_col), s), self.th, "u1"),
                    prefix=_view(lib().gto_seg_prefix(self.h, int(is_col), s), self.th, "<u4"),
                    ids=_view(lib().gto_seg_ids(self.h, int(is_col), s), n, "<u4"))

    def tile(self, rg: int, cg: int):
        nnz = lib().gto_tile_nnz(self.h, rg, cg)
        nc = lib().gto_seg_nnz(self.h, 1, cg)
        return dict(nnz=nnz, JA=_view(lib().gto_tile_JA(self.h, rg, cg), nc + 1, "<u4"),
                    IA=_view(lib().gto_tile_IA(self.h, rg, cg), nnz, "<u4"),
                    A=_view(lib().gto_tile_A(self.h, rg, cg), nnz, "<u4") if self.weighted else None)

    def spmv_f64(self, rg, cg, x, y, ordering=0):
        x = np.ascontiguousarray(x, dtype=np.float64); assert y.dtype == np.float64 and y.flags.c_contiguous
        lib().gto_tile_spmv_f64(self.h, rg, cg, ordering, _ptr(x), _ptr(y))

    def spmv_u32(self, rg, cg, x, y, t=None):
        x = np.ascontiguousarray(x, dtype=np.uint32); assert y.dtype == np.uint32 and y.flags.c_contiguous
        lib().gto_tile_spmv_u32(self.h, rg, cg, int(self.weighted), _ptr(x), _ptr(y), _ptr(t) if t is not None else None)

    def spmspv_u32(self, rg, cg, xi, xv, y, t=None):
        xi = np.ascontiguousarray(xi, dtype=np.uint32); xv = np.ascontiguousarray(xv, dtype=np.uint32)
        lib().gto_tile_spmspv_u32(self.h, rg, cg, int(self.weighted), _ptr(xi), _ptr(xv), len(xi), _ptr(y), _ptr(t) if t is not None else None)

    def degree(self, ordering=1):
        d = np.zeros(self.p * self.th, dtype=np.uint32)
        lib().gto_degree(self.h, ordering, _ptr(d))
        return d

    def classify(self, s: int) -> np.ndarray:
        """classify_vertices of vertex segment s: 1 regular, 2 source row, 3 sink column, 0 neither (u8[tile_height])."""
        out = np.zeros(self.th, dtype=np.uint8)
        lib().gto_classify(self.h, s, _ptr(out))
        return out

    def cf_tile(self, rg: int, cg: int) -> dict:
        """TCSC_CF_BASE::populate of tile (rg, cg): CF-ordered IA / A and the four pair lists
        (kind 0 REG_R_REG_C, 1 REG_R_SNK_C, 2 SRC_R_REG_C, 3 SRC_R_SNK_C)."""
        nnz = lib().gto_tile_nnz(self.h, rg, cg)
        if not nnz:
            return dict(nnz=0)
        t = lib().gto_cf_build(self.h, rg, cg)
        d = dict(nnz=nnz, IA=_view(lib().gto_cf_ia(t), nnz, "<u4"), A=_view(lib().gto_cf_a(t), nnz, "<u4"))
        for k in range(4):
            nc = lib().gto_cf_nc(t, k)
            d[f"NC{k}"], d[f"filled{k}"] = nc, lib().gto_cf_filled(t, k)
            d[f"JA{k}"] = _view(lib().gto_cf_ja(t, k), 2 * nc, "<u4")
            d[f"JC{k}"] = _view(lib().gto_cf_jc(t, k), nc, "<u4")
        lib().gto_cf_free(t)
        return d

    def pagerank(self, iters=20, alpha=0.15, tol=1e-5, cf=False):
        """iters == 0: until convergence.  cf: the _TCSC_CF_ computation-filtering schedule (pr.cpp) instead of _TCSC_ (pr1.cpp)."""
        n = self.p * self.th
        V = np.zeros(n, dtype=PR_STATE); rank = np.zeros(n); deg = np.zeros(n, dtype=np.uint32)
        it = (lib().gto_pagerank_cf if cf else lib().gto_pagerank)(self.h, iters, alpha, tol, _ptr(rank), _ptr(deg))
        V["rank"], V["degree"] = rank, deg
        return V, it

    def nonstationary(self, app: int, root=0, ratio=0.6):
        n = self.p * self.th
        a = np.zeros(n, dtype=np.uint32); b = np.zeros(n, dtype=np.uint32); ns = C.c_uint32()
        it = lib().gto_nonstationary(self.h, app, root, ratio, _ptr(a), _ptr(b), C.byref(ns))
        if app == BFS:
            V = np.zeros(n, dtype=BFS_STATE); V["parent"], V["hops"], V["vid"] = a, b, np.arange(n, dtype=np.uint32)
        else:
            V = a
        return V, it, ns.value


def run_app(app: str, triples: np.ndarray, nvertices: int, p: int = 1, arg: int | None = None):
    """The reference drivers (src/apps/*.cpp) on the CPU oracle: returns (V[p*tile_height], iterations)."""
    fl = dict(APP_FLAGS[app]); w = fl.pop("weighted")
    g = OracleGraph(triples, nvertices, p, weighted=w, **fl)
    try:
        if app == "pr":
            V, it = g.pagerank(20 if arg is None else arg)
        elif app == "deg":
            V, it = g.degree(0), 1
        else:
            V, it, _ = g.nonstationary({"bfs": BFS, "cc": CC, "sssp": SSSP}[app], 0 if arg is None else arg)
    finally:
        g.close()
    return V, it


def checksum(app: str, V: np.ndarray, nrows: int):
    """Vertex_Program::checksum (src/vp/vertex_program.hpp:1926-1960) on the first `nrows` states."""
    s, c = C.c_uint64(), C.c_uint64()
    if app == "pr":
        v = np.ascontiguousarray(V["rank"][:nrows])
        lib().gto_checksum_f64(_ptr(v), nrows, C.byref(s), C.byref(c))
    else:
        f = {"bfs": "hops", "deg": None, "cc": None, "sssp": None}[app]
        v = np.ascontiguousarray((V[f] if f else V)[:nrows], dtype=np.uint32)
        lib().gto_checksum_u32(_ptr(v), nrows, 0 if app == "deg" else INF, C.byref(s), C.byref(c))
    return s.value, c.value


def rmat_edges(scale: int, nedges: int | None = None, seed: int | None = None, weighted: bool = False, first_edge: int = 0):
    """Host RMAT stream in C (same stream as graphtap_b200.rmat.rmat_edges, ~50x faster); used to write the
    CPU-baseline samples."""
    n = (16 << scale) if nedges is None else nedges
    out = np.empty((n, 3 if weighted else 2), dtype="<u4")
    lib().gto_rmat(scale, first_edge, n, scale if seed is None else seed, int(weighted), _ptr(out))
    return out


def write_rmat(path: str, scale: int, seed: int | None = None, weighted: bool = False, chunk: int = 1 << 22, threads: int | None = None) -> int:
    """Writes the reference's headerless binary edge file (src/ds/triple.hpp:9-18).  The generator is a ctypes call
    (the GIL is released), so chunks are produced by a thread pool and written at their offsets."""
    from concurrent.futures import ThreadPoolExecutor
    n = 16 << scale
    rec = 12 if weighted else 8
    lib()
    fd = os.open(path, os.O_CREAT | os.O_WRONLY | os.O_TRUNC, 0o600)
    try:
        os.ftruncate(fd, n * rec)

        def work(first):
            a = rmat_edges(scale, min(chunk, n - first), seed, weighted, first)
            os.pwrite(fd, a.tobytes() if not a.flags.c_contiguous else memoryview(a).cast("B"), first * rec)

        with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
            list(ex.map(work, range(0, n, chunk)))
    finally:
        os.close(fd)
    return n


# ---- the unmodified reference (oracle/_ref) --------------------------------------------------------------
STATE_OF = {"pr": PR_STATE, "pr1": PR_STATE, "bfs": BFS_STATE, "cc": U32_STATE, "sssp": U32_STATE, "deg": U32_STATE}


def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_pr"))


def ref_run(app: str, path: str, nvertices: int, arg: int | None = None, np_ranks: int = 1, tiles: bool = False,
            keep_dir: str | None = None, extra: list[str] | None = None):
    """Run oracle/_ref/ref_<app> (the reference's own code) and gather the per-rank dumps into one
    array of p*tile_height states.  Returns (V, iterations, stdout, dump_prefix)."""
    d = keep_dir or tempfile.mkdtemp(prefix="gtref_")
    prefix = os.path.join(d, "o")
    cmd = [os.path.join(REF_DIR, "ref_" + app), path, str(nvertices)] + ([str(arg)] if arg is not None else []) + ["--dump", prefix]
    if tiles:
        cmd.append("--tiles")
    if extra:
        cmd += extra
    env = dict(os.environ)
    if np_ranks > 1:
        env["GT_MPI_NP"] = str(np_ranks)
    else:
        env.pop("GT_MPI_NP", None)
    out = subprocess.run(cmd, capture_output=True, text=True, check=True, env=env).stdout
    dt = STATE_OF[app]
    segs = {}
    th = None
    for r in range(np_ranks):
        meta = open(f"{prefix}.r{r}.meta").read().split()
        seg, th = int(meta[2]), int(meta[3])
        segs[seg] = np.fromfile(f"{prefix}.r{r}.V.bin", dtype=dt)
        assert len(segs[seg]) == th
    V = np.concatenate([segs[s] for s in sorted(segs)])
    it = int([l for l in out.splitlines() if l.startswith("Iterations:")][-1].split()[1])
    return V, it, out, prefix


def ref_execute_seconds(stdout: str) -> float:
    """The reference's own 'Execute time' line (src/vp/vertex_program.hpp:437); the last one printed."""
    lines = [l for l in stdout.splitlines() if l.startswith("Execute time:")]
    return float(lines[-1].split()[2])
