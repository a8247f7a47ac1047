"""Host-side mirror of the reference's public interface, on top of the C ABI.

Same names, argument meaning and call order as the reference drivers (``src/apps/*.cpp``):

    Env.init(); G = Graph(); G.load(path, n, n, directed, transpose, self_loops, acyclic,
    parallel_edges, _2DT_, _TCSC_); V = BFS_Program(G, stationary, gather_depends_on_apply,
    apply_depends_on_iter, _ROW_); V.root = r; V.execute(); V.checksum(); V.display(); V.free();
    G.free(); Env.finalize()

(``src/mat/graph.hpp:41-43``, ``src/vp/vertex_program.hpp:27-62``).  Everything that computes goes through
``libgraphtap_b200.so``; this file holds no numerical code.  Errors surface as ``GraphTapError``
(the reference prints to stderr and calls ``Env::exit(1)``, ``src/mpi/env.hpp:159-162``).
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import capi
from .capi import (GT_APP_BFS, GT_APP_CC, GT_APP_DEG, GT_APP_PR, GT_APP_SSSP, GT_COL, GT_ROW, GT_TCSC, GT_TCSC_CF,
                   check, lib)

# reference enum spellings (src/mat/tiling.hpp:12-15, src/ds/compressed_column.hpp:17-23, vertex_program.hpp:17-21)
_2D_, _2DT_ = 0, 1
_CSC_, _DCSC_, _TCSC_, _TCSC_CF_ = 0, 1, GT_TCSC, GT_TCSC_CF
_ROW_, _COL_ = GT_ROW, GT_COL

# Vertex_State layouts of the reference (src/apps/*.h), as numpy structured dtypes
PR_STATE = np.dtype([("degree", "<u4"), ("_pad", "<u4"), ("rank", "<f8")])
BFS_STATE = np.dtype([("parent", "<u4"), ("hops", "<u4"), ("vid", "<u4")])
DEG_STATE = np.dtype([("degree", "<u4")])
CC_STATE = np.dtype([("label", "<u4")])
SSSP_STATE = np.dtype([("distance", "<u4")])
INF = capi.GT_INF_U32


class Env:
    """``src/mpi/env.hpp``: process-wide rank / nranks / communicators.  One process per GPU."""
    rank = 0
    nranks = 1
    is_master = True
    ctx = None
    _dist = None
    quiet = False          # True: no "Execute time" / checksum lines on stdout (bench.py prints exactly one JSON line)

    @classmethod
    def init(cls, device: int | None = None, use_torch_distributed: bool | None = None) -> None:
        if cls.ctx is not None:
            return
        rank = int(os.environ.get("RANK", "0"))
        nranks = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = local
        uid = None
        if nranks > 1:
            # torch.distributed is plumbing only: it carries the 128-byte NCCL id from rank 0.
            import torch
            import torch.distributed as dist
            if not dist.is_initialized():
                torch.cuda.set_device(device)
                dist.init_process_group(backend="nccl", device_id=torch.device("cuda", device))
            cls._dist = dist
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                check(lib().gt_nccl_unique_id(buf))
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=f"cuda:{device}")
            dist.broadcast(t, src=0)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        ctx = C.c_void_p()
        check(lib().gt_ctx_create(device, rank, nranks, uid, C.byref(ctx)))
        cls.ctx, cls.rank, cls.nranks, cls.is_master = ctx, rank, nranks, rank == 0

    @classmethod
    def barrier(cls) -> None:
        if cls.ctx is not None:
            check(lib().gt_ctx_barrier(cls.ctx))          # Env::barrier = MPI_Barrier on the world (src/mpi/env.hpp:164-166)

    @classmethod
    def finalize(cls) -> None:
        if cls.ctx is not None:
            check(lib().gt_ctx_destroy(cls.ctx))
            cls.ctx = None
        if cls._dist is not None and cls._dist.is_initialized():
            cls._dist.destroy_process_group()
            cls._dist = None

    @classmethod
    def print_time(cls, preamble: str, seconds: float) -> None:
        if cls.is_master and not cls.quiet:
            print(f"{preamble} time: {seconds:f} seconds")


def _looks_like_text(path: str) -> bool:
    with open(path, "rb") as f:
        head = f.read(4096)
    return len(head) > 0 and all(b in b"0123456789 \t\r\n#%.-+eE" or 32 <= b < 127 for b in head) and b"\x00" not in head


def read_edge_list(path: str, weighted: bool) -> np.ndarray:
    """Host-side reader of the reference's two input formats, returning (n, 2|3) uint32 records.

    binary: headless array of ``{u32 row, u32 col[, u32 weight]}`` (``src/mat/graph.hpp:307-372``);
    text:   leading lines starting with '#' or '%' (or empty) are skipped, then one ``row col[ weight]`` per line,
            single-space separated, up to the first empty line (``src/mat/graph.hpp:194-304``); a line with the wrong
            number of fields is the reference's ``read() failure`` error."""
    rec = 3 if weighted else 2
    if not _looks_like_text(path):
        data = np.fromfile(path, dtype="<u4")
        if data.size % rec:
            raise capi.GraphTapError(capi.GT_ERR_INVALID, f"{path}: size is not a multiple of the {rec * 4}-byte record")
        return data.reshape(-1, rec)
    rows = []
    with open(path, "r") as f:
        started = False
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if not started:
                if line == "" or line[0] in "#%":
                    continue
                started = True
            if line == "":
                break
            fields = line.split(" ")
            if len(fields) != rec:
                raise capi.GraphTapError(capi.GT_ERR_INVALID, f'read() failure "{line}"')
            rows.append([int(x) for x in fields])
    return np.asarray(rows, dtype="<u4").reshape(-1, rec)


def share_of_file(filesize: int, rec_bytes: int, rank: int, nranks: int) -> tuple[int, int]:
    """Byte range [offset, endpos) of a binary edge file that ``rank`` reads, exactly ``Graph::parread_binary``'s split
    (``src/mat/graph.hpp:317-323``): equal whole-record shares, the last rank also takes the remainder."""
    share = (filesize // nranks) // rec_bytes * rec_bytes
    offset = share * rank
    endpos = filesize if rank == nranks - 1 else offset + share
    return offset, endpos


def read_edge_list_share(path: str, weighted: bool, rank: int, nranks: int) -> np.ndarray:
    """This rank's share of the records.  Binary files: only the share's bytes are read from disk.  Text files: the
    reference splits at byte offsets and re-synchronises on line ends (``src/mat/graph.hpp:194-304``); here every rank
    parses the file and keeps the records whose index falls in its equal share (same union, simpler)."""
    rec = 3 if weighted else 2
    if not _looks_like_text(path):
        size = os.path.getsize(path)
        if size % (rec * 4):
            raise capi.GraphTapError(capi.GT_ERR_INVALID, f"{path}: size is not a multiple of the {rec * 4}-byte record")
        offset, endpos = share_of_file(size, rec * 4, rank, nranks)
        return np.fromfile(path, dtype="<u4", count=(endpos - offset) // 4, offset=offset).reshape(-1, rec)
    allrec = read_edge_list(path, weighted)
    share = allrec.shape[0] // nranks
    return allrec[share * rank: allrec.shape[0] if rank == nranks - 1 else share * (rank + 1)]


class Graph:
    """``Graph<Weight, Integer_Type, Fractional_Type>`` (``src/mat/graph.hpp:33-67``)."""

    def __init__(self, weighted: bool = False):
        self.weighted = weighted          # the reference selects this at compile time (-DHAS_WEIGHT)
        self.handle = None

    def _flags(self, directed, transpose, self_loops, acyclic, parallel_edges):
        return capi.GraphFlags(int(directed), int(transpose), int(self_loops), int(acyclic), int(parallel_edges))

    def load(self, filepath, nrows, ncols, directed=True, transpose=False, self_loops=True, acyclic=False,
             parallel_edges=True, tiling_type=_2DT_, compression_type=_TCSC_, partitioned=None):
        """``Graph::load`` (``src/mat/graph.hpp:104-148``): sniffs the file type (the reference shells out to
        file(1): "ASCII" -> text, "data" -> binary).  With several ranks every rank reads ITS SHARE of the file and the
        entries are redistributed on the device (gt_graph_build_partitioned — the reference's parread + distribute);
        ``partitioned=False`` makes every rank read the whole list and keep its own tiles (gt_graph_build)."""
        Env.init()
        if partitioned is None:
            partitioned = Env.nranks > 1
        if partitioned:
            triples = read_edge_list_share(filepath, self.weighted, Env.rank, Env.nranks)
        else:
            triples = read_edge_list(filepath, self.weighted)
        return self.load_triples(triples, nrows, directed, transpose, self_loops, acyclic, parallel_edges, tiling_type, compression_type,
                                 partitioned=partitioned)

    load_binary = load
    load_text = load

    def load_triples(self, triples: np.ndarray, nvertices, directed=True, transpose=False, self_loops=True, acyclic=False,
                     parallel_edges=True, tiling_type=_2DT_, compression_type=_TCSC_, partitioned=False):
        """``triples``: the global record list, or — ``partitioned`` — this rank's share of it."""
        if tiling_type != _2DT_:
            raise capi.GraphTapError(capi.GT_ERR_UNSUPPORTED, "only _2DT_ tiling is provided (every reference app uses it)")
        Env.init()
        triples = np.ascontiguousarray(triples, dtype="<u4")
        fl = self._flags(directed, transpose, self_loops, acyclic, parallel_edges)
        h = C.c_void_p()
        build = lib().gt_graph_build_partitioned if partitioned else lib().gt_graph_build
        check(build(Env.ctx, triples.ctypes.data_as(C.c_void_p), triples.shape[0], int(self.weighted), 0,
                    int(nvertices), C.byref(fl), int(compression_type), C.byref(h)))
        self.handle = h
        return self

    def load_rmat(self, scale, nedges=None, seed=None, directed=True, transpose=False, self_loops=True, acyclic=False,
                  parallel_edges=True, compression_type=_TCSC_, partitioned=False):
        """Synthetic RMAT input generated on the device (bench tooling; same stream as graphtap_b200.rmat).
        ``partitioned``: every rank generates 1/p of the records and routes them to the tile owners."""
        Env.init()
        fl = self._flags(directed, transpose, self_loops, acyclic, parallel_edges)
        h = C.c_void_p()
        build = lib().gt_graph_build_rmat_partitioned if partitioned else lib().gt_graph_build_rmat
        check(build(Env.ctx, scale, (16 << scale) if nedges is None else nedges,
                    scale if seed is None else seed, int(self.weighted), C.byref(fl),
                    int(compression_type), C.byref(h)))
        self.handle = h
        return self

    def info(self) -> capi.GraphInfo:
        gi = capi.GraphInfo()
        check(lib().gt_graph_info_get(self.handle, C.byref(gi)))
        return gi

    # -- test access to the device-resident tiles (gt_graph_tile_view) --
    def _download(self, ptr, count, dtype):
        out = np.empty(count, dtype=dtype)
        if count:
            check(lib().gt_dev_download(Env.ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes))
        return out

    def tile(self, k: int) -> dict:
        tv = capi.TileView()
        check(lib().gt_graph_tile_view(self.handle, k, C.byref(tv)))
        d = dict(rg=tv.rg, cg=tv.cg, row_slot=tv.row_slot, col_slot=tv.col_slot, nnz=tv.nnz, nnzcols=tv.nnzcols, nnzrows=tv.nnzrows)
        d["JA"] = self._download(tv.JA, tv.nnzcols + 1, "<u4")
        d["IA"] = self._download(tv.IA, tv.nnz, "<u4")
        d["A"] = self._download(tv.A, tv.nnz, "<u4") if tv.A else None
        d["JC"] = self._download(tv.JC, tv.nnzcols, "<u4")
        d["IR"] = self._download(tv.IR, tv.nnzrows, "<u4")
        d["_view"] = tv
        return d

    def tile_cf(self, k: int) -> dict:
        """TCSC_CF_BASE's four computation-filtering lists of local tile k (GT_TCSC_CF graphs)."""
        cv = capi.TileCfView()
        check(lib().gt_graph_tile_cf_view(self.handle, k, C.byref(cv)))
        d = {}
        for kind in range(4):
            d[f"NC{kind}"], d[f"filled{kind}"] = cv.NC[kind], cv.filled[kind]
            d[f"JA{kind}"] = self._download(cv.JA[kind], 2 * cv.NC[kind], "<u4")
            d[f"JC{kind}"] = self._download(cv.JC[kind], cv.NC[kind], "<u4")
        return d

    def classify_lists(self):
        """(regular_rows, source_rows, sink_columns) of the owned segment: local vertex ids (classify_vertices)."""
        ptrs = [C.c_void_p() for _ in range(3)]
        ns = [C.c_uint32() for _ in range(3)]
        check(lib().gt_graph_classify_lists(self.handle, C.byref(ptrs[0]), C.byref(ns[0]), C.byref(ptrs[1]), C.byref(ns[1]), C.byref(ptrs[2]), C.byref(ns[2])))
        return tuple(self._download(p.value, n.value, "<u4") for p, n in zip(ptrs, ns))

    def rowgrp_maps(self, slot: int):
        I, IV, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().gt_graph_rowgrp_maps(self.handle, slot, C.byref(I), C.byref(IV), C.byref(n)))
        th = self.info().layout.tile_height
        return self._download(I.value, th, "u1"), self._download(IV.value, th, "<u4"), n.value

    def colgrp_maps(self, slot: int):
        J, JV, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().gt_graph_colgrp_maps(self.handle, slot, C.byref(J), C.byref(JV), C.byref(n)))
        th = self.info().layout.tile_height
        return self._download(J.value, th, "u1"), self._download(JV.value, th, "<u4"), n.value

    def free(self):
        if self.handle is not None:
            check(lib().gt_graph_free(self.handle))
            self.handle = None


class Vertex_Program:
    """``Vertex_Program<Weight, Integer_Type, Fractional_Type, Vertex_State>`` (``src/vp/vertex_program.hpp:23-62``).

    User-defined messenger/combiner/applicator virtuals cannot run on the device; the five shipped
    programs are recognised by class and mapped to the library's app enums."""
    APP = None
    STATE = None
    _state_label = ""

    def __init__(self, graph: Graph, stationary=False, gather_depends_on_apply=False, apply_depends_on_iter=False,
                 ordering_type=_ROW_):
        if self.APP is None:
            raise capi.GraphTapError(capi.GT_ERR_UNSUPPORTED, "only Deg/PR/BFS/CC/SSSP programs run on the device; there is no CPU fallback")
        self.graph = graph
        self.stationary = bool(stationary)
        self.gather_depends_on_apply = bool(gather_depends_on_apply)
        self.apply_depends_on_iter = bool(apply_depends_on_iter)
        self.ordering_type = ordering_type
        self.root = 0
        self.alpha = 0.15          # src/apps/pr.h:13
        self.tol = 1e-5            # src/apps/pr.h:12
        self.iteration = 0
        self.handle = None

    def _ensure(self):
        if self.handle is None:
            prm = capi.Params(self.alpha, self.tol, int(self.root))
            h = C.c_void_p()
            check(lib().gt_program_create(self.graph.handle, self.APP, int(self.stationary), int(self.gather_depends_on_apply),
                                          int(self.apply_depends_on_iter), int(self.ordering_type), C.byref(prm), C.byref(h)))
            self.handle = h
        return self.handle

    def set(self, name: str, value: float):
        check(lib().gt_program_set(self._ensure(), name.encode(), float(value)))

    def initialize(self, other: "Vertex_Program | None" = None):
        if other is not None:
            check(lib().gt_program_init_from(self._ensure(), other._ensure()))

    def execute(self, num_iterations: int = 0):
        done = C.c_uint32()
        check(lib().gt_program_execute(self._ensure(), int(num_iterations), C.byref(done)))
        self.iteration = done.value
        Env.print_time("Execute", self.timing().execute_ms * 1e-3)
        return self.iteration

    def run_phase(self, phase: int):
        """One phase of one iteration in isolation (0 scatter_gather, 1 combine, 2 apply); `iteration` does not advance."""
        check(lib().gt_program_run_phase(self._ensure(), int(phase)))

    def timing(self) -> capi.Timing:
        t = capi.Timing()
        check(lib().gt_program_timing(self._ensure(), C.byref(t)))
        return t

    @property
    def V(self) -> np.ndarray:
        """The owned segment's vertex states in the reference's AoS layout (public ``V``, :61)."""
        th = self.graph.info().layout.tile_height
        out = np.empty(th, dtype=self.STATE)
        check(lib().gt_program_state_to_host(self._ensure(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_V(self, states: np.ndarray):
        states = np.ascontiguousarray(states, dtype=self.STATE)
        check(lib().gt_program_state_from_host(self._ensure(), states.ctypes.data_as(C.c_void_p), states.nbytes))

    def checksum(self, quiet: bool = False):
        s, c = C.c_uint64(), C.c_uint64()
        check(lib().gt_program_checksum(self._ensure(), C.byref(s), C.byref(c)))
        if Env.is_master and not quiet:          # line formats of :1942-1958, grepped by graphtap.slurm:101-104
            print(f"Iterations: {self.iteration}")
            print(f"Value checksum: {s.value}")
            print(f"Reachable vertices: {c.value}")
        return s.value, c.value

    def display(self, count: int = 31):
        if Env.rank != 0:
            return
        V = self.V
        lay = self.graph.info().layout
        base = lay.owned_segment * lay.tile_height
        for i in range(min(count, len(V))):
            print(f"vertex[{base + i}]:{self._print_state(V[i])}")

    def _print_state(self, s) -> str:
        return str(s)

    def free(self):
        if self.handle is not None:
            check(lib().gt_program_free(self.handle))
            self.handle = None


class Deg_Program(Vertex_Program):
    APP, STATE = GT_APP_DEG, DEG_STATE

    def _print_state(self, s):
        return f"Degree={s['degree']}"


class PR_Program(Vertex_Program):
    APP, STATE = GT_APP_PR, PR_STATE

    def _print_state(self, s):
        return f"Rank={s['rank']:.6f},Degree={s['degree']}"


class BFS_Program(Vertex_Program):
    APP, STATE = GT_APP_BFS, BFS_STATE

    def _print_state(self, s):
        return f"Parent={s['parent']},Hops=" + ("INF" if s["hops"] == INF else str(s["hops"]))


class CC_Program(Vertex_Program):
    APP, STATE = GT_APP_CC, CC_STATE

    def _print_state(self, s):
        return f"Label={s['label']}"


class SSSP_Program(Vertex_Program):
    APP, STATE = GT_APP_SSSP, SSSP_STATE

    def _print_state(self, s):
        return "Distance=" + ("INF" if s["distance"] == INF else str(s["distance"]))


# ---- the reference drivers, as functions (src/apps/{pr,bfs,cc,sssp}.cpp) ---------------------------------
def run_pr(graph_loader, num_iterations=20, compression=_TCSC_CF_, pr_layout=None):
    """src/apps/pr.cpp:26-53.  ``graph_loader(G, **flags)`` loads the edge list into ``G`` with the given flags.
    ``pr_layout``: 0 = push SpMV over the TCSC arrays, 1 = derived pull layout (library default)."""
    G = Graph(weighted=False)
    graph_loader(G, directed=True, transpose=True, self_loops=True, acyclic=False, parallel_edges=True, compression_type=compression)
    V = Deg_Program(G, True, False, False, _COL_)
    V.execute(1)
    VR = PR_Program(G, True, False, False, _ROW_)
    if pr_layout is not None:
        VR.set("pr_layout", pr_layout)
    VR.initialize(V)
    V.free()
    VR.execute(num_iterations)
    return G, VR


def run_bfs(graph_loader, root=0):
    """src/apps/bfs.cpp:26-45"""
    G = Graph(weighted=False)
    graph_loader(G, directed=False, transpose=False, self_loops=False, acyclic=False, parallel_edges=False, compression_type=_TCSC_)
    V = BFS_Program(G, False, False, True, _ROW_)
    V.root = root
    V.execute()
    return G, V


def run_cc(graph_loader):
    """src/apps/cc.cpp:25-44"""
    G = Graph(weighted=False)
    graph_loader(G, directed=False, transpose=False, self_loops=True, acyclic=False, parallel_edges=False, compression_type=_TCSC_)
    V = CC_Program(G, False, True, False, _ROW_)
    V.execute()
    return G, V


def run_sssp(graph_loader, root=0):
    """src/apps/sssp.cpp:26-44 (-DHAS_WEIGHT build; `transpose = not transpose` for the directed non-stationary engine)"""
    G = Graph(weighted=True)
    graph_loader(G, directed=True, transpose=True, self_loops=False, acyclic=False, parallel_edges=False, compression_type=_TCSC_)
    V = SSSP_Program(G, False, True, False, _ROW_)
    V.root = root
    V.execute()
    return G, V


class DeviceArray:
    """A raw device buffer (gt_dev_alloc) with numpy upload/download; used by tests and the bench to
    feed the kernel-level entry points (gt_tile_spmv / gt_tile_spmspv)."""

    def __init__(self, host: np.ndarray | None = None, nbytes: int | None = None):
        Env.init()
        self.nbytes = int(host.nbytes if host is not None else nbytes)
        self.ptr = C.c_void_p()
        check(lib().gt_dev_alloc(Env.ctx, max(self.nbytes, 1), C.byref(self.ptr)))
        if host is not None:
            self.upload(host)

    def upload(self, host: np.ndarray):
        host = np.ascontiguousarray(host)
        assert host.nbytes <= self.nbytes
        check(lib().gt_dev_upload(Env.ctx, self.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))

    def download(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        check(lib().gt_dev_download(Env.ctx, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            check(lib().gt_dev_free(Env.ctx, self.ptr))
            self.ptr = None
