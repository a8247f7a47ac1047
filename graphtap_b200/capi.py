"""ctypes binding of ``libgraphtap_b200.so`` (the C ABI declared in ``include/graphtap_b200.h``).

This module is deliberately thin: it declares the prototypes, turns non-zero status codes into
``GraphTapError`` and nothing else.  The host-side mirror of the reference's ``Graph`` /
``Vertex_Program`` interface lives in ``graphtap_b200.engine``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GT_LIB: another build of the same library (tools/build_variants.sh: compile-time A/B experiments), never a fallback
LIB_PATH = os.environ.get("GT_LIB") or os.path.join(_HERE, "libgraphtap_b200.so")

# enums (include/graphtap_b200.h)
GT_OK, GT_ERR_INVALID, GT_ERR_NO_DEVICE, GT_ERR_CUDA, GT_ERR_NCCL, GT_ERR_OOM, GT_ERR_UNSUPPORTED = range(7)
GT_TCSC, GT_TCSC_CF = 2, 3
GT_ROW, GT_COL = 0, 1
GT_APP_DEG, GT_APP_PR, GT_APP_BFS, GT_APP_CC, GT_APP_SSSP = range(5)
GT_PLUS_TIMES_F64, GT_MIN_PLUS_U32, GT_MIN_SELECT_U32 = range(3)
GT_INF_U32 = 2147483647
(GT_LT_TILE_RANK, GT_LT_LEADER_RANKS, GT_LT_LOCAL_TILES_ROW_ORDER, GT_LT_LOCAL_TILES_COL_ORDER,
 GT_LT_LOCAL_ROW_SEGMENTS, GT_LT_LOCAL_COL_SEGMENTS, GT_LT_ALL_ROWGRP_RANKS, GT_LT_ALL_COLGRP_RANKS,
 GT_LT_FOLLOWER_ROWGRP_RANKS, GT_LT_FOLLOWER_COLGRP_RANKS) = range(10)


class GraphTapError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[gt status {code}] {msg}")
        self.code = code


class Layout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("nranks", "rank", "nrows", "nrowgrps", "ncolgrps", "tile_height",
                                          "rowgrp_nranks", "colgrp_nranks", "rank_nrowgrps", "rank_ncolgrps")] + \
               [(n, C.c_int32) for n in ("owned_segment", "accu_segment_rg", "accu_segment_cg",
                                         "accu_segment_row", "accu_segment_col")]


class GraphFlags(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("directed", "transpose", "self_loops", "acyclic", "parallel_edges")]


class GraphInfo(C.Structure):
    _fields_ = [("nedges_input", C.c_uint64), ("nnz_local", C.c_uint64), ("nnz_global", C.c_uint64),
                ("ntiles_local", C.c_uint32), ("weighted", C.c_uint32), ("nvertices", C.c_uint32), ("layout", Layout)]


class TileView(C.Structure):
    _fields_ = [("rg", C.c_uint32), ("cg", C.c_uint32), ("row_slot", C.c_uint32), ("col_slot", C.c_uint32),
                ("nnz", C.c_uint64), ("nnzcols", C.c_uint32), ("nnzrows", C.c_uint32),
                ("JA", C.c_void_p), ("IA", C.c_void_p), ("A", C.c_void_p), ("JC", C.c_void_p), ("IR", C.c_void_p)]


class TileCfView(C.Structure):
    _fields_ = [("NC", C.c_uint32 * 4), ("filled", C.c_uint32 * 4), ("JA", C.c_void_p * 4), ("JC", C.c_void_p * 4)]


class Params(C.Structure):
    _fields_ = [("alpha", C.c_double), ("tol", C.c_double), ("root", C.c_uint32)]


class Timing(C.Structure):
    _fields_ = [("execute_ms", C.c_double), ("scatter_gather_ms", C.c_double), ("combine_ms", C.c_double),
                ("apply_ms", C.c_double), ("kernel_launches", C.c_uint64), ("bytes_algorithmic", C.c_uint64),
                ("iterations", C.c_uint32), ("sparse_iterations", C.c_uint32), ("combine_bytes", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/graphtap_b200.h declares
PROTOTYPES = {
    "gt_last_error": (C.c_char_p, []),
    "gt_abi_version": (C.c_int, []),
    "gt_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "gt_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "gt_ctx_destroy": (C.c_int, [C.c_void_p]),
    "gt_ctx_sync": (C.c_int, [C.c_void_p]),
    "gt_ctx_barrier": (C.c_int, [C.c_void_p]),
    "gt_ctx_timer_begin": (C.c_int, [C.c_void_p]),
    "gt_ctx_timer_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "gt_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "gt_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "gt_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "gt_dev_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "gt_dev_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "gt_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "gt_host_free_pinned": (C.c_int, [C.c_void_p]),
    "gt_layout_query": (C.c_int, [C.c_uint32, C.c_int, C.c_int, C.POINTER(Layout)]),
    "gt_layout_table": (C.c_int, [C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_uint32, C.POINTER(C.c_uint32)]),
    "gt_graph_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.POINTER(GraphFlags), C.c_int, C.POINTER(C.c_void_p)]),
    "gt_graph_build_rmat": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(GraphFlags), C.c_int, C.POINTER(C.c_void_p)]),
    "gt_graph_build_partitioned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.POINTER(GraphFlags), C.c_int, C.POINTER(C.c_void_p)]),
    "gt_graph_build_rmat_partitioned": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(GraphFlags), C.c_int, C.POINTER(C.c_void_p)]),
    "gt_ingest_route_plan": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gt_rmat_generate": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]),
    "gt_graph_free": (C.c_int, [C.c_void_p]),
    "gt_graph_info_get": (C.c_int, [C.c_void_p, C.POINTER(GraphInfo)]),
    "gt_graph_tile_view": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TileView)]),
    "gt_graph_rowgrp_maps": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "gt_graph_colgrp_maps": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "gt_graph_classify": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "gt_graph_classify_lists": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "gt_graph_tile_cf_view": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TileCfView)]),
    "gt_tile_spmv": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gt_tile_spmspv": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gt_program_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]),
    "gt_program_free": (C.c_int, [C.c_void_p]),
    "gt_program_init_from": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_program_execute": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "gt_program_state_bytes": (C.c_uint32, [C.c_void_p]),
    "gt_program_state_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_program_state_from_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_program_checksum": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gt_program_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "gt_program_timing_samples": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_uint32)]),
    "gt_program_run_phase": (C.c_int, [C.c_void_p, C.c_int]),
    "gt_program_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the library (built in-tree by ``__graft_entry__.build()``) and bind the prototypes.
    Fails loudly if the extension is missing: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GraphTapError(GT_ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                                  "(graphtap_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != GT_OK:
        raise GraphTapError(status, lib().gt_last_error().decode("utf-8", "replace"))
