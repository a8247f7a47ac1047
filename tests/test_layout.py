"""gt_layout_* (the product's host-side restatement of Matrix::init_matrix) against the tables the
unmodified reference builds for every (p, rank) — tests/golden/layouts.json, from oracle/_ref/ref_layout."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from graphtap_b200 import capi
from oracle import oracle as O

LAY = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "layouts.json")))


def table(n, p, r, which):
    cnt = C.c_uint32()
    capi.check(capi.lib().gt_layout_table(n, p, r, which, None, 0, C.byref(cnt)))
    out = (C.c_int32 * max(1, cnt.value))()
    capi.check(capi.lib().gt_layout_table(n, p, r, which, out, cnt.value, C.byref(cnt)))
    return list(out[: cnt.value])


@pytest.mark.parametrize("key", sorted(LAY, key=lambda k: tuple(map(int, k.split(":")))))
def test_layout_matches_reference(key):
    p, r = map(int, key.split(":"))
    ref = LAY[key]
    lay = capi.Layout()
    capi.check(capi.lib().gt_layout_query(1024, p, r, C.byref(lay)))
    assert lay.tile_height == ref["tile_height"][0]
    assert [lay.nrowgrps, lay.ncolgrps, lay.rowgrp_nranks, lay.colgrp_nranks, lay.rank_nrowgrps, lay.rank_ncolgrps] == ref["grid"]
    assert lay.owned_segment == ref["owned_segment"][0]
    assert [lay.accu_segment_rg, lay.accu_segment_cg, lay.accu_segment_row, lay.accu_segment_col] == ref["accu"]
    for which, name in ((capi.GT_LT_TILE_RANK, "tile_rank"), (capi.GT_LT_LEADER_RANKS, "leader_ranks"),
                        (capi.GT_LT_LOCAL_TILES_ROW_ORDER, "local_tiles_row_order"), (capi.GT_LT_LOCAL_TILES_COL_ORDER, "local_tiles_col_order"),
                        (capi.GT_LT_LOCAL_ROW_SEGMENTS, "local_row_segments"), (capi.GT_LT_LOCAL_COL_SEGMENTS, "local_col_segments"),
                        (capi.GT_LT_ALL_ROWGRP_RANKS, "all_rowgrp_ranks"), (capi.GT_LT_ALL_COLGRP_RANKS, "all_colgrp_ranks"),
                        (capi.GT_LT_FOLLOWER_ROWGRP_RANKS, "follower_rowgrp_ranks"), (capi.GT_LT_FOLLOWER_COLGRP_RANKS, "follower_colgrp_ranks")):
        assert table(1024, p, r, which) == ref.get(name, []), name
    # group-local ranks used as bcast roots / reduce targets equal the reference's rank_rg / rank_cg
    assert table(1024, p, r, capi.GT_LT_ALL_ROWGRP_RANKS).index(r) == ref["rank_rg_cg"][0]
    assert table(1024, p, r, capi.GT_LT_ALL_COLGRP_RANKS).index(r) == ref["rank_rg_cg"][1]


@pytest.mark.parametrize("p", [1, 2, 4, 8, 16])
def test_oracle_layout_agrees(p):
    o = O.layout(1024, p)
    assert list(o["tile_rank"].ravel()) == LAY[f"{p}:0"]["tile_rank"]
    assert list(o["leader_ranks"]) == LAY[f"{p}:0"]["leader_ranks"]


def test_survey_probe_layouts():
    # SURVEY.md §8(a) D5 [probe]: p=4 grid and leaders, p=8 leaders
    assert LAY["4:0"]["tile_rank"] == [0, 1, 0, 1, 2, 3, 2, 3, 2, 3, 2, 3, 0, 1, 0, 1]
    assert LAY["4:0"]["leader_ranks"] == [0, 3, 2, 1]
    assert LAY["8:0"]["leader_ranks"] == [0, 3, 4, 7, 2, 1, 6, 5]


def test_group_consistency():
    """What the NCCL schedule relies on: all ranks of a column (row) group hold the same column (row)
    segments in the same order, and every segment's leader is a member of the group."""
    for p in (2, 4, 8, 16):
        for r in range(p):
            cg = table(1024, p, r, capi.GT_LT_ALL_COLGRP_RANKS)
            rg = table(1024, p, r, capi.GT_LT_ALL_ROWGRP_RANKS)
            lead = table(1024, p, r, capi.GT_LT_LEADER_RANKS)
            mycols = table(1024, p, r, capi.GT_LT_LOCAL_COL_SEGMENTS)
            myrows = table(1024, p, r, capi.GT_LT_LOCAL_ROW_SEGMENTS)
            for q in cg:
                assert table(1024, p, q, capi.GT_LT_LOCAL_COL_SEGMENTS) == mycols
                assert table(1024, p, q, capi.GT_LT_ALL_COLGRP_RANKS) == cg
            for q in rg:
                assert table(1024, p, q, capi.GT_LT_LOCAL_ROW_SEGMENTS) == myrows
                assert table(1024, p, q, capi.GT_LT_ALL_ROWGRP_RANKS) == rg
            assert all(lead[s] in cg for s in mycols) and all(lead[s] in rg for s in myrows)
