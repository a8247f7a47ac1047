// gt_layout.cpp — the reference's 2D-"transposed" (_2DT_) tile layout, restated from its formulas.
//
// What it reproduces (reference paths relative to the GraphTap repo):
//   grid      p x p tiles, p = #ranks; tile_height = (n+1)/p + 1        src/mat/matrix.hpp:188-194, graph.hpp:89-98
//   tiling    rowgrp_nranks x colgrp_nranks = p, from sqrt(p) upwards   src/mat/tiling.hpp:39-73
//   owners    rank(i,j) = (i % colgrp_nranks)*rowgrp_nranks + (j % rowgrp_nranks)      matrix.hpp:301-302
//   leaders   whole-row swaps until every diagonal tile has a distinct owner           matrix.hpp:327-341
//   locals    row-major / column-major scans, segment lists in first-seen order        matrix.hpp:343-380
//   groups    ranks sharing my diagonal tile's row / column, sorted                    matrix.hpp:382-465
//   accu_*    my position inside those lists                                           matrix.hpp:466-485
// Pure host arithmetic: these entry points work without a GPU, and tests/test_layout.py diffs every
// table against the unmodified reference (oracle/_ref/ref_layout) for p = 1, 2, 4, 8, 16.
#include "gt_internal.h"
#include <algorithm>
#include <cmath>
#include <numeric>

namespace gt {

static bool contains(const std::vector<int32_t>& v, int32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }

Layout make_layout(uint32_t nvertices, int nranks, int rank) {
    GT_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "layout: bad rank/nranks");
    Layout L;
    gt_layout& I = L.info;
    const uint32_t p = (uint32_t) nranks;
    I.nranks = p;
    I.rank = (uint32_t) rank;
    I.nrows = nvertices + 1;
    I.nrowgrps = I.ncolgrps = p;
    I.tile_height = I.nrows / p + 1;
    // Tiling::integer_factorize
    uint32_t a = (uint32_t) std::sqrt((double) p), b = a;
    while (a * b != p) { b++; a = p / b; }
    I.rowgrp_nranks = a;
    I.colgrp_nranks = b;
    I.rank_nrowgrps = p / b;
    I.rank_ncolgrps = p / a;

    // owners, one vector per tile-row so that rows can be swapped whole
    std::vector<std::vector<int32_t>> rows(p, std::vector<int32_t>(p));
    for (uint32_t i = 0; i < p; i++)
        for (uint32_t j = 0; j < p; j++) rows[i][j] = (int32_t) ((i % b) * a + (j % a));
    L.leader_ranks.assign(p, -1);
    for (uint32_t i = 0; i < p; i++) {
        for (uint32_t j = i; j < p; j++) {
            if (!contains(L.leader_ranks, rows[j][i])) { std::swap(rows[j], rows[i]); break; }
        }
        L.leader_ranks[i] = rows[i][i];
    }
    L.tile_rank.resize((size_t) p * p);
    for (uint32_t i = 0; i < p; i++)
        for (uint32_t j = 0; j < p; j++) L.tile_rank[(size_t) i * p + j] = rows[i][j];

    I.owned_segment = -1;
    for (uint32_t i = 0; i < p; i++)
        for (uint32_t j = 0; j < p; j++) {
            if (rows[i][j] != rank) continue;
            L.local_tiles_row_order.push_back((int32_t) (i * p + j));
            if (!contains(L.local_col_segments, (int32_t) j)) L.local_col_segments.push_back((int32_t) j);
            if (!contains(L.local_row_segments, (int32_t) i)) L.local_row_segments.push_back((int32_t) i);
            if (i == j) I.owned_segment = (int32_t) i;
        }
    for (uint32_t j = 0; j < p; j++)
        for (uint32_t i = 0; i < p; i++)
            if (rows[i][j] == rank) L.local_tiles_col_order.push_back((int32_t) (i * p + j));
    GT_REQUIRE(I.owned_segment >= 0, "layout: rank owns no diagonal tile");

    // ranks that share my diagonal tile's row (row group) and column (column group)
    const uint32_t s = (uint32_t) I.owned_segment;
    for (uint32_t j = 0; j < p; j++) {
        int32_t r = rows[s][j];
        if (!contains(L.all_rowgrp_ranks, r)) {
            L.all_rowgrp_ranks.push_back(r);
            if (r != rank) L.follower_rowgrp_ranks.push_back(r);
        }
    }
    for (uint32_t i = 0; i < p; i++) {
        int32_t r = rows[i][s];
        if (!contains(L.all_colgrp_ranks, r)) {
            L.all_colgrp_ranks.push_back(r);
            if (r != rank) L.follower_colgrp_ranks.push_back(r);
        }
    }
    std::sort(L.all_rowgrp_ranks.begin(), L.all_rowgrp_ranks.end());
    std::sort(L.all_colgrp_ranks.begin(), L.all_colgrp_ranks.end());
    std::sort(L.follower_rowgrp_ranks.begin(), L.follower_rowgrp_ranks.end());
    std::sort(L.follower_colgrp_ranks.begin(), L.follower_colgrp_ranks.end());

    I.accu_segment_rg = I.accu_segment_cg = I.accu_segment_row = I.accu_segment_col = -1;
    for (size_t j = 0; j < L.all_rowgrp_ranks.size(); j++) if (L.all_rowgrp_ranks[j] == rank) I.accu_segment_rg = (int32_t) j;
    for (size_t j = 0; j < L.all_colgrp_ranks.size(); j++) if (L.all_colgrp_ranks[j] == rank) I.accu_segment_cg = (int32_t) j;
    for (size_t j = 0; j < L.local_row_segments.size(); j++) if (L.leader_ranks[L.local_row_segments[j]] == rank) I.accu_segment_row = (int32_t) j;
    for (size_t j = 0; j < L.local_col_segments.size(); j++) if (L.leader_ranks[L.local_col_segments[j]] == rank) I.accu_segment_col = (int32_t) j;
    return L;
}

// The exchange plan of a partitioned ingest (gt_build.cu) for one rank, from the world's count matrix
// counts[r * nranks + q] = entries rank r holds for rank q.  Blocks sit in destination order in the send buffer and in
// sender order in every receive buffer — so a block from `rank` lands in q's buffer behind the blocks of the ranks
// before it, which is where q's own plan expects it.  Host arithmetic only.
RoutePlan make_route_plan(int nranks, int rank, const uint64_t* counts) {
    RoutePlan P;
    P.send_offset.assign(nranks, 0); P.recv_offset.assign(nranks, 0); P.remote_offset.assign(nranks, 0);
    for (int q = 0; q < nranks; q++) {
        P.send_offset[q] = P.nsend; P.nsend += counts[(size_t) rank * nranks + q];
        P.recv_offset[q] = P.nrecv; P.nrecv += counts[(size_t) q * nranks + rank];
        uint64_t col = 0;
        for (int r = 0; r < nranks; r++) {
            if (r == rank) P.remote_offset[q] = col;
            col += counts[(size_t) r * nranks + q];
        }
        P.max_recv = std::max(P.max_recv, col);
    }
    return P;
}

}  // namespace gt

extern "C" int gt_ingest_route_plan(int nranks, int rank, const uint64_t* counts, uint64_t* send_offset, uint64_t* recv_offset,
                                    uint64_t* remote_offset, uint64_t* nsend, uint64_t* nrecv, uint64_t* max_recv) {
    return gt::guarded([&] {
        GT_REQUIRE(counts && nranks >= 1 && rank >= 0 && rank < nranks, "gt_ingest_route_plan: bad arguments");
        const gt::RoutePlan P = gt::make_route_plan(nranks, rank, counts);
        for (int q = 0; q < nranks; q++) {
            if (send_offset) send_offset[q] = P.send_offset[q];
            if (recv_offset) recv_offset[q] = P.recv_offset[q];
            if (remote_offset) remote_offset[q] = P.remote_offset[q];
        }
        if (nsend) *nsend = P.nsend;
        if (nrecv) *nrecv = P.nrecv;
        if (max_recv) *max_recv = P.max_recv;
    });
}

extern "C" int gt_layout_query(uint32_t nvertices, int nranks, int rank, gt_layout* out) {
    return gt::guarded([&] {
        GT_REQUIRE(out, "gt_layout_query: out is NULL");
        *out = gt::make_layout(nvertices, nranks, rank).info;
    });
}

extern "C" int gt_layout_table(uint32_t nvertices, int nranks, int rank, int which, int32_t* out, uint32_t cap, uint32_t* n) {
    return gt::guarded([&] {
        gt::Layout L = gt::make_layout(nvertices, nranks, rank);
        const std::vector<int32_t>* v = nullptr;
        switch (which) {
            case GT_LT_TILE_RANK: v = &L.tile_rank; break;
            case GT_LT_LEADER_RANKS: v = &L.leader_ranks; break;
            case GT_LT_LOCAL_TILES_ROW_ORDER: v = &L.local_tiles_row_order; break;
            case GT_LT_LOCAL_TILES_COL_ORDER: v = &L.local_tiles_col_order; break;
            case GT_LT_LOCAL_ROW_SEGMENTS: v = &L.local_row_segments; break;
            case GT_LT_LOCAL_COL_SEGMENTS: v = &L.local_col_segments; break;
            case GT_LT_ALL_ROWGRP_RANKS: v = &L.all_rowgrp_ranks; break;
            case GT_LT_ALL_COLGRP_RANKS: v = &L.all_colgrp_ranks; break;
            case GT_LT_FOLLOWER_ROWGRP_RANKS: v = &L.follower_rowgrp_ranks; break;
            case GT_LT_FOLLOWER_COLGRP_RANKS: v = &L.follower_colgrp_ranks; break;
            default: throw gt::Error(GT_ERR_INVALID, "gt_layout_table: unknown table id");
        }
        if (n) *n = (uint32_t) v->size();
        if (out) {
            GT_REQUIRE(cap >= v->size(), "gt_layout_table: output capacity too small");
            std::copy(v->begin(), v->end(), out);
        }
    });
}
