// gt_graph.h — device-resident graph: 2DT tiles in TCSC + the group-wide index maps.
#pragma once
#include "gt_internal.h"

namespace gt {

// Index maps of one local row segment (I/IV) or column segment (J/JV):
// reference src/mat/matrix.hpp:82-85 and filter_vertices :860-1122.
struct SegMaps {
    int32_t segment = -1;          // global segment id
    uint32_t nnz = 0;              // group-wide non-empty count = length of the compressed vector
    DevBuf<uint8_t> bits;          // I / J   [tile_height]  1 = non-empty anywhere in the group
    DevBuf<uint32_t> prefix;       // IV / JV [tile_height]  compressed id, 0 where empty
    DevBuf<uint32_t> ids;          // IR / JC [nnz]          compressed id -> local id
};

// "Hot order" of one vertex segment: the vertices that have any entry in their row or column, by
// decreasing global degree (in + out, counted at ingest over the whole edge list, so every rank derives
// the same order without communication).  The pull layout of the plus-times SpMV indexes both x and y of
// the segment in this order (gt_pull.cu).
// On a _TCSC_CF_ graph the order is by vertex class first — regular (row and column non-empty), then source rows (row
// only), then sink columns (column only), src/mat/matrix.hpp:1135-1144 — and by degree inside a class, so the three
// sets the computation-filtering schedule treats differently are three contiguous ranges of every x and y vector.
struct HotOrder {
    int32_t segment = -1;
    uint32_t n = 0;                // vertices in the order
    uint32_t nreg = 0, nsrc = 0;   // _TCSC_CF_: [0, nreg) regular, [nreg, nreg + nsrc) source rows, the rest sink columns; else nreg = n
    DevBuf<uint32_t> ids;          // [n]           position -> local vertex id
    DevBuf<uint32_t> pos;          // [tile_height] local vertex id -> position, 0xffffffff if absent
};

struct Tile {
    uint32_t rg = 0, cg = 0, row_slot = 0, col_slot = 0;
    uint64_t nnz = 0;
    uint64_t offset = 0;           // into IA_pool / A_pool, multiple of 4 entries (128-bit loads)
    DevBuf<uint32_t> JA;           // [cols[col_slot].nnz + 1]
    DevBuf<uint32_t> chunk_col;    // first column of every GT_PUSH_CHUNK-edge chunk (+ sentinel)
    uint32_t max_col_entries = 0;  // longest column of the tile (decides whether the heavy-column path can trigger)
};

// TCSC_CF_BASE's computation-filtering lists of one tile (src/ds/compressed_column.hpp:749-1114):
// kind 0 REG_R_REG_C, 1 REG_R_SNK_C, 2 SRC_R_REG_C, 3 SRC_R_SNK_C; NC pairs (start, end) into IA + NC compressed column ids.
struct CfTile {
    uint32_t NC[4] = {0, 0, 0, 0}, filled[4] = {0, 0, 0, 0};
    DevBuf<uint32_t> JA[4], JC[4];
};
// classify_vertices of the owned segment (src/mat/matrix.hpp:1124-1144, :853-855): local vertex ids
struct CfOwned {
    DevBuf<uint32_t> regular_rows, source_rows, sink_columns;
    uint32_t nreg = 0, nsrc = 0, nsnk = 0;
};

struct PullLayout;
void pull_free(PullLayout* P);

}  // namespace gt

struct gt_graph {
    gt_ctx* ctx = nullptr;
    gt::Layout lay;
    gt_graph_flags flags{};
    int weighted = 0;
    int compression = GT_TCSC;
    uint32_t nvertices = 0;
    uint64_t nedges_input = 0, nnz_local = 0, nnz_global = 0;
    std::vector<gt::SegMaps> rows, cols;     // by local slot
    std::vector<gt::Tile> tiles;             // local_tiles_row_order
    gt::DevBuf<uint32_t> IA_pool, A_pool;    // concatenated per-tile IA / A
    gt::DevBuf<uint2> heavy_list;            // frontier SpMSpV scratch: (frontier position, chunk) of heavy columns + a counter
    gt::DevBuf<unsigned int> heavy_count;
    std::vector<gt::DevBuf<uint32_t>> col_deg;   // per local column slot: entries in the vertex's column over the whole matrix (raw records)
    std::vector<uint64_t> col_edges;         // their sum
    std::vector<gt::CfTile> cf_tiles;        // _TCSC_CF_ only, parallel to `tiles`
    gt::CfOwned cf_owned;                    // _TCSC_CF_ only
    std::vector<gt::DevBuf<uint8_t>> cls;    // _TCSC_CF_ only, per distinct local segment (same index as `hot`): 1 regular, 2 source row, 3 sink column
    std::vector<gt::HotOrder> hot;           // one per distinct local segment
    std::vector<int> hot_of_row_slot, hot_of_col_slot;
    gt::PullLayout* pull = nullptr;          // derived layout of the plus-times SpMV, built on first use (gt_pull.cu)
    ~gt_graph() { if (pull) gt::pull_free(pull); }
};

// edges per CTA work item of the load-balanced push kernels (gt_kernels.cu)
#define GT_PUSH_CHUNK 2048
