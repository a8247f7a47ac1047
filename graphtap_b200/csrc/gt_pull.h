// gt_pull.h — derived pull layout of the stationary plus-times SpMV (see gt_pull.cu).
#pragma once
#include "gt_kernels.cuh"

namespace gt {

constexpr uint32_t kPullMaxSegs = 8;            // local column segments per rank (rank_ncolgrps: 1,2,2,4 at p = 1,2,4,8)
constexpr uint32_t kPullVRow = 512;             // longest run of entries one lane sums before the row is split (GT_PULL_VROW)
constexpr uint32_t kPullSplit = 0x80000000u;    // vtgt flag: partial sum of a split row -> RED.ADD
constexpr uint32_t kPullHotBit = 0x80000000u;   // column-code flag: one of the hottest columns -> L1-allocating gather
constexpr int kPullThreads = 1024;

struct PullSell {                               // one SELL-32 array
    uint32_t nv = 0, nslices = 0;
    uint64_t nnz = 0, sell_len = 0;
    DevBuf<uint32_t> sell;                      // column codes, slice-major then column-major
    DevBuf<uint64_t> slice_ptr;                 // [nslices_all + 1]
    DevBuf<uint32_t> vtgt;                      // [nv] y index (| kPullSplit)
};

struct PullRows {                               // one local row segment
    uint32_t ny = 0;                            // length of its y vector = vertices in the segment's hot order
    uint64_t nnz = 0;
    // Multi-GPU: entries whose column lies in this rank's OWN x chunk are kept apart, so that part of the SpMV can
    // run while the all-gather of the other chunks is still in flight.  Single GPU: `own` is empty.
    PullSell own, rest;
    // _TCSC_CF_ graphs (computation filtering, src/vp/vertex_program.hpp:1218-1325): `own` / `rest` hold only the entries
    // of REGULAR rows in REGULAR columns, which every iteration needs; `snk` = regular rows x sink columns (iteration 0
    // only), `src` = source rows x every column (last iteration only).  Empty on _TCSC_ graphs.
    PullSell snk, src;
    uint64_t nnz_rr = 0;                        // entries of own + rest
};

struct PullLayout {
    // Concatenated x: one equal-sized chunk per member of the column group, chunk q = the segment led by group
    // rank q, so that the whole exchange is ONE in-place ncclAllGather (the reference: one Ibcast per segment,
    // src/vp/vertex_program.hpp:843-862).  Same for y and the row group with ncclReduceScatter.
    std::vector<uint32_t> xoff, xn;             // per column slot: chunk start, vertices in the segment's hot order
    uint32_t xchunk = 0, xlen = 0;              // chunk size, total; x[xlen] is a permanent 0.0 (padding target)
    std::vector<uint32_t> xreg, xsnk0;          // per column slot: x positions [0, xreg) regular, [xsnk0, xn) sink columns (_TCSC_: xreg = xn)
    std::vector<uint32_t> yoff, yn;             // per row slot
    std::vector<uint32_t> yreg, ysrc;           // per row slot: y positions [0, yreg) regular rows, [yreg, yreg + ysrc) source rows
    bool cf = false;                            // the graph is _TCSC_CF_: rows / columns are split as above
    uint32_t ychunk = 0, ylen = 0;
    std::vector<PullRows> rows;                 // per row slot
    uint32_t vrow = kPullVRow;                  // tuning knobs, fixed at build time (GT_PULL_* environment)
    uint32_t band = 0;                          // single GPU: columns [0, band) of the hot order form a pass of their own (0 = one pass)
    bool band_smem = false;                     // GT_PULL_BAND_SMEM: that pass gathers from a shared-memory copy of x[0, band)
    uint32_t l1hot = 0;                         // hottest columns (per rank) gathered L1::evict_last, the rest L1::evict_first; 0 = no distinction
    bool l2hint = true;                         // L2 eviction hints: index stream evict-first, x evict-last (-8 % on RMAT-26)
    int unroll = 8;
    int threads = kPullThreads, ctas_per_sm = 2;
};

PullLayout* pull_build(gt_graph* g);
void pull_free(PullLayout* P);
// part 0: the rank's own x chunk (y = ...), part 1: everything else (y += ... when part 0 exists),
// part 2: regular rows x sink columns (y += ...), part 3: source rows (y = ...)
void pull_spmv(gt_ctx* ctx, const PullLayout* P, uint32_t row_slot, int part, const double* x, double* y);

}  // namespace gt
