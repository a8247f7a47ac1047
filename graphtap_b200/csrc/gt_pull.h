// gt_pull.h — derived pull layout of the stationary plus-times SpMV (see gt_pull.cu).
#pragma once
#include "gt_kernels.cuh"

namespace gt {

constexpr uint32_t kPullMaxSegs = 8;            // local column segments per rank (rank_ncolgrps: 1,2,2,4 at p = 1,2,4,8)
constexpr uint32_t kPullHotDoubles = 25600;     // 200 KB of shared memory for the hot x values
constexpr uint32_t kPullVRow = 2048;            // longest run of entries one lane sums before the row is split
constexpr uint32_t kPullSplit = 0x80000000u;    // vtgt flag: partial sum of a split row -> RED.ADD
constexpr int kPullThreads = 1024;

struct PullHot {                                // passed by value to the kernel
    uint32_t total, per_seg;
    uint32_t xoff[kPullMaxSegs];                // start of segment s in the concatenated x buffer
    uint32_t seg_len[kPullMaxSegs];
};

struct PullRows {                               // one local row segment
    uint32_t nrows = 0, nv = 0, nslices = 0;
    uint64_t nnz = 0, sell_len = 0;
    DevBuf<uint32_t> sell;                      // SELL-32 column codes, slice-major then column-major
    DevBuf<uint64_t> slice_ptr;                 // [nslices_all + 1]
    DevBuf<uint32_t> vtgt;                      // [nv] hot row id (| kPullSplit)
};

struct PullLayout {
    std::vector<uint32_t> xoff;                 // [S + 1] concatenated x offsets
    uint32_t xlen = 0;
    PullHot hot{};
    std::vector<DevBuf<uint32_t>> col_rank, col_hot_local, col_code;   // per column slot
    std::vector<DevBuf<uint32_t>> row_rank, row_hot_local;             // per row slot
    std::vector<PullRows> rows;
};

PullLayout* pull_build(gt_graph* g);
void pull_free(PullLayout* P);
void pull_spmv(gt_ctx* ctx, const PullLayout* P, uint32_t row_slot, const double* x, double* y);

}  // namespace gt
